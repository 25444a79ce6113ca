// K2b: haplotype edit distance (compute_haplotype_edit_distance, SVIM_clustering.py:32-45;
// edlib.align NW distance) as a bit-parallel Myers/Hyyro kernel.
//
// One warp per insertion pair.  The longer haplotype is the pattern: its rows are cut into
// 64-bit words, WPL consecutive words per lane, 32 lanes = one strip of 32*WPL*64 rows; longer
// patterns take several strips with the horizontal deltas of the strip's bottom row parked in a
// per-warp column buffer.  Lanes run as a systolic wavefront: at step s lane l handles text
// column s-l and hands its bottom horizontal delta to lane l+1 with one shuffle.
// Symbol equality is computed from bit-planes of a compact bijective symbol code (no
// shared-memory Peq table): 2 planes when both haplotypes are pure ACGT, 3 with N, 8 otherwise,
// so arbitrary bytes stay exact.
//
// Integer-ALU bound (~30 instructions per 64-cell word step); DRAM traffic is the two
// haplotypes per pair.  No tensor cores: there is no dense contraction here.
#pragma once
#include "ctx.cuh"

#define MYERS_WPL 4

struct MyersWork { uint32_t a, b, slot, pad; };

struct HapSource {
    const uint8_t* p1; int64_t l1;   // reference left of the insertion point
    const uint8_t* p2; int64_t l2;   // inserted sequence
    const uint8_t* p3; int64_t l3;   // reference right of it
};

__constant__ uint8_t c_symcode[256];

static void myers_init_symcode() {
    uint8_t t[256];
    for (int i = 0; i < 256; ++i) t[i] = (uint8_t)i;
    const uint8_t acgtn[5] = {'A', 'C', 'G', 'T', 'N'};
    for (int k = 0; k < 5; ++k) { uint8_t x = t[k]; t[k] = t[acgtn[k]]; t[acgtn[k]] = x; }   // swap -> bijection
    // after the swaps t['A']=0.. and t[0]='A'..; apply .upper() folding on the input side
    uint8_t lut[256];
    for (int i = 0; i < 256; ++i) { int c = (i >= 'a' && i <= 'z') ? i - 32 : i; lut[i] = t[c]; }
    cudaMemcpyToSymbol(c_symcode, lut, 256);
}

__device__ __forceinline__ int64_t clampi(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// hap = ref[max(0,ws):max(0,st)] + seq + ref[max(0,st):max(0,we)], fetch clamped to the contig
__device__ __forceinline__ HapSource make_hap(const uint8_t* contig, int64_t clen, int64_t ws, int64_t we, int64_t st,
                                              const uint8_t* ins, int64_t ins_len) {
    int64_t a = clampi(ws, 0, clen), b = clampi(st, 0, clen), c = clampi(we, 0, clen);
    HapSource h;
    h.p1 = contig + a; h.l1 = b > a ? b - a : 0;
    h.p2 = ins; h.l2 = ins_len;
    h.p3 = contig + b; h.l3 = c > b ? c - b : 0;
    return h;
}

// materialise symbol codes; returns OR of all codes (warp-uniform)
__device__ __forceinline__ uint32_t hap_write_codes(const HapSource& h, uint8_t* dst, int lane) {
    uint32_t orall = 0;
    const int64_t n = h.l1 + h.l2 + h.l3;
    for (int64_t k = lane; k < n; k += 32) {
        uint8_t c = k < h.l1 ? h.p1[k] : (k < h.l1 + h.l2 ? h.p2[k - h.l1] : h.p3[k - h.l1 - h.l2]);
        uint8_t code = c_symcode[c];
        dst[k] = code; orall |= code;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) orall |= __shfl_xor_sync(0xffffffffu, orall, o);
    return orall;
}

template <int NP>
__device__ int32_t myers_run(const uint8_t* __restrict__ pat, int64_t m, const uint8_t* __restrict__ txt, int64_t n,
                             int8_t* __restrict__ hbuf, int lane) {
    const int64_t W = (m + 63) >> 6;
    int64_t score = 0;
    const int64_t STRIP = 32 * MYERS_WPL;
    for (int64_t sb = 0; sb < W; sb += STRIP) {
        const bool first_strip = (sb == 0), last_strip = (sb + STRIP >= W);
        const int64_t ws_cnt = (W - sb) < STRIP ? (W - sb) : STRIP;
        const int nl = (int)((ws_cnt + MYERS_WPL - 1) / MYERS_WPL);   // active lanes
        uint64_t pl[MYERS_WPL][NP], vm[MYERS_WPL], Pv[MYERS_WPL], Mv[MYERS_WPL];
#pragma unroll
        for (int k = 0; k < MYERS_WPL; ++k) {
            const int64_t row0 = (sb + (int64_t)lane * MYERS_WPL + k) * 64;
            vm[k] = 0; Pv[k] = ~0ull; Mv[k] = 0;
#pragma unroll
            for (int b = 0; b < NP; ++b) pl[k][b] = 0;
            if (row0 < m) {
                const int cnt = (m - row0) < 64 ? (int)(m - row0) : 64;
                for (int r = 0; r < cnt; ++r) {
                    const uint64_t code = pat[row0 + r];
#pragma unroll
                    for (int b = 0; b < NP; ++b) pl[k][b] |= ((code >> b) & 1ull) << r;
                }
                vm[k] = cnt == 64 ? ~0ull : ((1ull << cnt) - 1ull);
            }
        }
        // where the pattern's last row lives (last strip only)
        const int64_t wl = W - 1 - sb;
        const int l_last = (int)(wl / MYERS_WPL), k_last = (int)(wl % MYERS_WPL), bit_last = (int)((m - 1) & 63);
        int carry = 0;      // hout of this lane's last word at the previous step, for lane+1
        const int64_t steps = n + nl - 1;
        uint8_t c_next = (lane == 0 && n > 0) ? txt[0] : 0;
        int8_t h_next = (!first_strip && lane == 0 && n > 0) ? hbuf[0] : 0;
        for (int64_t s = 0; s < steps; ++s) {
            const int recv = __shfl_up_sync(0xffffffffu, carry, 1);
            const int64_t j = s - lane;
            const bool act = (lane < nl) && j >= 0 && j < n;
            const uint8_t c = c_next; const int8_t hb = h_next;
            // prefetch for the next step (column j+1)
            if (lane < nl && j + 1 >= 0 && j + 1 < n) {
                c_next = txt[j + 1];
                if (!first_strip && lane == 0) h_next = hbuf[j + 1];
            }
            if (act) {
                int hin = lane == 0 ? (first_strip ? 1 : (int)hb) : recv;
                uint64_t mk[NP];
#pragma unroll
                for (int b = 0; b < NP; ++b) mk[b] = 0ull - (uint64_t)((c >> b) & 1u);
#pragma unroll
                for (int k = 0; k < MYERS_WPL; ++k) {
                    uint64_t Eq = vm[k];
#pragma unroll
                    for (int b = 0; b < NP; ++b) Eq &= ~(pl[k][b] ^ mk[b]);
                    const uint64_t pv = Pv[k], mv = Mv[k];
                    const uint64_t hneg = hin < 0 ? 1ull : 0ull, hpos = hin > 0 ? 1ull : 0ull;
                    const uint64_t Xv = Eq | mv;
                    Eq |= hneg;
                    const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                    uint64_t Ph = mv | ~(Xh | pv);
                    uint64_t Mh = pv & Xh;
                    if (last_strip && lane == l_last && k == k_last) score += (int64_t)((Ph >> bit_last) & 1ull) - (int64_t)((Mh >> bit_last) & 1ull);
                    const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
                    Ph = (Ph << 1) | hpos; Mh = (Mh << 1) | hneg;
                    Pv[k] = Mh | ~(Xv | Ph);
                    Mv[k] = Ph & Xv;
                    hin = hout;
                }
                carry = hin;
                if (!last_strip && lane == 31) hbuf[j] = (int8_t)hin;
            }
        }
        __syncwarp();
    }
    // score lives on lane l_last of the last strip
    const int64_t wl = (W - 1) % STRIP;
    const int src = (int)(wl / MYERS_WPL);
    score = __shfl_sync(0xffffffffu, score, src);
    return (int32_t)(m + score);
}

// haplotype pair -> edit distance; both haplotypes materialised as symbol codes in `scratch`
__device__ int32_t myers_pair(const HapSource& ha, const HapSource& hb, uint8_t* scratch, int64_t maxlen, int lane) {
    const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
    if (la == 0) return (int32_t)lb;
    if (lb == 0) return (int32_t)la;
    const bool a_is_pat = la >= lb;
    const HapSource& hp = a_is_pat ? ha : hb;
    const HapSource& ht = a_is_pat ? hb : ha;
    const int64_t m = a_is_pat ? la : lb, n = a_is_pat ? lb : la;
    uint8_t* pat = scratch; uint8_t* txt = scratch + maxlen; int8_t* hbuf = (int8_t*)(scratch + 2 * maxlen);
    uint32_t orall = hap_write_codes(hp, pat, lane) | hap_write_codes(ht, txt, lane);
    __syncwarp();
    if (orall < 4) return myers_run<2>(pat, m, txt, n, hbuf, lane);
    if (orall < 8) return myers_run<3>(pat, m, txt, n, hbuf, lane);
    return myers_run<8>(pat, m, txt, n, hbuf, lane);
}

struct GenomeView { const uint8_t* bytes; const int64_t* off; int32_t n; const int32_t* rank_to_tid; int32_t n_ranks; };

// pairs between cluster-stage INS signatures (positions in the key-sorted array)
__global__ void __launch_bounds__(128) k_myers_pairs(const svim_csig* sig, const uint8_t* ins_blob, GenomeView g, const MyersWork* work,
                                                      uint32_t n_work, int32_t* ed_out, uint8_t* scratch, int64_t maxlen, uint32_t* next,
                                                      unsigned long long* cells, uint32_t* err) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t* my = scratch + (size_t)warp * 3 * maxlen;
    unsigned long long my_cells = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(next, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_work) break;
        const MyersWork wk = work[w];
        const svim_csig a = sig[wk.a], b = sig[wk.b];
        // contig of both signatures is the same (same partition)
        const int32_t rank = a.contig_a;
        int32_t tid = (g.rank_to_tid && rank >= 0 && rank < g.n_ranks) ? g.rank_to_tid[rank] : -1;
        if (tid < 0 || tid >= g.n) { if (lane == 0) { atomicExch(err, 1u); ed_out[wk.slot] = 0; } continue; }
        const uint8_t* contig = g.bytes + g.off[tid];
        const int64_t clen = g.off[tid + 1] - g.off[tid];
        const int64_t s1 = (int64_t)a.start, s2 = (int64_t)b.start;
        const int64_t ws = (s1 < s2 ? s1 : s2) - 100, we = (s1 > s2 ? s1 : s2) + 100;
        HapSource ha = make_hap(contig, clen, ws < 0 ? 0 : ws, we < 0 ? 0 : we, s1 < 0 ? 0 : s1, ins_blob + a.seq_off, a.seq_len);
        HapSource hb = make_hap(contig, clen, ws < 0 ? 0 : ws, we < 0 ? 0 : we, s2 < 0 ? 0 : s2, ins_blob + b.seq_off, b.seq_len);
        const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
        if (la > maxlen || lb > maxlen) { if (lane == 0) { atomicExch(err, 2u); ed_out[wk.slot] = 0; } continue; }
        int32_t ed = myers_pair(ha, hb, my, maxlen, lane);
        if (lane == 0) { ed_out[wk.slot] = ed; my_cells += (unsigned long long)la * (unsigned long long)lb; }
    }
    if (lane == 0 && my_cells) atomicAdd(cells, my_cells);
}

// unit-test entry: explicit string pairs
__global__ void __launch_bounds__(128) k_myers_strings(const uint8_t* blob, const int64_t* a_off, const int32_t* a_len, const int64_t* b_off,
                                                        const int32_t* b_len, uint32_t n_pairs, int32_t* out, uint8_t* scratch, int64_t maxlen,
                                                        uint32_t* next) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t* my = scratch + (size_t)warp * 3 * maxlen;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(next, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_pairs) break;
        HapSource ha{blob, 0, blob + a_off[w], a_len[w], blob, 0};
        HapSource hb{blob, 0, blob + b_off[w], b_len[w], blob, 0};
        int32_t ed = myers_pair(ha, hb, my, maxlen, lane);
        if (lane == 0) out[w] = ed;
    }
}

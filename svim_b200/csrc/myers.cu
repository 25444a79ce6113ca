// K2b: haplotype edit distance (compute_haplotype_edit_distance, SVIM_clustering.py:32-45;
// edlib.align NW distance) as a bit-parallel Myers/Hyyro kernel.
//
// * Exact work reduction first: both haplotypes of a pair carry the same 100 bp genome flanks, and a common prefix /
//   suffix never changes a Levenshtein distance, so the window is cut to [min(start), max(start)] (pair_haps).
// * The longer haplotype is the pattern: rows are cut into 64-bit words held in registers as 32-bit halves.
//   A group of G lanes owns one pair with WPL consecutive words per lane; pairs are binned by pattern length into
//   ten (G, WPL) shapes (4x1 ... 32x4) so short patterns share a warp and the per-step overhead is amortised.
//   Bin 9 (32 lanes x 4 words) strip-mines patterns longer than 8 192 rows, parking the horizontal deltas of a
//   strip's bottom row in a per-warp column buffer.
// * Lanes run as a systolic wavefront: at step s lane l handles text column s-l.  The text is pre-expanded to one
//   uint32 of byte masks per column; lane 0 injects it together with the incoming horizontal delta, every lane
//   forwards the same word with its own bottom delta in the top byte: one shuffle per step, no per-lane loads.
// * Symbol equality comes from 3 bit-planes of a compact bijective symbol code (A,C,G,T,N + 3); a pair holding any
//   other byte is deferred to an 8-plane kernel, so arbitrary bytes stay exact.
// * The distance is read from the vertical deltas of the last column: D[m][n] = n + sum_rows(Pv - Mv).
//
// Integer-ALU bound: ~34-40 ALU-pipe instructions per 64-cell word step (ncu: ALU pipe 90-97 % busy, FMA pipe idle);
// DRAM traffic is the two haplotypes per pair.  No tensor cores: there is no dense contraction here.
#pragma once
#include "ctx.cuh"
#include "myers_band.cuh"   // bins, word steps, banded wavefront (host-checkable)
#include "myers_tpp.cuh"    // thread-per-pair banded window (host-checkable)

struct MyersWork { uint32_t a, b, slot, pad; };

struct HapSource {
    const uint8_t* p1; int64_t l1;   // reference left of the insertion point
    const uint8_t* p2; int64_t l2;   // inserted sequence
    const uint8_t* p3; int64_t l3;   // reference right of it
};

struct GenomeView { const uint8_t* bytes; const int64_t* off; int32_t n; const int32_t* rank_to_tid; int32_t n_ranks; };

// Symbol codes of the thread-per-pair kernels: ((c & 0xDF) >> 1 & 3) << 5 for A/C/G/T in either case (the value Eq is indexed
// with), bit 7 set for any other byte.  The genome is encoded once per upload, the INS blob once per cluster call; a pair's two
// haplotypes are then assembled from code bytes by plain copies (k_myers_tpp), instead of classifying every base of every pair again.
__global__ void k_tpp_encode(const uint8_t* __restrict__ src, int64_t n, uint8_t* __restrict__ dst) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    if (i0 + 16 <= n && (((uintptr_t)(src + i0) | (uintptr_t)(dst + i0)) & 15) == 0) {
        const uint4 v = *(const uint4*)(src + i0);
        uint32_t w[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t r = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t up = (w[q] >> (8 * b)) & 0xDFu;
                const uint32_t ok = (up == 'A') | (up == 'C') | (up == 'G') | (up == 'T');
                r |= ((((up >> 1) & 3u) << 5) | (ok ? 0u : 0x80u)) << (8 * b);
            }
            o[q] = r;
        }
        *(uint4*)(dst + i0) = make_uint4(o[0], o[1], o[2], o[3]);
        return;
    }
    for (int64_t i = i0; i < n && i < i0 + 16; ++i) {
        const uint32_t up = src[i] & 0xDFu;
        const uint32_t ok = (up == 'A') | (up == 'C') | (up == 'G') | (up == 'T');
        dst[i] = (uint8_t)((((up >> 1) & 3u) << 5) | (ok ? 0u : 0x80u));
    }
}

__constant__ uint8_t c_symcode[256];

static void myers_init_symcode() {
    uint8_t t[256];
    for (int i = 0; i < 256; ++i) t[i] = (uint8_t)i;
    const uint8_t acgtn[5] = {'A', 'C', 'G', 'T', 'N'};
    for (int k = 0; k < 5; ++k) { uint8_t x = t[k]; t[k] = t[acgtn[k]]; t[acgtn[k]] = x; }   // swaps keep it a bijection
    uint8_t lut[256];
    for (int i = 0; i < 256; ++i) { int c = (i >= 'a' && i <= 'z') ? i - 32 : i; lut[i] = t[c]; }   // .upper()
    cudaMemcpyToSymbol(c_symcode, lut, 256);
}

__host__ __device__ __forceinline__ int64_t clampi(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

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

// both haplotypes of an insertion pair; false if the contig is not in the genome table
__device__ __forceinline__ bool pair_haps(const svim_csig& a, const svim_csig& b, const uint8_t* ins_blob, const GenomeView& g,
                                          HapSource& ha, HapSource& hb) {
    const int32_t rank = a.contig_a;
    const int32_t tid = (g.rank_to_tid && rank >= 0 && rank < g.n_ranks) ? g.rank_to_tid[rank] : -1;
    if (tid < 0 || tid >= g.n) return false;
    const uint8_t* contig = g.bytes + g.off[tid];
    const int64_t clen = g.off[tid + 1] - g.off[tid];
    const int64_t s1 = (int64_t)a.start, s2 = (int64_t)b.start;
    // The reference pads both haplotypes with the same 100 bp of genome on either side (window_start/window_end,
    // SVIM_clustering.py:33-34).  A common prefix and a common suffix never change a Levenshtein distance, so the
    // window is cut down to [min(start), max(start)] — exact, and (200+L)^2 -> L^2 cells for co-located insertions.
    const int64_t lo = s1 < s2 ? s1 : s2, hi = s1 > s2 ? s1 : s2;
    ha = make_hap(contig, clen, lo < 0 ? 0 : lo, hi < 0 ? 0 : hi, s1 < 0 ? 0 : s1, ins_blob + a.seq_off, a.seq_len);
    hb = make_hap(contig, clen, lo < 0 ? 0 : lo, hi < 0 ? 0 : hi, s2 < 0 ? 0 : s2, ins_blob + b.seq_off, b.seq_len);
    return true;
}

// materialise symbol codes with the G lanes of a group; returns OR of all codes (group-uniform)
template <int G>
__device__ __forceinline__ uint32_t hap_write_codes(const HapSource& h, uint8_t* dst, int gl, bool valid) {
    uint32_t orall = 0;
    if (valid) {
        const int64_t n = h.l1 + h.l2 + h.l3;
        for (int64_t k = gl; k < n; k += G) {
            uint8_t c = k < h.l1 ? h.p1[k] : (k < h.l1 + h.l2 ? h.p2[k - h.l1] : h.p3[k - h.l1 - h.l2]);
            uint8_t code = c_symcode[c];
            dst[k] = code; orall |= code;
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) orall |= __shfl_xor_sync(0xffffffffu, orall, o, G);
    return orall;
}

// one Myers word step (Hyyro): updates Pv/Mv, returns hout; ph/mh are the pre-shift horizontal vectors
__device__ __forceinline__ int myers_word(uint64_t Eq, uint64_t& Pv, uint64_t& Mv, int hin, uint64_t& ph_out, uint64_t& mh_out) {
    const uint64_t pv = Pv, mv = Mv;
    const uint64_t hneg = hin < 0 ? 1ull : 0ull, hpos = hin > 0 ? 1ull : 0ull;
    const uint64_t Xv = Eq | mv;
    Eq |= hneg;
    const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
    uint64_t Ph = mv | ~(Xh | pv);
    uint64_t Mh = pv & Xh;
    ph_out = Ph; mh_out = Mh;
    const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
    Ph = (Ph << 1) | hpos; Mh = (Mh << 1) | hneg;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return hout;
}

template <int NP>
__device__ __forceinline__ void build_planes(const uint8_t* __restrict__ pat, int64_t m, int64_t row0, uint64_t (&pl)[NP], uint64_t& vm) {
    vm = 0;
#pragma unroll
    for (int b = 0; b < NP; ++b) pl[b] = 0;
    if (row0 < m) {
        const int cnt = (m - row0) < 64 ? (int)(m - row0) : 64;
        for (int r = 0; r < cnt; ++r) {
            const uint64_t code = pat[row0 + r];
#pragma unroll
            for (int b = 0; b < NP; ++b) pl[b] |= ((code >> b) & 1ull) << r;
        }
        vm = cnt == 64 ? ~0ull : ((1ull << cnt) - 1ull);
    }
}

// ---- generic kernel (NP bit-planes, any bytes): a whole warp per pair, WPL words per lane, strip-mined -----------
template <int WPL, int NP>
__device__ int32_t myers_run(const uint8_t* __restrict__ pat, int64_t m, const uint8_t* __restrict__ txt, int64_t n,
                             int8_t* __restrict__ hbuf, int lane) {
    const int64_t W = (m + 63) >> 6;
    int64_t score = 0;
    const int64_t STRIP = 32 * WPL;
    for (int64_t sb = 0; sb < W; sb += STRIP) {
        const bool first_strip = (sb == 0), last_strip = (sb + STRIP >= W);
        const int64_t ws_cnt = (W - sb) < STRIP ? (W - sb) : STRIP;
        const int nl = (int)((ws_cnt + WPL - 1) / WPL);   // active lanes
        uint64_t pl[WPL][NP], vm[WPL], Pv[WPL], Mv[WPL];
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            build_planes<NP>(pat, m, (sb + (int64_t)lane * WPL + k) * 64, pl[k], vm[k]);
            Pv[k] = ~0ull; Mv[k] = 0;
        }
        const int64_t wl = W - 1 - sb;   // where the pattern's last row lives (last strip only)
        const int l_last = (int)(wl / WPL), k_last = (int)(wl % WPL), bit_last = (int)((m - 1) & 63);
        int carry = 0;
        const int64_t steps = n + nl - 1;
        uint8_t c_next = (lane == 0 && n > 0) ? txt[0] : 0;
        int8_t h_next = (!first_strip && lane == 0 && n > 0) ? hbuf[0] : 0;
        for (int64_t s = 0; s < steps; ++s) {
            const int recv = __shfl_up_sync(0xffffffffu, carry, 1);
            const int64_t j = s - lane;
            const bool act = (lane < nl) && j >= 0 && j < n;
            const uint8_t c = c_next; const int8_t hb = h_next;
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
                for (int k = 0; k < WPL; ++k) {
                    uint64_t Eq = vm[k];
#pragma unroll
                    for (int b = 0; b < NP; ++b) Eq &= ~(pl[k][b] ^ mk[b]);
                    uint64_t ph, mh;
                    const int hout = myers_word(Eq, Pv[k], Mv[k], hin, ph, mh);
                    if (last_strip && lane == l_last && k == k_last) score += (int64_t)((ph >> bit_last) & 1ull) - (int64_t)((mh >> bit_last) & 1ull);
                    hin = hout;
                }
                carry = hin;
                if (!last_strip && lane == 31) hbuf[j] = (int8_t)hin;
            }
        }
        __syncwarp();
    }
    const int64_t wl = (W - 1) % STRIP;
    score = __shfl_sync(0xffffffffu, score, (int)(wl / WPL));
    return (int32_t)(m + score);
}

// ---- fast path (3 bit-planes, 32-bit halves, text masks travel with the carry) -------------------------
// Text is pre-expanded to one uint32 per column: byte b = 0xFF if bit b of the symbol code is set.  Lane 0
// injects it together with the incoming horizontal delta (+1 at the top boundary, or the previous strip's
// bottom delta), every lane forwards the same word with its own bottom delta in the top byte, so each step
// costs one shuffle and no per-lane text load.  The distance is read at the end from the vertical deltas of
// the last column:  D[m][n] = n + sum_rows (Pv - Mv).
template <int G>
__device__ __forceinline__ uint32_t hap_write_txt32(const HapSource& h, uint32_t* dst, int gl, bool valid) {
    uint32_t orall = 0;
    if (valid) {
        const int64_t n = h.l1 + h.l2 + h.l3;
        for (int64_t k = gl; k < n; k += G) {
            const uint8_t c = k < h.l1 ? h.p1[k] : (k < h.l1 + h.l2 ? h.p2[k - h.l1] : h.p3[k - h.l1 - h.l2]);
            const uint32_t code = c_symcode[c];
            dst[k] = ((code & 1u) ? 0xFFu : 0u) | ((code & 2u) ? 0xFF00u : 0u) | ((code & 4u) ? 0xFF0000u : 0u);
            orall |= code;
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) orall |= __shfl_xor_sync(0xffffffffu, orall, o, G);
    return orall;
}

__device__ __forceinline__ void word_init(Word32& w, const uint8_t* __restrict__ pat, int64_t m, int64_t row0) {
    uint64_t pl[3], vm;
    build_planes<3>(pat, m, row0, pl, vm);
    w.p0l = (uint32_t)pl[0]; w.p0h = (uint32_t)(pl[0] >> 32); w.p1l = (uint32_t)pl[1]; w.p1h = (uint32_t)(pl[1] >> 32);
    w.p2l = (uint32_t)pl[2]; w.p2h = (uint32_t)(pl[2] >> 32);
    w.pvl = w.pvh = 0xffffffffu; w.mvl = w.mvh = 0u;
}

// vertical-delta sum of the rows of this word that belong to the pattern
__device__ __forceinline__ int word_vsum(const Word32& w, int64_t m, int64_t row0) {
    if (row0 >= m) return 0;
    const int cnt = (m - row0) < 64 ? (int)(m - row0) : 64;
    const uint64_t vm = cnt == 64 ? ~0ull : ((1ull << cnt) - 1ull);
    const uint64_t pv = ((uint64_t)w.pvh << 32) | w.pvl, mv = ((uint64_t)w.mvh << 32) | w.mvl;
    return __popcll(pv & vm) - __popcll(mv & vm);
}

// ---- two-plane (A/C/G/T) words on the FMA pipe (WordQ / block_step, myers_band.cuh) ----
__device__ __forceinline__ void wordq_init(WordQ& w, const uint8_t* __restrict__ pat, int64_t m, int64_t row0) {
    uint64_t pl[2], vm;
    build_planes<2>(pat, m, row0, pl, vm);
    wordq_from_planes(w, pl[0], pl[1], vm, ~0ull);
}

__device__ __forceinline__ int wordq_vsum(const WordQ& w, int64_t m, int64_t row0) {
    if (row0 >= m) return 0;
    const int cnt = (m - row0) < 64 ? (int)(m - row0) : 64;
    const uint64_t vm = cnt == 64 ? ~0ull : ((1ull << cnt) - 1ull);
    const uint64_t pv = ((uint64_t)w.pv[1] << 32) | w.pv[0], mv = ((uint64_t)w.mv[1] << 32) | w.mv[0];
    return __popcll(pv & vm) - __popcll(mv & vm);
}

// G lanes per pair, WPL words per lane; strips of G*WPL words (only G = 32 ever needs more than one strip)
template <int G, int WPL, int NP>
__device__ int32_t myers_fast(const uint8_t* __restrict__ pat, int64_t m, const uint32_t* __restrict__ txt32, int n,
                              uint8_t* __restrict__ hbuf, int gl, bool valid) {
    const int64_t W = valid ? ((m + 63) >> 6) : 0;
    const int64_t STRIP = G * WPL;
    int vsum = 0;
    int64_t n_strips = (W + STRIP - 1) / STRIP;
    if (G < 32) n_strips = 1;            // bins guarantee W <= G*WPL; keeps the loop warp-uniform across groups
    for (int64_t si = 0; si < n_strips; ++si) {
        const int64_t sb = si * STRIP;
        const bool first_strip = (si == 0), last_strip = (si + 1 >= n_strips);
        const int64_t ws_cnt = (W - sb) < STRIP ? (W - sb) : STRIP;
        const int nl = valid ? (int)((ws_cnt + WPL - 1) / WPL) : 0;
        Word32 w[WPL];
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            if (valid) word_init(w[k], pat, m, (sb + (int64_t)gl * WPL + k) * 64);
            else { w[k].p0l = w[k].p0h = w[k].p1l = w[k].p1h = w[k].p2l = w[k].p2h = 0; w[k].pvl = w[k].pvh = w[k].mvl = w[k].mvh = 0; }
        }
        int steps = valid ? (n + nl - 1) : 0;
        if (G < 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, steps, o); steps = t > steps ? t : steps; }
        }
        const bool lane_on = valid && gl < nl, head = lane_on && gl == 0;
        uint32_t pk_out = 0;
        uint32_t t_nxt = head ? txt32[0] : 0u;
        uint32_t h_nxt = (head && !first_strip) ? hbuf[0] : 2u;
        for (int s = 0; s < steps; ++s) {
            const uint32_t recv = __shfl_up_sync(0xffffffffu, pk_out, 1, G);
            const uint32_t t_cur = t_nxt, h_cur = h_nxt;
            if (head && s + 1 < n) { t_nxt = txt32[s + 1]; if (G == 32 && !first_strip) h_nxt = hbuf[s + 1]; }
            const uint32_t pk = gl == 0 ? (t_cur | (h_cur << 24)) : recv;
            const int j = s - gl;
            if (lane_on && (unsigned)j < (unsigned)n) {
                const uint32_t m0 = __byte_perm(pk, 0, 0x0000), m1 = __byte_perm(pk, 0, 0x1111);
                const uint32_t m2 = NP == 2 ? 0u : __byte_perm(pk, 0, 0x2222);
                uint32_t e = pk >> 24;
#pragma unroll
                for (int k = 0; k < WPL; ++k) e = word_step<NP>(w[k], m0, m1, m2, e);
                pk_out = (pk & 0x00ffffffu) | (e << 24);
                if (G == 32 && !last_strip && gl == 31) hbuf[j] = (uint8_t)e;
            }
        }
#pragma unroll
        for (int k = 0; k < WPL; ++k) vsum += lane_on ? word_vsum(w[k], m, (sb + (int64_t)gl * WPL + k) * 64) : 0;
        __syncwarp();
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o, G);
    return n + vsum;
}

// two-plane variant of myers_fast on WordQ blocks (same wavefront, same packet format)
template <int G, int WPL, bool HI>
__device__ int32_t myers_fast_q(const uint8_t* __restrict__ pat, int64_t m, const uint32_t* __restrict__ txt32, int n,
                                uint8_t* __restrict__ hbuf, int gl, bool valid, uint32_t one, uint32_t two) {
    const int64_t W = valid ? ((m + 63) >> 6) : 0;
    const int64_t STRIP = G * WPL;
    int vsum = 0;
    int64_t n_strips = (W + STRIP - 1) / STRIP;
    if (G < 32) n_strips = 1;
    for (int64_t si = 0; si < n_strips; ++si) {
        const int64_t sb = si * STRIP;
        const bool first_strip = (si == 0), last_strip = (si + 1 >= n_strips);
        const int64_t ws_cnt = (W - sb) < STRIP ? (W - sb) : STRIP;
        const int nl = valid ? (int)((ws_cnt + WPL - 1) / WPL) : 0;
        WordQ w[WPL];
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            if (valid) wordq_init(w[k], pat, m, (sb + (int64_t)gl * WPL + k) * 64);
            else {
#pragma unroll
                for (int h = 0; h < 2; ++h) { w[k].q0[h] = w[k].d1[h] = w[k].d2[h] = w[k].d3[h] = 0; w[k].pv[h] = w[k].mv[h] = 0; }
            }
        }
        int steps = valid ? (n + nl - 1) : 0;
        if (G < 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, steps, o); steps = t > steps ? t : steps; }
        }
        const bool lane_on = valid && gl < nl, head = lane_on && gl == 0;
        uint32_t pk_out = 0;
        uint32_t t_nxt = head ? txt32[0] : 0u;
        uint32_t h_nxt = (head && !first_strip) ? hbuf[0] : 2u;
        for (int s = 0; s < steps; ++s) {
            const uint32_t recv = __shfl_up_sync(0xffffffffu, pk_out, 1, G);
            const uint32_t t_cur = t_nxt, h_cur = h_nxt;
            if (head && s + 1 < n) { t_nxt = txt32[s + 1]; if (G == 32 && !first_strip) h_nxt = hbuf[s + 1]; }
            const uint32_t pk = gl == 0 ? (t_cur | (h_cur << 24)) : recv;
            const int j = s - gl;
            if (lane_on && (unsigned)j < (unsigned)n) {
                // packet: byte 0 / byte 1 = 0xFF when bit 0 / bit 1 of the symbol code is set, byte 3 = hin + 1
                const uint32_t b0 = pk & 1u, b1 = (pk >> 8) & 1u;
                const uint32_t c1 = b0, c2 = b1, c3 = b0 & b1;      // bilinear Eq (wordq_from_planes)
                const uint32_t e = pk >> 24;
                uint32_t hp = e >> 1, hn = 1u >> e;
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    block_step<HI>(w[k].q0[0], w[k].d1[0], w[k].d2[0], w[k].d3[0], w[k].pv[0], w[k].mv[0], c1, c2, c3, one, two, hp, hn);
                    block_step<HI>(w[k].q0[1], w[k].d1[1], w[k].d2[1], w[k].d3[1], w[k].pv[1], w[k].mv[1], c1, c2, c3, one, two, hp, hn);
                }
                const uint32_t eo = 1u + hp - hn;
                pk_out = (pk & 0x00ffffffu) | (eo << 24);
                if (G == 32 && !last_strip && gl == 31) hbuf[j] = (uint8_t)eo;
            }
        }
#pragma unroll
        for (int k = 0; k < WPL; ++k) vsum += lane_on ? wordq_vsum(w[k], m, (sb + (int64_t)gl * WPL + k) * 64) : 0;
        __syncwarp();
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o, G);
    return n + vsum;
}

// ---- kernels -------------------------------------------------------------------------------------
struct MyersArgs {
    const svim_csig* sig; const uint8_t* ins_blob; GenomeView g;
    const MyersWork* work; uint32_t n_work;
    int32_t* ed_out;
    uint8_t* scratch; int64_t maxlen;         // per group: 6*maxlen bytes (pattern codes, 4-byte text masks, column buffer)
    uint32_t* next;                           // work cursor
    MyersWork* fallback; uint32_t* n_fallback;   // pairs that need the 8-plane kernel
    unsigned long long* cells; uint32_t* err;
    uint32_t one, two;                        // 1 and 2, opaque to the compiler: multiplicands that keep adds/shifts on the FMA pipe
    // banded first pass (k_myers_band): pairs whose result exceeds the bound go to `retry`, one region per unbanded bin;
    // the unbanded kernel of that bin reads them as its second list (`extra`, count on the device)
    const MyersWork* extra; const uint32_t* n_extra;
    MyersWork* retry; uint32_t* n_retry; uint32_t retry_off[MYERS_BINS];
    int32_t band_num, band_add;
    unsigned long long* band_cells;           // cells the banded pass actually computed (statistics)
    unsigned long long* tpp_cells;            // cells the thread-per-pair kernels computed (columns x 32 x window blocks)
    unsigned long long* unb_cells;            // cells the unbanded kernels computed (full matrices: own pairs + hand-overs)
    // second wave (pairs the first pass handed over): MyersWork.pad == 1 marks a pair holding symbols outside A/C/G/T.
    uint32_t tpp_second;                      // k_myers_tpp: take the `extra` list, run every unmarked pair unbanded
    unsigned long long* trace;                // debug (SVIM_MYERS_TRACE=1): [2k] first start / [2k+1] last end of launch k, %globaltimer ns
    uint32_t trace_slot;
    const uint8_t* genome_codes; const uint8_t* ins_codes; const uint8_t* ins_base; const uint8_t* str_codes;   // k_tpp_encode images of g.bytes / the INS blob / the string blob
    uint32_t extra_marked_only;               // k_myers_fast: of the `extra` list take only the marked pairs
};

// explicit string pairs (unit-test entry svimgpu_edit_distance) share the kernels below through this view
struct StringPairs { const uint8_t* blob; const int64_t* a_off; const int32_t* a_len; const int64_t* b_off; const int32_t* b_len; const uint32_t* list; };

// work item w of a kernel's lists -> the two haplotypes; false (and an error flag) if the contig is unknown
template <bool STRINGS>
__device__ __forceinline__ bool myers_load_pair(const MyersArgs& a, const StringPairs& sp, uint32_t w, MyersWork& wk, HapSource& ha, HapSource& hb, int gl) {
    if (STRINGS) {
        const uint32_t i = w < a.n_work ? sp.list[w] : a.extra[w - a.n_work].slot;
        wk.slot = i; wk.a = i; wk.b = 0; wk.pad = 0;
        ha = HapSource{sp.blob, 0, sp.blob + sp.a_off[i], sp.a_len[i], sp.blob, 0};
        hb = HapSource{sp.blob, 0, sp.blob + sp.b_off[i], sp.b_len[i], sp.blob, 0};
        return true;
    }
    wk = w < a.n_work ? a.work[w] : a.extra[w - a.n_work];
    if (!pair_haps(a.sig[wk.a], a.sig[wk.b], a.ins_blob, a.g, ha, hb)) { if (gl == 0) { atomicExch(a.err, 1u); a.ed_out[wk.slot] = 0; } return false; }
    return true;
}

// MODE: 0 = ALU-pipe formulation, 1 = two-plane pairs on the FMA-pipe formulation, 2 = same with IMAD.HI for the top bits
template <int G, int WPL, bool STRINGS, int MODE>
__global__ void __launch_bounds__(128) k_myers_fast(MyersArgs a, StringPairs sp) {
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), grp = lane / G;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t* my = a.scratch + ((size_t)warp * GPW + grp) * 6 * a.maxlen;
    if (a.trace && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(a.trace + 2 * a.trace_slot, t); }
    unsigned long long my_cells = 0;
    const uint32_t n_total = a.n_work + (a.n_extra ? *a.n_extra : 0u);
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.next, (uint32_t)GPW);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_total) break;
        const uint32_t w = base + grp;
        bool valid = w < n_total;
        MyersWork wk{0, 0, 0, 0};
        HapSource ha{nullptr, 0, nullptr, 0, nullptr, 0}, hb = ha;
        if (valid) valid = myers_load_pair<STRINGS>(a, sp, w, wk, ha, hb, gl);
        if (valid && a.extra_marked_only && w >= a.n_work && wk.pad == 0) valid = false;      // the thread-per-pair second wave takes it
        const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
        if (valid && (la > a.maxlen || lb > a.maxlen)) { if (gl == 0) { atomicExch(a.err, 2u); a.ed_out[wk.slot] = 0; } valid = false; }
        if (valid && (la == 0 || lb == 0)) { if (gl == 0) a.ed_out[wk.slot] = (int32_t)(la + lb); valid = false; }
        const bool a_is_pat = la >= lb;
        const HapSource& hp = a_is_pat ? ha : hb;
        const HapSource& ht = a_is_pat ? hb : ha;
        const int64_t m = a_is_pat ? la : lb, n = a_is_pat ? lb : la;
        uint8_t* pat = my; uint32_t* txt32 = (uint32_t*)(my + a.maxlen); uint8_t* hbuf = my + 5 * a.maxlen;
        const uint32_t orall = hap_write_codes<G>(hp, pat, gl, valid) | hap_write_txt32<G>(ht, txt32, gl, valid);
        __syncwarp();
        if (valid && orall >= 8) {   // symbols outside the 3-plane code space: defer to the 8-plane kernel
            if (gl == 0) { uint32_t f = atomicAdd(a.n_fallback, 1u); a.fallback[f] = wk; }
            valid = false;
        }
        // pure A/C/G/T pairs (codes 0-3) need two planes only; the choice is warp-uniform
        const bool three = __any_sync(0xffffffffu, valid && orall >= 4);
        int32_t ed;
        if (three) ed = myers_fast<G, WPL, 3>(pat, m, txt32, (int)n, hbuf, gl, valid);
        else if (MODE == 0) ed = myers_fast<G, WPL, 2>(pat, m, txt32, (int)n, hbuf, gl, valid);
        else ed = myers_fast_q<G, WPL, MODE == 2>(pat, m, txt32, (int)n, hbuf, gl, valid, a.one, a.two);
        if (valid && gl == 0) { a.ed_out[wk.slot] = ed; my_cells += (unsigned long long)la * (unsigned long long)lb; }
        __syncwarp();
    }
    if (a.trace && lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(a.trace + 2 * a.trace_slot + 1, t); }
    if (my_cells) { atomicAdd(a.cells, my_cells); if (a.unb_cells) atomicAdd(a.unb_cells, my_cells); }
}

// Banded first pass (myers_band.cuh): G lanes rotate over the word groups the band crosses, so the shape follows the
// band width.  Pure A/C/G/T pairs only (two planes, FMA-pipe words); anything else, and every pair whose banded
// result exceeds its bound, is handed to the unbanded kernel of the pattern's bin.
__host__ __device__ __forceinline__ size_t myers_band_scratch(int64_t maxlen, int WPL) {
    return (size_t)((2 * maxlen + 24 * ((maxlen >> 6) + WPL + 1) + 15) & ~15ll);   // pattern codes, text codes, 3 plane words per pattern word
}

// the wavefront is a long dependent chain per warp: resident warps matter, so the register budget is pinned per WPL
template <int G, int WPL, bool STRINGS, bool HI>
__global__ void __launch_bounds__(128, WPL <= 1 ? 8 : WPL == 2 ? 6 : WPL == 3 ? 5 : 4) k_myers_band(MyersArgs a, StringPairs sp) {
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), grp = lane / G;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t* my = a.scratch + ((size_t)warp * GPW + grp) * myers_band_scratch(a.maxlen, WPL);
    uint8_t* pat = my; uint8_t* txt = my + a.maxlen; uint64_t* planes = (uint64_t*)(my + 2 * a.maxlen);
    unsigned long long my_cells = 0, my_band = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.next, (uint32_t)GPW);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= a.n_work) break;
        const uint32_t w = base + grp;
        bool valid = w < a.n_work;
        MyersWork wk{0, 0, 0, 0};
        HapSource ha{nullptr, 0, nullptr, 0, nullptr, 0}, hb = ha;
        if (valid) valid = myers_load_pair<STRINGS>(a, sp, w, wk, ha, hb, gl);
        const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
        if (valid && (la > a.maxlen || lb > a.maxlen)) { if (gl == 0) { atomicExch(a.err, 2u); a.ed_out[wk.slot] = 0; } valid = false; }
        if (valid && (la == 0 || lb == 0)) { if (gl == 0) a.ed_out[wk.slot] = (int32_t)(la + lb); valid = false; }
        const bool a_is_pat = la >= lb;
        const HapSource& hp = a_is_pat ? ha : hb;
        const HapSource& ht = a_is_pat ? hb : ha;
        const int64_t m = a_is_pat ? la : lb, n = a_is_pat ? lb : la;
        const uint32_t orall = hap_write_codes<G>(hp, pat, gl, valid) | hap_write_codes<G>(ht, txt, gl, valid);
        const int64_t k = myers_band_k(m, n, a.band_num, a.band_add);
        const int rbin = myers_bin_of(m);
        if (valid && (orall >= 4 || k < 0 || !myers_band_fits(m, n, k, G, WPL))) {   // not a banded pair after all
            if (gl == 0) { wk.pad = orall >= 4 ? 1u : 0u; const uint32_t f = atomicAdd(a.n_retry + rbin, 1u); a.retry[a.retry_off[rbin] + f] = wk; }
            valid = false;
        }
        BandGeom ge = band_geom(valid ? m : 0, valid ? n : 0, valid ? k : 0, WPL);
        __syncwarp();
        for (int64_t x = gl; x < (int64_t)ge.NG * WPL; x += G) band_build_word(pat, m, ge.pad, x, planes + 3 * x);
        __syncwarp();
        BandLane<WPL> L;
        band_lane_init(L, ge, planes, txt, gl, valid);
        int steps = valid ? ge.n + ge.NG - 1 : 0;
        if (G < 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, steps, o); steps = t > steps ? t : steps; }
        }
        uint32_t e_out = 0;
        for (int s = 0; s < steps; ++s) {
            const uint32_t recv = __shfl_sync(0xffffffffu, e_out, (gl - 1) & (G - 1), G);
            e_out = band_lane_step<G, WPL, HI>(L, ge, planes, s, recv, e_out, a.one, a.two);
        }
        int score = L.score;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) score += __shfl_xor_sync(0xffffffffu, score, o, G);
        if (valid && gl == 0) {
            const int64_t ed = m + score;
            my_band += (unsigned long long)ge.n * (unsigned long long)(ge.a + ge.b + 64 * WPL);
            if (ed <= k) { a.ed_out[wk.slot] = (int32_t)ed; my_cells += (unsigned long long)la * (unsigned long long)lb; }
            else { const uint32_t f = atomicAdd(a.n_retry + rbin, 1u); a.retry[a.retry_off[rbin] + f] = wk; }
        }
        __syncwarp();
    }
    if (my_cells) atomicAdd(a.cells, my_cells);
    if (my_band && a.band_cells) atomicAdd(a.band_cells, my_band);
}

// ---- thread-per-pair banded window (myers_tpp.cuh) -----------------------------------------------------------------------
// A warp takes 32 pairs at a time.  Preparation is cooperative (coalesced): for each of the 32 pairs all lanes turn the
// pattern into symbol codes and then into per-block match masks (one lane per 32-row block), and the text into pre-scaled
// symbol bytes stored at column index j + phase; then every lane runs tpp_thread on its own pair.  Pairs holding a symbol
// outside A/C/G/T, or not fitting this bucket after all, are handed to the unbanded kernel of their pattern's bin, and so
// is every banded pair whose result exceeds its bound.
template <int B>
struct TppEq {
    uint32_t* base;                                   // shared memory of this warp, + lane; layout [block][symbol][lane]
    __device__ __forceinline__ void put(int slot, const uint32_t v[4]) {
#pragma unroll
        for (int c = 0; c < 4; ++c) base[(slot * 4 + c) * 32] = v[c];
    }
    // sym = code * 32.  The load is a volatile asm so that it stays where tpp_thread puts it — TPP_AHEAD block steps before its use;
    // left to the compiler it sinks next to the use and puts ~30 cycles of shared-memory latency on every step of the carry chain.
    __device__ __forceinline__ uint32_t get(int slot, uint32_t sym) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + (sym + slot * 128) * 4u));
        return v;
    }
    uint32_t sbase;                                   // shared-window address of `base`
    __device__ __forceinline__ void shift_up() {
#pragma unroll
        for (int i = 0; i + 1 < B; ++i) {
#pragma unroll
            for (int c = 0; c < 4; ++c) base[(i * 4 + c) * 32] = base[((i + 1) * 4 + c) * 32];
        }
    }
};
struct TppPeq {
    const uint4* blocks; int32_t nblk;
    __device__ __forceinline__ void block(int b, uint32_t v[4]) const {
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (b >= 0 && b < nblk) x = blocks[b];
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    }
};
struct TppTxt {
    const uint8_t* t;                                 // 4-byte aligned, index = column + phase
    __device__ __forceinline__ uint32_t byte(int s) const { return t[s]; }
    __device__ __forceinline__ uint32_t word(int s) const { return *(const uint32_t*)(t + s); }
};

__host__ __device__ __forceinline__ size_t tpp_pair_scratch(int64_t maxlen) {   // masks, pattern codes, text symbols of one pair
    return (size_t)(16 * ((maxlen + 31) / 32 + 1) + ((maxlen + 32 + 15) & ~15ll) + ((maxlen + 64 + 15) & ~15ll));
}

__device__ __forceinline__ HapSource hap_bcast(const HapSource& h, int src) {
    HapSource r;
    r.p1 = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)h.p1, src); r.l1 = (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)h.l1, src);
    r.p2 = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)h.p2, src); r.l2 = (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)h.l2, src);
    r.p3 = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)h.p3, src); r.l3 = (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)h.l3, src);
    return r;
}

// all lanes: dst[lead, lead + len) = the three pieces of a haplotype, read from their code images (k_tpp_encode); dst[0, lead) = 0.
// Returns (per lane) the OR of what was copied — bit 7 set means a symbol outside A/C/G/T.  `dst` is 4-byte aligned; every lane
// assembles whole words.
__device__ __forceinline__ uint32_t tpp_copy_codes(const uint8_t* c1, int32_t l1, const uint8_t* c2, int32_t l2, const uint8_t* c3, int32_t l3,
                                                   uint8_t* dst, int32_t lead, int lane) {
    uint32_t seen = 0;
    const int32_t len = l1 + l2 + l3, l12 = l1 + l2;
    for (int32_t k0 = 4 * lane; k0 < lead + len; k0 += 128) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int32_t k = k0 + b - lead;
            uint32_t c = 0;
            if (k >= 0 && k < len) c = k < l1 ? c1[k] : (k < l12 ? c2[k - l1] : c3[k - l12]);
            w |= c << (8 * b);
        }
        seen |= w;
        *(uint32_t*)(dst + k0) = w;
    }
    return seen;
}

// code image of a haplotype piece: same offset as the raw bytes, in the image of the array the piece lives in
template <bool STRINGS>
__device__ __forceinline__ const uint8_t* tpp_code_ptr(const MyersArgs& a, const uint8_t* raw, int is_ins) {
    if (STRINGS) return a.str_codes + (raw - a.ins_base);           // explicit string pairs: everything lives in one blob
    return is_ins ? a.ins_codes + (raw - a.ins_base) : a.genome_codes + (raw - a.g.bytes);
}

// one lane: match masks of pattern block `blk` from its 32 symbol codes (code << 5 per byte, 16-byte aligned)
__device__ __forceinline__ uint4 tpp_block_masks(const uint8_t* codes, int32_t m, int32_t blk) {
    const uint4 lo = *(const uint4*)(codes + 32 * blk), hi = *(const uint4*)(codes + 32 * blk + 16);
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t p0 = 0, p1 = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {   // bit 0 / bit 1 of the four code bytes of w[q] -> four adjacent bits (multiply-gather)
        p0 |= ((((w[q] >> 5) & 0x01010101u) * 0x00204081u >> 21) & 0xFu) << (4 * q);
        p1 |= ((((w[q] >> 6) & 0x01010101u) * 0x00204081u >> 21) & 0xFu) << (4 * q);
    }
    const int32_t cnt = m - 32 * blk;
    const uint32_t real = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
    return make_uint4(~p0 & ~p1 & real, p0 & ~p1 & real, ~p0 & p1 & real, p0 & p1 & real);
}

#ifndef TPP_MINB
#define TPP_MINB(B) ((B) <= 8 ? 8 : (B) <= 12 ? 6 : (B) <= 20 ? 5 : 4)      // CTAs per SM the register budget is cut for
#endif
template <int B, bool STRINGS>
__global__ void __launch_bounds__(128, TPP_MINB(B)) k_myers_tpp(MyersArgs a, StringPairs sp) {
    extern __shared__ __align__(16) uint32_t tpp_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t per_pair = tpp_pair_scratch(a.maxlen);
    uint8_t* wscr = a.scratch + (size_t)warp * 32 * per_pair;
    const size_t off_codes = (size_t)16 * ((a.maxlen + 31) / 32 + 1), off_txt = off_codes + (size_t)((a.maxlen + 32 + 15) & ~15ll);
    TppEq<B> eq; eq.base = tpp_smem + wib * (B * 128) + lane; eq.sbase = (uint32_t)__cvta_generic_to_shared(eq.base);
    unsigned long long t_begin = 0;
    if (a.trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin)); if (threadIdx.x == 0) atomicMin(a.trace + 2 * a.trace_slot, t_begin); }
    unsigned long long my_cells = 0, my_comp = 0;
    const uint32_t n_total = a.n_work + (a.n_extra ? *a.n_extra : 0u);
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.next, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_total) break;
        const uint32_t w = base + lane;
        bool valid = w < n_total;
        MyersWork wk{0, 0, 0, 0};
        HapSource ha{nullptr, 0, nullptr, 0, nullptr, 0}, hb = ha;
        if (valid) valid = myers_load_pair<STRINGS>(a, sp, w, wk, ha, hb, 0);
        if (valid && a.tpp_second && wk.pad) valid = false;            // marked pairs run on the plane kernels
        const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
        if (valid && (la > a.maxlen || lb > a.maxlen)) { atomicExch(a.err, 2u); a.ed_out[wk.slot] = 0; valid = false; }
        if (valid && (la == 0 || lb == 0)) { a.ed_out[wk.slot] = (int32_t)(la + lb); valid = false; }
        const bool a_is_pat = la >= lb;
        const HapSource& hp = a_is_pat ? ha : hb;
        const HapSource& ht = a_is_pat ? hb : ha;
        const int32_t m = (int32_t)(a_is_pat ? la : lb), n = (int32_t)(a_is_pat ? lb : la);
        TppPlan plan = tpp_plan(valid ? m : 1, valid ? n : 1, a.band_num, a.band_add);
        if (a.tpp_second) { plan.a = -1; plan.B = (int32_t)tpp_blocks_full(valid ? m : 1); }      // second wave: the whole pattern
        const int64_t k = myers_band_k(m, n, a.band_num, a.band_add);
        const int rbin = myers_bin_of(m);
        bool hand_over = valid && plan.B > B;            // not a pair of this bucket after all
        const int32_t phase = plan.a >= 0 ? ((-plan.a) & 31) : 0;
        // ---- cooperative preparation of the 32 pairs --------------------------------------------------------
        const uint32_t todo = __ballot_sync(0xffffffffu, valid && !hand_over);
        for (uint32_t rest = todo; rest; rest &= rest - 1) {
            const int p = __ffs(rest) - 1;
            const HapSource pp = hap_bcast(hp, p), pt = hap_bcast(ht, p);
            const int32_t pm = __shfl_sync(0xffffffffu, m, p), pph = __shfl_sync(0xffffffffu, phase, p);
            uint8_t* slot = wscr + (size_t)p * per_pair;
            // text symbols start at column index `phase`: the word-aligned copy begins at the aligned offset below it
            uint32_t bad = tpp_copy_codes(tpp_code_ptr<STRINGS>(a, pp.p1, 0), (int32_t)pp.l1, tpp_code_ptr<STRINGS>(a, pp.p2, 1), (int32_t)pp.l2,
                                          tpp_code_ptr<STRINGS>(a, pp.p3, 0), (int32_t)pp.l3, slot + off_codes, 0, lane);
            bad |= tpp_copy_codes(tpp_code_ptr<STRINGS>(a, pt.p1, 0), (int32_t)pt.l1, tpp_code_ptr<STRINGS>(a, pt.p2, 1), (int32_t)pt.l2,
                                  tpp_code_ptr<STRINGS>(a, pt.p3, 0), (int32_t)pt.l3, slot + off_txt + (pph & ~3), pph & 3, lane);
            bad &= 0x80808080u;
            __syncwarp();
            const int32_t nblk = (pm + 31) >> 5;
            for (int32_t blk = lane; blk < nblk; blk += 32) ((uint4*)slot)[blk] = tpp_block_masks(slot + off_codes, pm, blk);
            if (__any_sync(0xffffffffu, bad) && lane == p) hand_over = true;
        }
        __syncwarp();
        if (hand_over) {
            if (a.tpp_second) { const uint32_t f = atomicAdd(a.n_fallback, 1u); a.fallback[f] = wk; }       // cannot happen for a well-formed plan: exact any-byte kernel
            else { wk.pad = plan.B > B ? 0u : 1u; const uint32_t f = atomicAdd(a.n_retry + rbin, 1u); a.retry[a.retry_off[rbin] + f] = wk; }
            valid = false;
        }
        // ---- one pair per lane ------------------------------------------------------------------------------------
        if (valid) {
            const uint8_t* slot = wscr + (size_t)lane * per_pair;
            TppPeq peq{(const uint4*)slot, (m + 31) >> 5};
            TppTxt txt{slot + off_txt};
            const int32_t ed = tpp_thread<B>(m, n, plan.a, eq, peq, txt, a.one, a.two);
            my_comp += (unsigned long long)n * (unsigned long long)(32 * B);
            if (plan.a < 0 || ed <= k) { a.ed_out[wk.slot] = ed; my_cells += (unsigned long long)la * (unsigned long long)lb; }
            else { const uint32_t f = atomicAdd(a.n_retry + rbin, 1u); a.retry[a.retry_off[rbin] + f] = wk; }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o); my_comp += __shfl_xor_sync(0xffffffffu, my_comp, o); }
    if (lane == 0 && my_cells) atomicAdd(a.cells, my_cells);
    if (lane == 0 && my_comp && a.tpp_cells) atomicAdd(a.tpp_cells, my_comp);
    if (a.trace && lane == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(a.trace + 2 * a.trace_slot + 1, t);
        atomicAdd(a.trace + 128 + 2 * a.trace_slot, t - t_begin); atomicAdd(a.trace + 128 + 2 * a.trace_slot + 1, 1ull);      // warp residency: sum, count
    }
}

// any bytes: 8 bit-planes, a warp per pair, 4 words per lane, strip-mined
template <bool STRINGS>
__global__ void __launch_bounds__(128) k_myers_generic(MyersArgs a, StringPairs sp) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t* my = a.scratch + (size_t)warp * 6 * a.maxlen;
    unsigned long long my_cells = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(a.next, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= a.n_work) break;
        const MyersWork wk = a.work[w];
        HapSource ha, hb;
        if (STRINGS) {
            const uint32_t i = wk.slot;
            ha = HapSource{sp.blob, 0, sp.blob + sp.a_off[i], sp.a_len[i], sp.blob, 0};
            hb = HapSource{sp.blob, 0, sp.blob + sp.b_off[i], sp.b_len[i], sp.blob, 0};
        } else if (!pair_haps(a.sig[wk.a], a.sig[wk.b], a.ins_blob, a.g, ha, hb)) { if (lane == 0) { atomicExch(a.err, 1u); a.ed_out[wk.slot] = 0; } continue; }
        const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
        if (la > a.maxlen || lb > a.maxlen) { if (lane == 0) { atomicExch(a.err, 2u); a.ed_out[wk.slot] = 0; } continue; }
        if (la == 0 || lb == 0) { if (lane == 0) a.ed_out[wk.slot] = (int32_t)(la + lb); continue; }
        const bool a_is_pat = la >= lb;
        const HapSource& hp = a_is_pat ? ha : hb;
        const HapSource& ht = a_is_pat ? hb : ha;
        const int64_t m = a_is_pat ? la : lb, n = a_is_pat ? lb : la;
        uint8_t* pat = my; uint8_t* txt = my + a.maxlen; int8_t* hbuf = (int8_t*)(my + 2 * a.maxlen);
        hap_write_codes<32>(hp, pat, lane, true); hap_write_codes<32>(ht, txt, lane, true);
        __syncwarp();
        const int32_t ed = myers_run<4, 8>(pat, m, txt, n, hbuf, lane);
        if (lane == 0) { a.ed_out[wk.slot] = ed; my_cells += (unsigned long long)la * (unsigned long long)lb; }
        __syncwarp();
    }
    if (lane == 0 && my_cells) { atomicAdd(a.cells, my_cells); if (a.unb_cells) atomicAdd(a.unb_cells, my_cells); }
}

// the Myers bins run on SVIM_AUX_STREAMS side streams: fork from / join to the context's main stream
static cudaError_t myers_fork(svimgpu_ctx* ctx) {
    cudaError_t e = cudaEventRecord(ctx->aux_ev[SVIM_AUX_STREAMS], ctx->stream);
    for (int i = 0; i < SVIM_AUX_STREAMS && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(ctx->aux_stream[i], ctx->aux_ev[SVIM_AUX_STREAMS], 0);
    return e;
}
static cudaError_t myers_join(svimgpu_ctx* ctx) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < SVIM_AUX_STREAMS && e == cudaSuccess; ++i) {
        e = cudaEventRecord(ctx->aux_ev[i], ctx->aux_stream[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[i], 0);
    }
    return e;
}

// extra_cap = upper bound of the second list (pairs the banded pass may hand over), for grid sizing only
template <bool STRINGS>
static cudaError_t myers_launch_bin(svimgpu_ctx* ctx, int bin, MyersArgs a, StringPairs sp, DevBuf& scratch, int sms, uint32_t extra_cap = 0) {
    const uint64_t n_items = (uint64_t)a.n_work + extra_cap;
    if (n_items == 0) return cudaSuccess;
    if (extra_cap == 0) { a.extra = nullptr; a.n_extra = nullptr; }
    const MyersBin spec = bin < MYERS_BINS ? myers_bin_spec(bin) : MyersBin{32, 4, 1 << 30};
    const int64_t cap = (bin < MYERS_BINS - 1) ? (int64_t)64 * spec.capW : a.maxlen;   // longest haplotype in the bin
    a.maxlen = (cap + 15) & ~15ll;
    const int groups = 32 / spec.G;
    int blocks = sms * 8;
    blocks = (int)std::min<int64_t>(blocks, ((int64_t)n_items + 4 * groups - 1) / (4 * groups));
    const size_t per_warp = (size_t)6 * groups * a.maxlen;
    while (blocks > sms && (size_t)blocks * 4 * per_warp > ((size_t)16 << 30)) blocks -= sms;
    cudaError_t e = scratch.ensure((size_t)blocks * 4 * per_warp);
    if (e != cudaSuccess) return e;
    a.scratch = scratch.as<uint8_t>();
    ctx->launches++;
    cudaStream_t stream = ctx->aux_stream[bin % SVIM_AUX_STREAMS];   // bins overlap: one kernel's tail is filled by the next
    a.one = 1u; a.two = 2u;
    a.trace = ctx->myers_trace ? ctx->d_myers_trace.as<unsigned long long>() : nullptr; a.trace_slot = 20u + (uint32_t)bin;
#define MYERS_LAUNCH(G_, W_)                                                                                     \
    switch (ctx->myers_mode) {                                                                                   \
        case 0: k_myers_fast<G_, W_, STRINGS, 0><<<blocks, 128, 0, stream>>>(a, sp); break;                      \
        case 2: k_myers_fast<G_, W_, STRINGS, 2><<<blocks, 128, 0, stream>>>(a, sp); break;                      \
        default: k_myers_fast<G_, W_, STRINGS, 1><<<blocks, 128, 0, stream>>>(a, sp); break;                     \
    }                                                                                                            \
    break;
    switch (bin) {
        case 0: MYERS_LAUNCH(4, 1)
        case 1: MYERS_LAUNCH(4, 2)
        case 2: MYERS_LAUNCH(4, 3)
        case 3: MYERS_LAUNCH(4, 4)
        case 4: MYERS_LAUNCH(8, 3)
        case 5: MYERS_LAUNCH(8, 4)
        case 6: MYERS_LAUNCH(16, 3)
        case 7: MYERS_LAUNCH(16, 4)
        case 8: MYERS_LAUNCH(32, 3)
        case 9: MYERS_LAUNCH(32, 4)
        default: k_myers_generic<STRINGS><<<blocks, 128, 0, stream>>>(a, sp); break;   // MYERS_BINS: any bytes
    }
#undef MYERS_LAUNCH
    return cudaGetLastError();
}

// banded first pass of shape `bin`; a.maxlen = longest haplotype of the whole run (the shape bounds the band, not the pattern)
template <bool STRINGS>
static cudaError_t myers_launch_band(svimgpu_ctx* ctx, int bin, MyersArgs a, StringPairs sp, DevBuf& scratch, int sms) {
    if (a.n_work == 0) return cudaSuccess;
    const MyersBin spec = myers_bin_spec(bin);
    a.maxlen = (a.maxlen + 15) & ~15ll;
    const int groups = 32 / spec.G;
    int blocks = sms * 8;
    blocks = (int)std::min<int64_t>(blocks, ((int64_t)a.n_work + 4 * groups - 1) / (4 * groups));
    const size_t per_warp = (size_t)groups * myers_band_scratch(a.maxlen, spec.WPL);
    while (blocks > sms && (size_t)blocks * 4 * per_warp > ((size_t)16 << 30)) blocks -= sms;
    cudaError_t e = scratch.ensure((size_t)blocks * 4 * per_warp);
    if (e != cudaSuccess) return e;
    a.scratch = scratch.as<uint8_t>();
    ctx->launches++;
    cudaStream_t stream = ctx->aux_stream[bin % SVIM_AUX_STREAMS];
    a.one = 1u; a.two = 2u;
#define MYERS_LAUNCH(G_, W_)                                                                                     \
    if (ctx->myers_mode == 2) k_myers_band<G_, W_, STRINGS, true><<<blocks, 128, 0, stream>>>(a, sp);           \
    else k_myers_band<G_, W_, STRINGS, false><<<blocks, 128, 0, stream>>>(a, sp);                               \
    break;
    switch (bin) {
        case 0: MYERS_LAUNCH(4, 1)
        case 1: MYERS_LAUNCH(4, 2)
        case 2: MYERS_LAUNCH(4, 3)
        case 3: MYERS_LAUNCH(4, 4)
        case 4: MYERS_LAUNCH(8, 3)
        case 5: MYERS_LAUNCH(8, 4)
        case 6: MYERS_LAUNCH(16, 3)
        case 7: MYERS_LAUNCH(16, 4)
        case 8: MYERS_LAUNCH(32, 3)
        default: MYERS_LAUNCH(32, 4)
    }
#undef MYERS_LAUNCH
    return cudaGetLastError();
}

// longest pattern a pair of TPP bucket q can have under the band policy (bounds the per-pair scratch)
static int64_t tpp_bucket_maxlen(int q, int64_t maxlen, int32_t num, int32_t add) {
    const int64_t B = tpp_bucket_B(q);
    int64_t cap = 32 * B;                                             // unbanded pairs of the bucket
    if (num > 0) {
        const int64_t K = 32 * B - 31;                                // banded: a + b <= 32B - 32, a + b in {k-1, k}
        if (K >= add) cap = std::max<int64_t>(cap, ((K - add + 1) * 1024 + num - 1) / num + 1);
    }
    return std::min<int64_t>(maxlen, cap + 64);
}

// second_bin >= 0: second wave — the pairs the first pass handed over to unbanded bin `second_bin` (patterns up to 64 * capW
// rows, i.e. 8 * (second_bin + 1) blocks), whole pattern in the window; extra_cap bounds their number for the grid size
template <bool STRINGS>
static cudaError_t myers_launch_tpp(svimgpu_ctx* ctx, int q, MyersArgs a, StringPairs sp, DevBuf& scratch, int sms, int second_bin = -1, uint32_t extra_cap = 0) {
    const uint64_t n_items = second_bin >= 0 ? (uint64_t)extra_cap : (uint64_t)a.n_work;
    if (n_items == 0) return cudaSuccess;
    const int B = tpp_bucket_B(q);
    if (second_bin >= 0) { a.maxlen = std::min<int64_t>(a.maxlen, 32 * B); a.tpp_second = 1u; a.n_work = 0; a.work = nullptr; sp.list = nullptr; }
    else { a.maxlen = tpp_bucket_maxlen(q, a.maxlen, a.band_num, a.band_add); a.tpp_second = 0u; a.extra = nullptr; a.n_extra = nullptr; }
    a.maxlen = (a.maxlen + 15) & ~15ll;
    const int occ = B <= 8 ? 8 : B <= 12 ? 6 : B <= 20 ? 5 : 4;
    int blocks = sms * occ;
    blocks = (int)std::min<int64_t>(blocks, ((int64_t)n_items + 127) / 128);
    const size_t per_warp = (size_t)32 * tpp_pair_scratch(a.maxlen);
    while (blocks > sms && (size_t)blocks * 4 * per_warp > ((size_t)8 << 30)) blocks -= sms;
    cudaError_t e = scratch.ensure((size_t)blocks * 4 * per_warp);
    if (e != cudaSuccess) return e;
    a.scratch = scratch.as<uint8_t>();
    ctx->launches++;
    cudaStream_t stream = ctx->aux_stream[q % SVIM_AUX_STREAMS];
    a.one = 1u; a.two = 2u;
    a.trace = ctx->myers_trace ? ctx->d_myers_trace.as<unsigned long long>() : nullptr; a.trace_slot = second_bin >= 0 ? 40u + second_bin : (uint32_t)q;
    const size_t smem = (size_t)4 * B * 128 * 4;
#define TPP_LAUNCH(B_)                                                                                                     \
    case B_:                                                                                                              \
        e = cudaFuncSetAttribute(k_myers_tpp<B_, STRINGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        if (e == cudaSuccess) k_myers_tpp<B_, STRINGS><<<blocks, 128, smem, stream>>>(a, sp);                              \
        break;
    switch (B) {
        TPP_LAUNCH(2) TPP_LAUNCH(3) TPP_LAUNCH(4) TPP_LAUNCH(5) TPP_LAUNCH(6) TPP_LAUNCH(7) TPP_LAUNCH(8) TPP_LAUNCH(9) TPP_LAUNCH(10)
        TPP_LAUNCH(12) TPP_LAUNCH(14) TPP_LAUNCH(16) TPP_LAUNCH(18) TPP_LAUNCH(20) TPP_LAUNCH(22) TPP_LAUNCH(24) TPP_LAUNCH(26) TPP_LAUNCH(28)
        default: return cudaErrorInvalidValue;
    }
#undef TPP_LAUNCH
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---- the whole edit-distance stage: banded shapes first, then the unbanded bins (their own pairs + the banded pass's
// hand-overs), then the 8-plane kernel for pairs with symbols outside the code space -------------------------------------
#define MYERS_LISTS (2 * MYERS_BINS + TPP_BUCKETS)
struct MyersPlan {
    uint32_t cnt[MYERS_LISTS];        // first-pass items: [0,10) unbanded bins, [10,20) banded wavefront shapes, [20,38) thread-per-pair buckets
    uint32_t off[MYERS_LISTS];        // their offsets in the work list (pipeline) / index list (string pairs)
    uint32_t retry_cap[MYERS_BINS];   // banded / thread-per-pair items whose pattern belongs to each unbanded bin
};
// control words on the device: [0, MYERS_LISTS) list counts of k_ins_pairs, [MYERS_CTL_CAP, +10) hand-over capacities,
// [MYERS_CTL_CURSOR, +MYERS_LISTS+1) work cursors, [MYERS_CTL_RETRY, +10) hand-over counts, then 64-bit statistics
enum { MYERS_CTL_CAP = 40, MYERS_CTL_CURSOR = 64, MYERS_CTL_RETRY = 112, MYERS_CTL_BANDCELLS = 128, MYERS_CTL_TPPCELLS = 130, MYERS_CTL_UNBCELLS = 132, MYERS_CTL_FALLBACK = 136,
       MYERS_CTL_N = 160 };

// ctl: MYERS_CTL_N device words; retry: room for sum(retry_cap) items; ma carries sig/ins/genome, ed_out, fallback list,
// cells / band_cells / err pointers and the band policy.  Synchronises ctx->stream.
template <bool STRINGS>
static cudaError_t myers_run_plan(svimgpu_ctx* ctx, const MyersPlan& pl, MyersArgs ma, StringPairs sp, const MyersWork* work, const uint32_t* list,
                                  MyersWork* retry, uint32_t* ctl, int64_t maxlen, int sms) {
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    cudaStream_t st = ctx->stream;
    if (ctx->myers_trace) {
        chk(ctx->d_myers_trace.ensure(64 * 32));
        unsigned long long init[256]; for (int k = 0; k < 64; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; init[128 + 2 * k] = 0ull; init[128 + 2 * k + 1] = 0ull; }
        chk(cudaMemcpyAsync(ctx->d_myers_trace.p, init, sizeof(init), cudaMemcpyHostToDevice, st)); chk(cudaStreamSynchronize(st));
    }
    chk(cudaMemsetAsync(ctl + MYERS_CTL_CURSOR, 0, (MYERS_CTL_N - MYERS_CTL_CURSOR) * 4, st));
    ma.n_fallback = ctl + MYERS_CTL_FALLBACK; ma.n_retry = ctl + MYERS_CTL_RETRY; ma.retry = retry;
    uint32_t acc = 0, n_banded = 0;
    for (int bb = 0; bb < MYERS_BINS; ++bb) { ma.retry_off[bb] = acc; acc += pl.retry_cap[bb]; n_banded += pl.cnt[MYERS_BINS + bb]; }
    for (int q = 0; q < TPP_BUCKETS; ++q) n_banded += pl.cnt[2 * MYERS_BINS + q];
    ma.band_cells = (unsigned long long*)(ctl + MYERS_CTL_BANDCELLS); ma.tpp_cells = (unsigned long long*)(ctl + MYERS_CTL_TPPCELLS); ma.unb_cells = (unsigned long long*)(ctl + MYERS_CTL_UNBCELLS);
    if (n_banded > 0) {
        chk(myers_fork(ctx));
        for (int q = TPP_BUCKETS - 1; q >= 0 && e == cudaSuccess; --q) {      // widest windows first: their batches run longest
            const int li = 2 * MYERS_BINS + q;
            ma.work = work ? work + pl.off[li] : nullptr; sp.list = list ? list + pl.off[li] : nullptr;
            ma.n_work = pl.cnt[li]; ma.next = ctl + MYERS_CTL_CURSOR + li; ma.maxlen = maxlen; ma.extra = nullptr; ma.n_extra = nullptr;
            chk(myers_launch_tpp<STRINGS>(ctx, q, ma, sp, ctx->d_myers_scratch[24 + q], sms));
        }
        for (int bb = MYERS_BINS - 1; bb >= 0 && e == cudaSuccess; --bb) {
            const int q = MYERS_BINS + bb;
            ma.work = work ? work + pl.off[q] : nullptr; sp.list = list ? list + pl.off[q] : nullptr;
            ma.n_work = pl.cnt[q]; ma.next = ctl + MYERS_CTL_CURSOR + q; ma.maxlen = maxlen; ma.extra = nullptr; ma.n_extra = nullptr;
            chk(myers_launch_band<STRINGS>(ctx, bb, ma, sp, ctx->d_myers_scratch[12 + bb], sms));
        }
        chk(myers_join(ctx));
    }
    // longest bins first so the tail of the launch sequence is made of short pairs; bins overlap on side streams
    // Second wave.  Hand-overs with patterns up to 768 rows run unbanded on the thread-per-pair kernels of 8 / 16 / 24 blocks (their
    // unbanded bins 0-2 then only take the pairs marked as holding symbols outside A/C/G/T); longer ones on the wavefront kernels.
    chk(myers_fork(ctx));
    for (int bb = MYERS_BINS - 1; bb >= 0 && e == cudaSuccess; --bb) {
        const bool second_tpp = ctx->myers_tpp && bb <= 2 && pl.retry_cap[bb] > 0;
        ma.work = work ? work + pl.off[bb] : nullptr; sp.list = list ? list + pl.off[bb] : nullptr;
        ma.n_work = pl.cnt[bb]; ma.next = ctl + MYERS_CTL_CURSOR + bb; ma.maxlen = maxlen;
        ma.extra = retry + ma.retry_off[bb]; ma.n_extra = ctl + MYERS_CTL_RETRY + bb;
        ma.extra_marked_only = second_tpp ? 1u : 0u;
        chk(myers_launch_bin<STRINGS>(ctx, bb, ma, sp, ctx->d_myers_scratch[bb], sms, pl.retry_cap[bb]));
        if (second_tpp) {
            ma.next = ctl + MYERS_CTL_CURSOR + MYERS_LISTS + 1 + bb; ma.maxlen = maxlen; ma.extra_marked_only = 0u;
            chk(myers_launch_tpp<STRINGS>(ctx, tpp_bucket_of(8 * (bb + 1)), ma, sp, ctx->d_myers_scratch[42 + bb], sms, bb, pl.retry_cap[bb]));
        }
    }
    ma.extra_marked_only = 0u; ma.tpp_second = 0u;
    chk(myers_join(ctx));
    uint32_t n_fb = 0;
    chk(cudaMemcpyAsync(&n_fb, ctl + MYERS_CTL_FALLBACK, 4, cudaMemcpyDeviceToHost, st));
    chk(cudaStreamSynchronize(st));
    if (ctx->myers_trace && e == cudaSuccess) {
        unsigned long long tr[256];
        if (cudaMemcpy(tr, ctx->d_myers_trace.p, sizeof(tr), cudaMemcpyDeviceToHost) == cudaSuccess) {
            unsigned long long t0 = ~0ull; for (int k = 0; k < 64; ++k) if (tr[2 * k] < t0) t0 = tr[2 * k];
            for (int k = 0; k < 64; ++k) if (tr[2 * k + 1]) fprintf(stderr, "[myers trace] %s %2d  items %8u  start %8.3f ms  end %8.3f ms  warps %6llu  mean warp residency %8.3f ms\n", k < 20 ? "tpp  B" : k < 40 ? "fast bin" : "tpp2 bin",
                                                              k < 20 ? tpp_bucket_B(k) : k < 40 ? k - 20 : k - 40, k < 20 ? pl.cnt[2 * MYERS_BINS + k] : k < 40 ? pl.cnt[k - 20] : pl.retry_cap[k - 40], (tr[2 * k] - t0) * 1e-6, (tr[2 * k + 1] - t0) * 1e-6,
                                                              tr[128 + 2 * k + 1], tr[128 + 2 * k + 1] ? tr[128 + 2 * k] * 1e-6 / tr[128 + 2 * k + 1] : 0.0);
        }
    }
    if (e == cudaSuccess && n_fb > 0) {   // pairs with symbols outside A,C,G,T,N(+3): exact 8-plane kernel
        sp.list = nullptr;
        ma.work = ma.fallback; ma.n_work = n_fb; ma.next = ctl + MYERS_CTL_CURSOR + MYERS_LISTS; ma.maxlen = maxlen; ma.fallback = nullptr; ma.n_fallback = nullptr;
        ma.extra = nullptr; ma.n_extra = nullptr;
        chk(myers_fork(ctx));
        chk(myers_launch_bin<STRINGS>(ctx, MYERS_BINS, ma, sp, ctx->d_myers_scratch[MYERS_BINS], sms));
        chk(myers_join(ctx));
    }
    return e;
}

// COLLECT kernels: analyze_alignment_file_coordsorted (SVIM_COLLECT.py:132-167).
//
//   k_cigar_scan      K1  one warp per alignment record; streams the BAM-encoded CIGAR
//                         with 128-bit loads (analyze_cigar_indel, SVIM_intra.py:8-30 and
//                         analyze_alignment_indel :33-51), emits DEL/INS records, and
//                         leaves a ChainWork item for primaries with an SA tag.
//   k_segment_chain   K1b one thread per ChainWork item: SA parsing + split-read tree.
//   k_ins_gather      K1c one warp per INS record: 4-bit SEQ -> ASCII blob.
//
// HBM-bound by design: the scan touches every CIGAR word exactly once
// (B_aln = 44 + 4*n_cigar + sa_len bytes per record) and nothing else.
#pragma once
#include <cub/cub.cuh>
#include <thread>
#include "ctx.cuh"

#define FULL 0xffffffffu

struct DeviceEmitter {
    SigQueue qm, qt;
    uint32_t* overflow;
    uint32_t slot_m = 0xffffffffu, slot_t = 0xffffffffu;   // next pre-reserved slot of this lane (scan kernel), else one atomic per record
    __device__ __forceinline__ void push(SigQueue& q, const svim_sig& s, uint32_t& reserved) {
        const uint32_t slot = reserved != 0xffffffffu ? reserved++ : atomicAdd(q.count, 1u);
        if (slot < q.cap) q.recs[slot] = s; else atomicExch(overflow, 1u);
    }
    __device__ __forceinline__ void sig(const svim_sig& s) { push(qm, s, slot_m); }
    __device__ __forceinline__ void twin(const svim_sig& s) { push(qt, s, slot_t); }
};

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}

// class bit of a CIGAR word: 1 << op
__device__ __forceinline__ uint32_t op_bit(uint32_t v) { return 1u << (v & 15u); }

// SUM: also accumulate N / H bases (only primaries with an SA tag need reference_end and the hard-clip test)
template <bool SUM>
__device__ __forceinline__ void acc_op(uint32_t v, uint32_t B, uint32_t& ref, uint32_t& read, uint32_t& nsum, uint32_t& hsum) {
    const uint32_t len = v >> 4;
    if (B & SVIM_MASK_REF_QUIRK) ref += len;
    if (B & SVIM_MASK_READ) read += len;
    if (SUM) {
        if (B & (1u << OP_N)) nsum += len;
        if (B & (1u << OP_H)) hsum += len;
    }
}
// SV-sized insertion or deletion: op in {I, D} and len >= min_sv_size (thresh = min_sv_size << 4)
__device__ __forceinline__ bool is_event(uint32_t v, uint32_t B, uint32_t thresh) { return (B & 0x6u) && v >= thresh; }

#define SCAN_UNROLL 4
#define SCAN_BATCH 4     // alignments fetched per atomic

// Queue slots are handed to a warp in chunks: one global atomic per SCAN_CHUNK signatures instead of one per signature
// (event-dense CIGARs serialise on that counter otherwise).  Slots a warp reserved but never filled are marked as holes
// and sort to the end of the queue.
#define SCAN_CHUNK 16
struct SlotChunk { uint32_t base, left; };

__device__ __forceinline__ void mark_hole(svim_sig& r) { r.aln_idx = 0xffffffffu; r.ordinal = 0xffffffffu; r.type = 0xff; }

__device__ __noinline__ uint32_t reserve_slots(const SigQueue& q, SlotChunk* c, uint32_t need, int lane, uint32_t* holes) {
    uint32_t base = c->base; const uint32_t left = c->left;
    __syncwarp();
    if (left >= need) { if (lane == 0) { c->base = base + need; c->left = left - need; } __syncwarp(); return base; }
    for (uint32_t k = lane; k < left; k += 32) if (base + k < q.cap) mark_hole(q.recs[base + k]);     // abandon the remainder
    const uint32_t take = need > SCAN_CHUNK ? need : SCAN_CHUNK;
    uint32_t nb = 0;
    if (lane == 0) { nb = atomicAdd(q.count, take); if (left) atomicAdd(holes, left); }
    nb = __shfl_sync(FULL, nb, 0);
    if (lane == 0) { c->base = nb + need; c->left = take - need; }
    __syncwarp();
    return nb;
}

__device__ __forceinline__ void flush_slots(const SigQueue& q, SlotChunk* c, int lane, uint32_t* holes) {
    const uint32_t base = c->base, left = c->left;
    for (uint32_t k = lane; k < left; k += 32) if (base + k < q.cap) mark_hole(q.recs[base + k]);
    if (lane == 0 && left) atomicAdd(holes, left);
}

struct ScanRec { uint32_t i, qid; int32_t tid; int64_t ref_start, l_seq; };
struct EvState { int64_t base_ref, base_read; uint32_t n_ev, n_tw, nsum, hsum; };

// one 128-op group (a uint4 per lane) that holds at least one SV-sized I/D: exact positions via warp scans.
// acc_ref/acc_read: lane-private consumption since the last fold; returned state has them folded in.
template <bool SUM>
__device__ __noinline__ EvState scan_events(const uint4 w, uint32_t thresh, int lane, const ChainParams& p, const ScanRec& r, EvState st,
                                            uint32_t acc_ref, uint32_t acc_read, DeviceEmitter out) {
    const uint32_t v[4] = {w.x, w.y, w.z, w.w};
    st.base_ref += warp_sum(acc_ref); st.base_read += warp_sum(acc_read);
    uint32_t g_ref = 0, g_read = 0, g_n = 0, g_h = 0;
    uint32_t pre_ref[4], pre_read[4]; uint32_t my_ev = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t B = op_bit(v[k]);
        pre_ref[k] = g_ref; pre_read[k] = g_read;
        acc_op<SUM>(v[k], B, g_ref, g_read, g_n, g_h);
        my_ev += is_event(v[k], B, thresh);
    }
    st.nsum += g_n; st.hsum += g_h;
    uint32_t tot_ref, tot_read, tot_ev;
    const uint32_t ex_ref = warp_excl_scan(g_ref, lane, tot_ref);
    const uint32_t ex_read = warp_excl_scan(g_read, lane, tot_read);
    const uint32_t ex_ev = warp_excl_scan(my_ev, lane, tot_ev);
    uint32_t ord = st.n_ev + ex_ev;
    uint32_t tw_before = 0;
    if (p.all_bnds) {   // twins only for deletions: count DEL events before this lane
        uint32_t my_del = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) my_del += (is_event(v[k], op_bit(v[k]), thresh) && (v[k] & 15u) == OP_D);
        uint32_t tot_del; const uint32_t ex_del = warp_excl_scan(my_del, lane, tot_del);
        tw_before = st.n_tw + ex_del; st.n_tw += tot_del;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!is_event(v[k], op_bit(v[k]), thresh)) continue;
        const uint32_t op = v[k] & 15u; const int64_t len = v[k] >> 4;
        const int64_t pr = st.base_ref + ex_ref + pre_ref[k];
        const int64_t pq = st.base_read + ex_read + pre_read[k];
        svim_sig s; memset(&s, 0, sizeof(s));
        s.contig1 = r.tid; s.contig2 = -1; s.start = (int32_t)(r.ref_start + pr); s.end = (int32_t)(r.ref_start + pr + len);
        s.aln_idx = r.i; s.qname_id = r.qid; s.ordinal = ord++;
        if (op == OP_D) {
            s.type = SVIM_DEL;
            out.sig(s);
            if (p.all_bnds) {   // SVIM_intra.py:43-44 (same contig, start < end: already canonical)
                svim_sig b = s; b.type = SVIM_BND; b.contig2 = r.tid; b.pos = s.end; b.end = s.start + 1; b.ordinal = tw_before++;
                if (len == 0) { b.flags = SVIM_F_DIR1_REV | SVIM_F_DIR2_REV; }   // pos1 == pos2: the else-branch flips both directions
                out.twin(b);
            }
        } else {
            s.type = SVIM_INS;
            int64_t lo, hi; py_slice(pq, len, r.l_seq, lo, hi);   // query_sequence[pos_read:pos_read+len]
            s.seq_off = (uint64_t)lo; s.seq_len = (uint32_t)(hi - lo);
            out.sig(s);
        }
    }
    st.n_ev += tot_ev;
    st.base_ref += tot_ref; st.base_read += tot_read;
    return st;
}

// one uint4 (4 ops) of the lane: fast path accumulates into registers, rare path goes through scan_events
#define SCAN_GROUP(W)                                                                                              \
    {                                                                                                              \
        const uint32_t B0 = op_bit((W).x), B1 = op_bit((W).y), B2 = op_bit((W).z), B3 = op_bit((W).w);             \
        const bool ev = is_event((W).x, B0, thresh) | is_event((W).y, B1, thresh) | is_event((W).z, B2, thresh) |  \
                        is_event((W).w, B3, thresh);                                                               \
        if (__ballot_sync(FULL, ev) == 0) {                                                                        \
            acc_op<SUM>((W).x, B0, a_ref, a_read, a_n, a_h); acc_op<SUM>((W).y, B1, a_ref, a_read, a_n, a_h);      \
            acc_op<SUM>((W).z, B2, a_ref, a_read, a_n, a_h); acc_op<SUM>((W).w, B3, a_ref, a_read, a_n, a_h);      \
        } else {                                                                                                   \
            st = scan_events<SUM>((W), thresh, lane, p, r, st, a_ref, a_read, out);                                \
            a_ref = 0; a_read = 0;                                                                                 \
        }                                                                                                          \
    }

// stream one record's CIGAR.  Main loop: whole 512-op blocks without any bounds test.
template <bool SUM, int UNROLL>
__device__ __forceinline__ void scan_cigar(const uint4* __restrict__ cg, uint32_t n, uint32_t thresh, int lane, const ChainParams& p, const ScanRec& r,
                                           EvState& st_out, uint32_t& acc_ref_out, uint32_t& acc_read_out, const DeviceEmitter& out) {
    const uint32_t n4 = (n + 3) >> 2;
    const uint32_t full = (n >> 2) / (32 * UNROLL) * (32 * UNROLL);   // uint4 groups in complete blocks of 128*UNROLL ops
    EvState st = st_out;
    uint32_t a_ref = 0, a_read = 0, a_n = 0, a_h = 0;
    uint32_t base = 0;
    for (; base < full; base += 32 * UNROLL) {
        uint4 w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) w[u] = __ldcs(cg + base + u * 32 + lane);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) SCAN_GROUP(w[u])
    }
    for (; base < n4; base += 32) {   // ragged tail: bounds-checked, words past n_cigar zeroed
        const uint32_t idx = base + lane;
        uint4 w = (idx < n4) ? __ldcs(cg + idx) : make_uint4(0, 0, 0, 0);
        if (idx == n4 - 1) {
            const uint32_t rr = n & 3u;
            if (rr == 1) { w.y = 0; w.z = 0; w.w = 0; } else if (rr == 2) { w.z = 0; w.w = 0; } else if (rr == 3) { w.w = 0; }
        }
        SCAN_GROUP(w)
    }
    st.nsum += a_n; st.hsum += a_h;    // lane-private partial sums (warp-reduced by the caller when needed)
    st_out = st; acc_ref_out = a_ref; acc_read_out = a_read;
}

template <int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_cigar_scan(DevSoa a, ChainParams p, SigQueue qm, SigQueue qt, ChainWork* work,
                                                           uint32_t work_cap, uint32_t* cnt) {
    const int lane = threadIdx.x & 31;
    DeviceEmitter out{qm, qt, cnt + CNT_OVERFLOW};
    const uint32_t thresh = p.min_sv <= 0 ? 0u : (p.min_sv >= (1 << 28) ? 0xffffffffu : ((uint32_t)p.min_sv << 4));
    const uint32_t n_aln = (uint32_t)a.n;
    uint32_t primaries = 0;
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(cnt + CNT_NEXT_ALN, (uint32_t)SCAN_BATCH);
        first = __shfl_sync(FULL, first, 0);
        if (first >= n_aln) break;
        const uint32_t last = min(first + SCAN_BATCH, n_aln);
        for (uint32_t i = first; i < last; ++i) {
            const uint32_t flag = a.flag[i];
            if ((flag & 0x104u) || (int32_t)a.mapq[i] < p.min_mapq) continue;   // SVIM_COLLECT.py:143
            const bool primary = !(flag & 0x800u);
            primaries += primary;
            const uint32_t n = a.n_cigar[i];
            const uint4* cg = reinterpret_cast<const uint4*>(a.cigar + a.cigar_off[i]);
            ScanRec r; r.i = i; r.qid = a.qname_id[i]; r.tid = a.tid[i]; r.ref_start = a.pos[i]; r.l_seq = a.l_seq[i];
            EvState st; st.base_ref = 0; st.base_read = 0; st.n_ev = 0; st.n_tw = 0; st.nsum = 0; st.hsum = 0;
            uint32_t acc_ref = 0, acc_read = 0;
            const bool need_summary = primary && a.sa_len[i] > 0;
            if (need_summary) scan_cigar<true, UNROLL>(cg, n, thresh, lane, p, r, st, acc_ref, acc_read, out);
            else scan_cigar<false, UNROLL>(cg, n, thresh, lane, p, r, st, acc_ref, acc_read, out);
            // ---- primaries with an SA tag: summary for the split-read analysis -----------------
            if (need_summary) {
                const uint32_t hard = warp_sum(st.hsum);
                if (hard == 0) {   // SVIM_COLLECT.py:47-48
                    const int64_t ref_q = st.base_ref + warp_sum(acc_ref);
                    const int64_t rd = st.base_read + warp_sum(acc_read);
                    const int64_t nsum = warp_sum(st.nsum);
                    if (lane == 0) {
                        const uint32_t* c32 = a.cigar + a.cigar_off[i];
                        const int64_t l_seq = r.l_seq;
                        CigarSummary cs; cigsum_init(cs);
                        if (l_seq == 0) {   // no SEQ: exact sequential summary (rare)
                            for (uint32_t k = 0; k < n; ++k) cigsum_add(cs, c32[k] & 15u, c32[k] >> 4);
                        } else {
                            cs.ref_len = ref_q + nsum; cs.qlen_h = rd; cs.hard = 0; cs.n_ops = (int32_t)n;
                            uint32_t k = 0;
                            for (; k < n; ++k) { uint32_t op = c32[k] & 15u; if (op == OP_H) continue; if (op != OP_S) break; cs.lead_s += c32[k] >> 4; }
                            for (uint32_t j = n; j-- > 1;) { uint32_t op = c32[j] & 15u; if (op == OP_H) continue; if (op != OP_S) break; cs.trail_s += c32[j] >> 4; }
                        }
                        Seg sg; int64_t rl;
                        cigsum_finish(cs, l_seq, r.ref_start, (flag & 0x10u) ? 1 : 0, sg, rl);
                        uint32_t slot = atomicAdd(cnt + CNT_WORK, 1u);
                        if (slot < work_cap) {
                            ChainWork wk; wk.aln_idx = i; wk.ord_sig = 0x80000000u | st.n_ev; wk.ord_twin = 0x80000000u | st.n_tw; wk.pad = 0;
                            wk.ref_end = sg.ref_end; wk.q_start = sg.q_start; wk.q_end = sg.q_end; wk.read_len = rl;
                            work[slot] = wk;
                        } else atomicExch(cnt + CNT_OVERFLOW, 1u);
                    }
                }
            }
        }
    }
    primaries = warp_sum(primaries);   // every lane counted the same records
    if (lane == 0 && primaries) atomicAdd(cnt + CNT_PRIMARIES, primaries / 32);
}

// ---- default scan kernel ---------------------------------------------------------------------------------------
// The scan moves ~11 KB of CIGAR per record at HBM speed, but per 4-byte op it also has to classify the op and add
// its length to two running sums: with plain LOP3/ISETP/IADD code that is ~12 ALU-pipe instructions per op, and the
// ALU pipe (one warp instruction every two cycles per SM sub-partition) saturates below the HBM roofline.  So:
//   * one funnel shift of a 64-bit constant by the op code yields all three class bits at once
//     (bit 0: consumes reference, bit 15: I/D, bit 31: consumes read);
//   * one IMAD.WIDE on the otherwise idle FMA pipe adds len to both sums, kept as two fields of a 64-bit
//     accumulator (reference in bits 0-30, read from bit 31 up; a valid record consumes < 2^31 of either);
//   * everything the rare path needs (running positions, ordinals, record identity, queues, options) lives in
//     shared memory, and the record metadata of a batch is fetched by four lanes at once, so the hot loop keeps
//     only the 16 loaded words, the accumulator and the cursor in registers (6 CTAs of 256 threads per SM).
struct ScanMeta { uint64_t cig; uint32_t i, n, bits, slot, qid; int32_t tid, pos, l_seq; };   // bits: 1 scan, 2 primary, 4 summary, 8 reverse
struct ScanShared { ChainParams p; SigQueue qm, qt; uint32_t* overflow; uint32_t* holes; uint32_t thresh; int use_chunks; };
struct ScanWarp { EvState st; SlotChunk cm, ct; uint32_t cur, pad; ScanMeta meta[SCAN_BATCH]; };
__shared__ ScanShared g_scan_sh;
__shared__ ScanWarp g_scan_ws[8];

constexpr uint64_t SCAN_K = (uint64_t)SVIM_MASK_REF_QUIRK | ((uint64_t)0x6u << 15) | ((uint64_t)SVIM_MASK_READ << 31);
constexpr uint64_t SCAN_K2 = ((uint64_t)1u << OP_N) | (((uint64_t)1u << OP_H) << 31);      // N / H lengths for the SA summary
#define SCAN_F_MASK 0x80000001u
#define SCAN_F_CAND 0x8000u

__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint32_t acc_lo(uint64_t acc) { return (uint32_t)acc & 0x7fffffffu; }
__device__ __forceinline__ uint32_t acc_hi(uint64_t acc) { return (uint32_t)(acc >> 31); }

// Rare path: a 128-op group with at least one SV-sized I/D.  Warp reductions (redux.sync) give the consumption before
// each event lane; event lanes are visited in order, so ordinals and queue slots stay in emission order.
// `acc` = the lane's consumption since the last fold.  Kept small on purpose (one emission body, no unrolling): the
// scan's hot loop shares the instruction cache with it.
__device__ __noinline__ void scan_events_s(const uint4 w, int lane, uint64_t acc, bool sum) {
    const ScanShared* sh = &g_scan_sh;
    ScanWarp* ws = &g_scan_ws[threadIdx.x >> 5];
    const uint32_t thresh = sh->thresh;
    uint32_t g_ref = 0, g_read = 0, g_n = 0, g_h = 0, n_evt = 0, n_del = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t vk = k == 0 ? w.x : k == 1 ? w.y : k == 2 ? w.z : w.w;
        const uint32_t B = op_bit(vk);
        acc_op<true>(vk, B, g_ref, g_read, g_n, g_h);
        if (is_event(vk, B, thresh)) { ++n_evt; n_del += (vk & 15u) == OP_D; }
    }
    unsigned m = __ballot_sync(FULL, n_evt != 0);
    EvState st = ws->st;
    const ScanMeta r = ws->meta[ws->cur];
    const int all_bnds = sh->p.all_bnds;
    __syncwarp();
    st.base_ref += __reduce_add_sync(FULL, acc_lo(acc)); st.base_read += __reduce_add_sync(FULL, acc_hi(acc));
#pragma unroll 1
    while (m) {
        const int L = __ffs(m) - 1; m &= m - 1;
        const uint32_t ex_ref = __reduce_add_sync(FULL, lane < L ? g_ref : 0u);
        const uint32_t ex_read = __reduce_add_sync(FULL, lane < L ? g_read : 0u);
        const uint32_t cntL = __shfl_sync(FULL, n_evt, L);
        const uint32_t ndel = all_bnds ? __shfl_sync(FULL, n_del, L) : 0u;
        uint32_t slot = 0, slot_t = 0;
        if (sh->use_chunks) { slot = reserve_slots(sh->qm, &ws->cm, cntL, lane, sh->holes); if (ndel) slot_t = reserve_slots(sh->qt, &ws->ct, ndel, lane, sh->holes + 1); }
        else if (lane == L) { slot = atomicAdd(sh->qm.count, cntL); if (ndel) slot_t = atomicAdd(sh->qt.count, ndel); }
        if (lane == L) {
            uint32_t ord = st.n_ev, tord = st.n_tw, cr = 0, cq = 0, dn = 0, dh = 0;
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const uint32_t vk = k == 0 ? w.x : k == 1 ? w.y : k == 2 ? w.z : w.w;
                const uint32_t B = op_bit(vk);
                if (is_event(vk, B, thresh)) {
                    const int64_t len = vk >> 4;
                    const int64_t pr = st.base_ref + ex_ref + cr;
                    const int64_t pq = st.base_read + ex_read + cq;
                    svim_sig s; memset(&s, 0, sizeof(s));
                    s.contig1 = r.tid; s.contig2 = -1; s.start = (int32_t)(r.pos + pr); s.end = (int32_t)(r.pos + pr + len);
                    s.aln_idx = r.i; s.qname_id = r.qid; s.ordinal = ord++;
                    if ((vk & 15u) == OP_D) {
                        s.type = SVIM_DEL;
                        if (all_bnds) {   // SVIM_intra.py:43-44 (same contig, start < end: already canonical)
                            svim_sig t = s; t.type = SVIM_BND; t.contig2 = r.tid; t.pos = s.end; t.end = s.start + 1; t.ordinal = tord++;
                            if (len == 0) { t.flags = SVIM_F_DIR1_REV | SVIM_F_DIR2_REV; }   // pos1 == pos2: the else-branch flips both directions
                            if (slot_t < sh->qt.cap) sh->qt.recs[slot_t] = t; else atomicExch(sh->overflow, 1u);
                            ++slot_t;
                        }
                    } else {
                        s.type = SVIM_INS;
                        int64_t lo, hi; py_slice(pq, len, (int64_t)r.l_seq, lo, hi);   // query_sequence[pos_read:pos_read+len]
                        s.seq_off = (uint64_t)lo; s.seq_len = (uint32_t)(hi - lo);
                    }
                    if (slot < sh->qm.cap) sh->qm.recs[slot] = s; else atomicExch(sh->overflow, 1u);
                    ++slot;
                }
                acc_op<true>(vk, B, cr, cq, dn, dh);
            }
        }
        st.n_ev += cntL; st.n_tw += ndel;
    }
    st.base_ref += __reduce_add_sync(FULL, g_ref); st.base_read += __reduce_add_sync(FULL, g_read);
    if (sum) { st.nsum += __reduce_add_sync(FULL, g_n); st.hsum += __reduce_add_sync(FULL, g_h); }   // the shared copy keeps warp totals
    __syncwarp();
    if (lane == 0) ws->st = st;
    __syncwarp();
}

// Class bits of the four ops of a uint4 and the warp's event ballot.  Written in PTX so that each step stays one
// instruction: and / funnel shift / and (multiplier) / and+setp (-> LOP3 with predicate output) / setp.ge.and.
__device__ __forceinline__ uint32_t scan_classify(const uint4 w, uint32_t thresh, uint32_t& f0, uint32_t& f1, uint32_t& f2, uint32_t& f3) {
    uint32_t ballot;
    asm volatile(
        "{\n\t"
        ".reg .b32 o, t, c;\n\t"
        ".reg .pred pc, p0, p1, p2, p3;\n\t"
        "and.b32 o, %5, 15;\n\t shf.r.wrap.b32 t, %10, %11, o;\n\t and.b32 %1, t, 0x80000001;\n\t and.b32 c, t, 0x8000;\n\t"
        "setp.ne.u32 pc, c, 0;\n\t setp.ge.and.u32 p0, %5, %9, pc;\n\t"
        "and.b32 o, %6, 15;\n\t shf.r.wrap.b32 t, %10, %11, o;\n\t and.b32 %2, t, 0x80000001;\n\t and.b32 c, t, 0x8000;\n\t"
        "setp.ne.u32 pc, c, 0;\n\t setp.ge.and.u32 p1, %6, %9, pc;\n\t"
        "and.b32 o, %7, 15;\n\t shf.r.wrap.b32 t, %10, %11, o;\n\t and.b32 %3, t, 0x80000001;\n\t and.b32 c, t, 0x8000;\n\t"
        "setp.ne.u32 pc, c, 0;\n\t setp.ge.and.u32 p2, %7, %9, pc;\n\t"
        "and.b32 o, %8, 15;\n\t shf.r.wrap.b32 t, %10, %11, o;\n\t and.b32 %4, t, 0x80000001;\n\t and.b32 c, t, 0x8000;\n\t"
        "setp.ne.u32 pc, c, 0;\n\t setp.ge.and.u32 p3, %8, %9, pc;\n\t"
        "or.pred p0, p0, p1;\n\t or.pred p2, p2, p3;\n\t or.pred p0, p0, p2;\n\t"
        "vote.sync.ballot.b32 %0, p0, 0xffffffff;\n\t"
        "}"
        : "=r"(ballot), "=r"(f0), "=r"(f1), "=r"(f2), "=r"(f3)
        : "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w), "r"(thresh), "r"((uint32_t)SCAN_K), "r"((uint32_t)(SCAN_K >> 32)));
    return ballot;
}
__device__ __forceinline__ uint32_t scan_f2(uint32_t v) { return __funnelshift_r((uint32_t)SCAN_K2, (uint32_t)(SCAN_K2 >> 32), v & 15u) & SCAN_F_MASK; }

#define SCAN_GROUP_S(W)                                                                                            \
    {                                                                                                              \
        uint32_t f0, f1, f2, f3;                                                                                   \
        if (scan_classify((W), thresh, f0, f1, f2, f3) == 0) {                                                     \
            acc = mad_wide((W).x >> 4, f0, acc); acc = mad_wide((W).y >> 4, f1, acc);                              \
            acc = mad_wide((W).z >> 4, f2, acc); acc = mad_wide((W).w >> 4, f3, acc);                              \
            if (sum) { acc2 = mad_wide((W).x >> 4, scan_f2((W).x), acc2); acc2 = mad_wide((W).y >> 4, scan_f2((W).y), acc2); \
                       acc2 = mad_wide((W).z >> 4, scan_f2((W).z), acc2); acc2 = mad_wide((W).w >> 4, scan_f2((W).w), acc2); } \
        } else {                                                                                                   \
            scan_events_s((W), lane, acc, sum);                                                                    \
            acc = 0;                                                                                               \
        }                                                                                                          \
    }

// acc: reference (bits 0-30) / read (bits 31+) consumption of this lane since the last fold; acc2: N / H lengths of
// the groups that took the fast path (`sum`: records whose summary feeds the split-read analysis).
// No unrolling beyond one 4-load block and a small rare path: the instruction cache is part of the budget.
// LD = 128-bit loads in flight per lane
template <int LD>
__device__ __forceinline__ void scan_cigar_s(const uint4* __restrict__ cg, uint32_t n, uint32_t thresh, int lane, bool sum, uint64_t& acc_out, uint64_t& acc2_out) {
    const uint32_t n4 = (n + 3) >> 2;
    const uint32_t full = (n >> 2) / (LD * 32) * (LD * 32);
    uint64_t acc = 0, acc2 = 0;
    uint32_t base = 0;
#pragma unroll 1
    for (; base < full; base += LD * 32) {
        uint4 w[LD];
#pragma unroll
        for (int u = 0; u < LD; ++u) w[u] = __ldcs(cg + base + u * 32 + lane);
#pragma unroll
        for (int u = 0; u < LD; ++u) SCAN_GROUP_S(w[u])
    }
    const uint32_t last = n4 - 1, rr = n & 3u;
#pragma unroll 1
    for (; base < n4; base += LD * 32) {   // ragged tail: loads in flight together, words past n_cigar zeroed
        uint4 w[LD];
#pragma unroll
        for (int u = 0; u < LD; ++u) {
            const uint32_t idx = base + u * 32 + lane;
            uint4 t = (idx <= last) ? __ldcs(cg + idx) : make_uint4(0, 0, 0, 0);
            const uint32_t valid = (idx == last && rr) ? rr : 4u;
            t.y = valid > 1 ? t.y : 0u; t.z = valid > 2 ? t.z : 0u; t.w = valid > 3 ? t.w : 0u;
            w[u] = t;
        }
#pragma unroll
        for (int u = 0; u < LD; ++u) { if (u == 0 || base + u * 32 < n4) SCAN_GROUP_S(w[u]) }
    }
    acc_out = acc; acc2_out = acc2;
}

// query-sorted mode (SVIM_COLLECT.py:96-129): the host grouped the records by read; qs.info[i] = role | slot << 3 with
// role 0 = skip, 1 = the read's only primary, 2 = good supplementary, bit 2 = the read has segments to chain.
struct QsView { const uint32_t* info; const uint32_t* grp; SegSum* segsum; };

template <int MINB, bool QS, int LD>
__global__ void __launch_bounds__(256, MINB) k_cigar_scan_s(DevSoa a, ChainParams p, SigQueue qm, SigQueue qt, ChainWork* work, uint32_t work_cap,
                                                             uint32_t* cnt, QsView qs, int use_chunks, uint32_t thresh) {
    const int lane = threadIdx.x & 31;
    ScanShared* sh = &g_scan_sh;
    ScanWarp* ws = &g_scan_ws[threadIdx.x >> 5];
    if (threadIdx.x == 0) { sh->p = p; sh->qm = qm; sh->qt = qt; sh->overflow = cnt + CNT_OVERFLOW; sh->holes = cnt + CNT_HOLES_MAIN; sh->thresh = thresh; sh->use_chunks = use_chunks; }
    if (lane == 0) { ws->cm = SlotChunk{0, 0}; ws->ct = SlotChunk{0, 0}; }
    __syncthreads();
    const uint32_t n_aln = (uint32_t)a.n;
    uint32_t primaries = 0;
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(cnt + CNT_NEXT_ALN, (uint32_t)SCAN_BATCH);
        first = __shfl_sync(FULL, first, 0);
        if (first >= n_aln) break;
        // ---- the batch's record metadata: one lane per record, every field load in flight at once ----
        if (lane < SCAN_BATCH) {
            const uint32_t i = first + lane;
            ScanMeta m; m.i = i; m.bits = 0; m.n = 0; m.cig = 0; m.slot = 0; m.qid = 0; m.tid = 0; m.pos = 0; m.l_seq = 0;
            if (i < n_aln) {
                const uint32_t flag = a.flag[i];
                bool take, primary, summary;
                if (QS) {
                    const uint32_t info = qs.info[i];
                    take = (info & 3u) != 0; primary = (info & 3u) == 1; summary = (info & 4u) != 0; m.slot = info >> 3;
                } else {
                    take = !((flag & 0x104u) || (int32_t)a.mapq[i] < p.min_mapq);      // SVIM_COLLECT.py:143
                    primary = !(flag & 0x800u); summary = primary && a.sa_len[i] > 0;
                }
                if (take) {
                    m.bits = 1u | (primary ? 2u : 0u) | (summary ? 4u : 0u) | ((flag & 0x10u) ? 8u : 0u);
                    m.n = a.n_cigar[i]; m.cig = a.cigar_off[i]; m.qid = a.qname_id[i]; m.tid = a.tid[i]; m.pos = a.pos[i]; m.l_seq = a.l_seq[i];
                    primaries += primary;
                }
            }
            ws->meta[lane] = m;
        }
        __syncwarp();
        for (int b = 0; b < SCAN_BATCH; ++b) {
            const uint32_t bits = ws->meta[b].bits;
            if (!(bits & 1u)) continue;
            const uint32_t n = ws->meta[b].n;
            const uint4* cg = reinterpret_cast<const uint4*>(a.cigar + ws->meta[b].cig);
            if (lane == 0) {
                const uint32_t slot = ws->meta[b].slot;
                EvState st; st.base_ref = 0; st.base_read = 0; st.n_ev = slot << 20; st.n_tw = slot << 20; st.nsum = 0; st.hsum = 0;
                ws->st = st; ws->cur = (uint32_t)b;
            }
            __syncwarp();
            uint64_t acc = 0, acc2 = 0;
            const bool need_summary = bits & 4u;
            scan_cigar_s<LD>(cg, n, thresh, lane, need_summary, acc, acc2);
            if (need_summary) {
                const EvState st = ws->st;     // st.nsum / st.hsum: warp totals of the rare-path groups (scan_events_s)
                const uint32_t hard = __reduce_add_sync(FULL, acc_hi(acc2)) + st.hsum;
                if (hard == 0 || QS) {     // the hard-clip rule belongs to the SA reconstruction only (SVIM_COLLECT.py:47)
                    const int64_t ref_q = st.base_ref + __reduce_add_sync(FULL, acc_lo(acc));
                    const int64_t rd = st.base_read + __reduce_add_sync(FULL, acc_hi(acc));
                    const int64_t nsum = (int64_t)__reduce_add_sync(FULL, acc_lo(acc2)) + st.nsum;
                    if (lane == 0) {
                        const ScanMeta m = ws->meta[b];
                        const uint32_t* c32 = a.cigar + m.cig;
                        const int64_t l_seq = m.l_seq;
                        CigarSummary cs; cigsum_init(cs);
                        if (l_seq == 0) {   // no SEQ: exact sequential summary (rare)
                            for (uint32_t k = 0; k < n; ++k) cigsum_add(cs, c32[k] & 15u, c32[k] >> 4);
                        } else {
                            cs.ref_len = ref_q + nsum; cs.qlen_h = rd + hard; cs.hard = hard; cs.n_ops = (int32_t)n;   // infer_read_length counts H (query-sorted mode keeps hard-clipped records)
                            uint32_t k = 0;
                            for (; k < n; ++k) { uint32_t op = c32[k] & 15u; if (op == OP_H) continue; if (op != OP_S) break; cs.lead_s += c32[k] >> 4; }
                            for (uint32_t j = n; j-- > 1;) { uint32_t op = c32[j] & 15u; if (op == OP_H) continue; if (op != OP_S) break; cs.trail_s += c32[j] >> 4; }
                        }
                        Seg sg; int64_t rl;
                        cigsum_finish(cs, l_seq, m.pos, (m.bits & 8u) ? 1 : 0, sg, rl);
                        if (QS) { SegSum ss; ss.ref_end = sg.ref_end; ss.q_start = sg.q_start; ss.q_end = sg.q_end; ss.read_len = rl; qs.segsum[m.i] = ss; }
                        if (!QS || (m.bits & 2u)) {
                            uint32_t wslot = atomicAdd(cnt + CNT_WORK, 1u);
                            if (wslot < work_cap) {
                                ChainWork wk; wk.aln_idx = m.i; wk.pad = 0;
                                wk.ord_sig = QS ? 0xFFF00000u : (0x80000000u | st.n_ev); wk.ord_twin = QS ? 0xFFF00000u : (0x80000000u | st.n_tw);
                                wk.ref_end = sg.ref_end; wk.q_start = sg.q_start; wk.q_end = sg.q_end; wk.read_len = rl;
                                work[wslot] = wk;
                            } else atomicExch(cnt + CNT_OVERFLOW, 1u);
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    flush_slots(qm, &ws->cm, lane, cnt + CNT_HOLES_MAIN);
    if (p.all_bnds) flush_slots(qt, &ws->ct, lane, cnt + CNT_HOLES_TWIN);
    primaries = warp_sum(primaries);   // lanes counted their own records
    if (lane == 0 && primaries) atomicAdd(cnt + CNT_PRIMARIES, primaries);
}

// ---- bulk-copy (TMA engine) variant of the scan ---------------------------------------------------------
// Each warp owns a ring of SCAN_STAGES shared-memory stages of 2 KiB (one 512-op block).  Lane 0 issues
// cp.async.bulk global->shared copies that complete on a per-stage mbarrier; in-flight bytes are bounded by
// shared memory, not registers, and the copy engine runs ahead across the records of the warp's batch.
#define SCAN_STAGES 4
#define SCAN_BLOCK_U4 128                  // uint4 per stage (2 KiB)
#define SCAN_BULK_WARPS 8
#define SCAN_BULK_BATCH 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(32 * SCAN_BULK_WARPS, 3) k_cigar_scan_bulk(DevSoa a, ChainParams p, SigQueue qm, SigQueue qt, ChainWork* work,
                                                                            uint32_t work_cap, uint32_t* cnt) {
    extern __shared__ __align__(128) unsigned char scan_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint4* ring = reinterpret_cast<uint4*>(scan_smem) + (size_t)wib * SCAN_STAGES * SCAN_BLOCK_U4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(scan_smem + (size_t)SCAN_BULK_WARPS * SCAN_STAGES * SCAN_BLOCK_U4 * 16) + wib * SCAN_STAGES;
    if (lane == 0) {
        for (int s = 0; s < SCAN_STAGES; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    DeviceEmitter out{qm, qt, cnt + CNT_OVERFLOW};
    const uint32_t thresh = p.min_sv <= 0 ? 0u : (p.min_sv >= (1 << 28) ? 0xffffffffu : ((uint32_t)p.min_sv << 4));
    const uint32_t n_aln = (uint32_t)a.n;
    uint32_t primaries = 0;
    uint32_t phase = 0;                       // bit s = parity the consumer waits for on stage s
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(cnt + CNT_NEXT_ALN, (uint32_t)SCAN_BULK_BATCH);
        first = __shfl_sync(FULL, first, 0);
        if (first >= n_aln) break;
        const uint32_t nrec = min((uint32_t)SCAN_BULK_BATCH, n_aln - first);
        // per-record block counts of this batch (lane r holds record r); filtered records have none
        uint32_t my_n = 0; const uint4* my_cg = nullptr; bool my_ok = false;
        if (lane < nrec) {
            const uint32_t i = first + lane;
            const uint32_t flag = a.flag[i];
            my_ok = !((flag & 0x104u) || (int32_t)a.mapq[i] < p.min_mapq);
            if (my_ok) { my_n = a.n_cigar[i]; my_cg = reinterpret_cast<const uint4*>(a.cigar + a.cigar_off[i]); }
        }
        const uint32_t my_blk = (((my_n + 3) >> 2) + SCAN_BLOCK_U4 - 1) / SCAN_BLOCK_U4;
        uint32_t tot_blk; const uint32_t blk_before = warp_excl_scan(my_blk, lane, tot_blk);
        // producer state (lane 0 issues): global block index -> (record lane, local block)
        uint32_t prod = 0;
        auto issue = [&](uint32_t g) {   // executed by all lanes (shuffles), copy issued by lane 0
            // owner = last lane with blk_before <= g and my_blk > 0
            const unsigned own = __ballot_sync(FULL, my_blk > 0 && blk_before <= g && g < blk_before + my_blk);
            const int ol = __ffs(own) - 1;
            const uint32_t lb = g - __shfl_sync(FULL, blk_before, ol);
            const uint32_t n4 = (__shfl_sync(FULL, my_n, ol) + 3) >> 2;
            const unsigned long long base = __shfl_sync(FULL, (unsigned long long)my_cg, ol);
            const uint32_t cnt4 = min((uint32_t)SCAN_BLOCK_U4, n4 - lb * SCAN_BLOCK_U4);
            if (lane == 0) {
                const int st = g % SCAN_STAGES;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this stage vs the async write
                mbar_expect_tx(bars + st, cnt4 * 16);
                bulk_g2s(ring + (size_t)st * SCAN_BLOCK_U4, reinterpret_cast<const uint4*>(base) + (size_t)lb * SCAN_BLOCK_U4, cnt4 * 16, bars + st);
            }
        };
        for (; prod < tot_blk && prod < SCAN_STAGES; ++prod) issue(prod);
        uint32_t g = 0;                        // consumer block index within the batch
        // NOTE: stages are indexed by g % SCAN_STAGES, so the ring restarts at stage 0 every batch; the phase word keeps
        // the per-stage parity across batches.
        for (uint32_t ri = 0; ri < nrec; ++ri) {
            const bool ok = __shfl_sync(FULL, (int)my_ok, ri);
            if (!ok) continue;
            const uint32_t i = first + ri;
            const uint32_t flag = a.flag[i];
            const bool primary = !(flag & 0x800u);
            primaries += primary;
            const uint32_t n = __shfl_sync(FULL, my_n, ri);
            const uint32_t n4 = (n + 3) >> 2;
            const uint32_t nb = __shfl_sync(FULL, my_blk, ri);
            ScanRec r; r.i = i; r.qid = a.qname_id[i]; r.tid = a.tid[i]; r.ref_start = a.pos[i]; r.l_seq = a.l_seq[i];
            EvState st; st.base_ref = 0; st.base_read = 0; st.n_ev = 0; st.n_tw = 0; st.nsum = 0; st.hsum = 0;
            uint32_t a_ref = 0, a_read = 0, a_n = 0, a_h = 0;
            const bool need_summary = primary && a.sa_len[i] > 0;
            for (uint32_t b = 0; b < nb; ++b, ++g) {
                const int stg = g % SCAN_STAGES;
                mbar_wait(bars + stg, (phase >> stg) & 1u);
                phase ^= 1u << stg;
                const uint4* blk = ring + (size_t)stg * SCAN_BLOCK_U4;
                uint4 w[4];
                const uint32_t base4 = b * SCAN_BLOCK_U4;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t idx = base4 + u * 32 + lane;
                    w[u] = idx < n4 ? blk[u * 32 + lane] : make_uint4(0, 0, 0, 0);
                    if (idx == n4 - 1) {
                        const uint32_t rr = n & 3u;
                        if (rr == 1) { w[u].y = 0; w[u].z = 0; w[u].w = 0; } else if (rr == 2) { w[u].z = 0; w[u].w = 0; } else if (rr == 3) { w[u].w = 0; }
                    }
                }
                __syncwarp();                                  // every lane has its words: the stage can be refilled
                if (prod < tot_blk) { issue(prod); ++prod; }
                if (need_summary) {
                    constexpr bool SUM = true;
#pragma unroll
                    for (int u = 0; u < 4; ++u) SCAN_GROUP(w[u])
                } else {
                    constexpr bool SUM = false;
#pragma unroll
                    for (int u = 0; u < 4; ++u) SCAN_GROUP(w[u])
                }
            }
            st.nsum += a_n; st.hsum += a_h;
            if (need_summary) {
                const uint32_t hard = warp_sum(st.hsum);
                if (hard == 0) {
                    const int64_t ref_q = st.base_ref + warp_sum(a_ref);
                    const int64_t rd = st.base_read + warp_sum(a_read);
                    const int64_t nsum = warp_sum(st.nsum);
                    if (lane == 0) {
                        const uint32_t* c32 = a.cigar + a.cigar_off[i];
                        const int64_t l_seq = r.l_seq;
                        CigarSummary cs; cigsum_init(cs);
                        if (l_seq == 0) {
                            for (uint32_t k = 0; k < n; ++k) cigsum_add(cs, c32[k] & 15u, c32[k] >> 4);
                        } else {
                            cs.ref_len = ref_q + nsum; cs.qlen_h = rd; cs.hard = 0; cs.n_ops = (int32_t)n;
                            uint32_t k = 0;
                            for (; k < n; ++k) { uint32_t op = c32[k] & 15u; if (op == OP_H) continue; if (op != OP_S) break; cs.lead_s += c32[k] >> 4; }
                            for (uint32_t j = n; j-- > 1;) { uint32_t op = c32[j] & 15u; if (op == OP_H) continue; if (op != OP_S) break; cs.trail_s += c32[j] >> 4; }
                        }
                        Seg sg; int64_t rl;
                        cigsum_finish(cs, l_seq, r.ref_start, (flag & 0x10u) ? 1 : 0, sg, rl);
                        uint32_t slot = atomicAdd(cnt + CNT_WORK, 1u);
                        if (slot < work_cap) {
                            ChainWork wk; wk.aln_idx = i; wk.ord_sig = 0x80000000u | st.n_ev; wk.ord_twin = 0x80000000u | st.n_tw; wk.pad = 0;
                            wk.ref_end = sg.ref_end; wk.q_start = sg.q_start; wk.q_end = sg.q_end; wk.read_len = rl;
                            work[slot] = wk;
                        } else atomicExch(cnt + CNT_OVERFLOW, 1u);
                    }
                }
            }
        }
        // the ring restarts at stage 0 for the next batch: rotate the phase word so bit s still belongs to stage s
        // (tot_blk blocks were consumed; stage s was used ceil((tot_blk - s)/STAGES) times — already folded into `phase`)
    }
    primaries = warp_sum(primaries);
    if (lane == 0 && primaries) atomicAdd(cnt + CNT_PRIMARIES, primaries / 32);
}

// svim_aln_soa.cigar16 -> BAM uint32 CIGAR words (include/svimgpu.h).  One warp per record: every lane takes 8 packed words
// (one 128-bit load) per round; a word without the 0xF nibble is its own uint32 image, so the common case is a widening copy
// at HBM speed.  Rounds that hold a length-extension word (operations of 4096 bases and more) are decoded by lane 0 in order.
__global__ void __launch_bounds__(256) k_expand_cigar16(const uint16_t* __restrict__ c16, const uint64_t* __restrict__ off16, const uint32_t* __restrict__ n_cigar,
                                                         const uint64_t* __restrict__ cigar_off, int64_t n, uint32_t* __restrict__ cigar, uint32_t* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
        const uint16_t* src = c16 + off16[i];
        const uint64_t nw = off16[i + 1] - off16[i];
        uint32_t* dst = cigar + cigar_off[i];
        const uint32_t nc = n_cigar[i];
        uint64_t out = 0;            // operations written so far (warp-uniform)
        uint64_t acc = 0;            // pending length-extension bits (warp-uniform: only lane 0's slow path changes it, then broadcasts)
        for (uint64_t base = 0; base < nw; base += 256) {
            const uint64_t at = base + (uint64_t)lane * 8;
            uint4 v = make_uint4(0x000F000Fu, 0x000F000Fu, 0x000F000Fu, 0x000F000Fu);
            if (at < nw) v = *(const uint4*)(src + at);
            const uint32_t w2[4] = {v.x, v.y, v.z, v.w};
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { w[2 * k] = w2[k] & 0xffffu; w[2 * k + 1] = w2[k] >> 16; }
            uint32_t ext = 0, real_ext = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const uint32_t e = (w[k] & 15u) == 15u; ext |= e << k; real_ext |= (e && w[k] != 0x000Fu) ? 1u : 0u; }
            if (__any_sync(0xffffffffu, real_ext) || acc) {
                // in-order decode of this round by lane 0 (rare: only operations of >= 4096 bases need extension words)
                uint64_t o = out, a = acc;
                for (int l = 0; l < 32; ++l) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t x = __shfl_sync(0xffffffffu, w[k], l);
                        if (lane == 0) {
                            if ((x & 15u) == 15u) a = (a << 12) | (x >> 4);
                            else { const uint64_t len = (a << 12) | (x >> 4); if (o < nc) dst[o] = (uint32_t)((len << 4) | (x & 15u)); ++o; a = 0; if (len >= (1ull << 28)) atomicExch(bad, 1u); }
                        }
                    }
                }
                out = __shfl_sync(0xffffffffu, o, 0); acc = __shfl_sync(0xffffffffu, a, 0);
                continue;
            }
            const uint32_t cnt = 8u - (uint32_t)__popc(ext);
            uint32_t pre = cnt;                              // inclusive warp scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
            const uint64_t mine = out + pre - cnt;
            if (cnt == 8u && mine + 8 <= nc && ((mine & 3) == 0)) {
                *(uint4*)(dst + mine) = make_uint4(w[0], w[1], w[2], w[3]); *(uint4*)(dst + mine + 4) = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
                uint64_t o2 = mine;
#pragma unroll
                for (int k = 0; k < 8; ++k) if (!((ext >> k) & 1u)) { if (o2 < nc) dst[o2] = w[k]; ++o2; }
            }
            out += __shfl_sync(0xffffffffu, pre, 31);
        }
        if (out != nc || acc) { if (lane == 0) atomicExch(bad, 1u); }
        for (uint32_t k = nc + lane; k < ((nc + 3u) & ~3u); k += 32) dst[k] = 0u;      // records are padded to 16 bytes with zero words
    }
}

// svim_aln_soa.cigar8 -> BAM uint32 CIGAR words.  One warp per record, 16 bytes (one 128-bit load) per lane and round.  A lane needs
// two things from its neighbours: where its first operation lands (warp scan of the per-lane operation counts) and the extension
// bits left pending by the bytes before it (the previous lane's c8_tail; lane 0 takes lane 31's of the previous round).
__global__ void __launch_bounds__(256) k_expand_cigar8(const uint8_t* __restrict__ c8, const uint64_t* __restrict__ off8, const uint32_t* __restrict__ n_cigar,
                                                        const uint64_t* __restrict__ cigar_off, int64_t n, uint32_t* __restrict__ cigar, uint32_t* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
        const uint8_t* src = c8 + off8[i];
        const uint64_t nb = off8[i + 1] - off8[i];
        uint32_t* dst = cigar + cigar_off[i];
        const uint64_t nc = n_cigar[i];
        uint64_t out = 0;            // operations written so far (warp-uniform)
        uint32_t carry = 0;          // extension bits pending at the end of the previous round (warp-uniform)
        uint32_t err = 0;
        for (uint64_t base = 0; base < nb; base += 512) {
            const uint64_t at = base + (uint64_t)lane * 16;
            uint4 v = make_uint4(0x0F0F0F0Fu, 0x0F0F0F0Fu, 0x0F0F0F0Fu, 0x0F0F0F0Fu);
            if (at < nb) v = *(const uint4*)(src + at);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t cnt = 16u - c8_ext_bytes(w);
            uint32_t pre = cnt;                              // inclusive warp scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
            const uint32_t tail = c8_tail(w);
            uint32_t init = __shfl_up_sync(0xffffffffu, tail, 1);
            if (lane == 0) init = carry;
            c8_decode_lane(w, init, out + pre - cnt, nc, dst, &err);
            out += __shfl_sync(0xffffffffu, pre, 31);
            carry = __shfl_sync(0xffffffffu, tail, 31);
        }
        if (__any_sync(0xffffffffu, err) || out != nc || carry) { if (lane == 0) atomicExch(bad, 1u); }
        for (uint32_t k = (uint32_t)nc + lane; k < (((uint32_t)nc + 3u) & ~3u); k += 32) dst[k] = 0u;      // records are padded to 16 bytes with zero words
    }
}

// The same expansion with the round's words staged in shared memory: the lanes decode into the warp's 512-word staging area (their
// slots are ~15 words apart: an odd stride, few bank conflicts) and the warp then writes the round out as whole 128-byte lines.
// k_expand_cigar8 stores straight from the decode loop instead: a warp store instruction touches ~16 lines there.
__global__ void __launch_bounds__(256) k_expand_cigar8_staged(const uint8_t* __restrict__ c8, const uint64_t* __restrict__ off8, const uint32_t* __restrict__ n_cigar,
                                                               const uint64_t* __restrict__ cigar_off, int64_t n, uint32_t* __restrict__ cigar, uint32_t* __restrict__ bad) {
    __shared__ uint32_t stage_all[8][512];
    uint32_t* stage = stage_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
        const uint8_t* src = c8 + off8[i];
        const uint64_t nb = off8[i + 1] - off8[i];
        uint32_t* dst = cigar + cigar_off[i];
        const uint64_t nc = n_cigar[i];
        uint64_t out = 0;
        uint32_t carry = 0, err = 0;
        for (uint64_t base = 0; base < nb; base += 512) {
            const uint64_t at = base + (uint64_t)lane * 16;
            uint4 v = make_uint4(0x0F0F0F0Fu, 0x0F0F0F0Fu, 0x0F0F0F0Fu, 0x0F0F0F0Fu);
            if (at < nb) v = *(const uint4*)(src + at);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t cnt = 16u - c8_ext_bytes(w);
            uint32_t pre = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
            const uint32_t tail = c8_tail(w);
            uint32_t init = __shfl_up_sync(0xffffffffu, tail, 1);
            if (lane == 0) init = carry;
            c8_decode_lane(w, init, pre - cnt, 512, stage, &err);
            __syncwarp();
            const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
            const int shift = (int)(((uintptr_t)(dst + out) >> 2) & 31u);      // lanes line up with the 128-byte lines of the destination
            for (int k = lane - shift; k < (int)total; k += 32)
                if (k >= 0 && out + (uint64_t)k < nc) dst[out + (uint64_t)k] = stage[k];
            __syncwarp();
            out += total;
            carry = __shfl_sync(0xffffffffu, tail, 31);
        }
        if (__any_sync(0xffffffffu, err) || out != nc || carry) { if (lane == 0) atomicExch(bad, 1u); }
        for (uint32_t k = (uint32_t)nc + lane; k < (((uint32_t)nc + 3u) & ~3u); k += 32) dst[k] = 0u;
    }
}

__global__ void __launch_bounds__(128) k_segment_chain(DevSoa a, ChainParams p, ContigTable ct, const ChainWork* work, uint32_t n_work,
                                                        SigQueue qm, SigQueue qt, uint32_t* cnt, uint32_t* big_list) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_work) return;
    const ChainWork wk = work[w];
    const uint32_t i = wk.aln_idx;
    DeviceEmitter out{qm, qt, cnt + CNT_OVERFLOW};
    Seg chain[SVIM_MAX_SEGMENTS];
    int n = 0;
    uint32_t err = 0;
    const uint32_t flag = a.flag[i];
    const int32_t rev = (flag & 0x10u) ? 1 : 0;
    if (rev && wk.read_len < 0) err |= CH_NO_READLEN;      // SVIM_inter.py:31-34
    else {
        Seg s; s.tid = a.tid[i]; s.ref_start = a.pos[i]; s.ref_end = wk.ref_end; s.q_start = wk.q_start; s.q_end = wk.q_end; s.rev = rev;
        chain[n++] = s;
    }
    n = parse_sa_segments(a.sa + a.sa_off[i], (int)a.sa_len[i], ct, p, a.l_seq[i], chain, n, err);
    if (err & CH_TOO_MANY) { big_list[atomicAdd(cnt + CNT_TOO_MANY, 1u)] = w; return; }      // more segments than the local arrays hold: nothing emitted here
    sort_chain(chain, n);
    PrimaryInfo pi; pi.aln_idx = i; pi.qname_id = a.qname_id[i]; pi.l_seq = a.l_seq[i]; pi.read_len = wk.read_len;
    Junction junc[SVIM_MAX_SEGMENTS]; Tandem tand[SVIM_MAX_SEGMENTS];
    if (n >= 2) analyze_chain(chain, n, pi, p, ct, out, wk.ord_sig, wk.ord_twin, err, junc, tand);
    if (err & CH_BAD_FIELDS) atomicAdd(cnt + CNT_BAD_FIELDS, 1u);
    if (err & CH_NO_READLEN) atomicAdd(cnt + CNT_NO_READLEN, 1u);
    if (err & CH_DATA_ERROR) atomicAdd(cnt + CNT_DATA_ERR, 1u);
}

// query-sorted mode: segments are the read's REAL supplementary records (analyze_alignment_file_querysorted,
// SVIM_COLLECT.py:113-123), listed per read group by the host in file order
__global__ void __launch_bounds__(128) k_segment_chain_qs(DevSoa a, ChainParams p, ContigTable ct, const ChainWork* work, uint32_t n_work,
                                                           const uint32_t* grp, const uint32_t* mem_off, const uint32_t* mem_idx, const SegSum* segsum,
                                                           SigQueue qm, SigQueue qt, uint32_t* cnt, uint32_t* big_list) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_work) return;
    const ChainWork wk = work[w];
    const uint32_t i = wk.aln_idx;
    DeviceEmitter out{qm, qt, cnt + CNT_OVERFLOW};
    Seg chain[SVIM_MAX_SEGMENTS];
    int n = 0;
    uint32_t err = 0;
    const int32_t rev = (a.flag[i] & 0x10u) ? 1 : 0;
    if (rev && wk.read_len < 0) err |= CH_NO_READLEN;
    else { Seg s; s.tid = a.tid[i]; s.ref_start = a.pos[i]; s.ref_end = wk.ref_end; s.q_start = wk.q_start; s.q_end = wk.q_end; s.rev = rev; chain[n++] = s; }
    const uint32_t g = grp[i];
    for (uint32_t m = mem_off[g]; m < mem_off[g + 1]; ++m) {
        const uint32_t j = mem_idx[m];
        const SegSum ss = segsum[j];
        const int32_t jr = (a.flag[j] & 0x10u) ? 1 : 0;
        if (jr && ss.read_len < 0) { err |= CH_NO_READLEN; continue; }
        if (n >= SVIM_MAX_SEGMENTS) { big_list[atomicAdd(cnt + CNT_TOO_MANY, 1u)] = w; return; }
        Seg s; s.tid = a.tid[j]; s.ref_start = a.pos[j]; s.ref_end = ss.ref_end; s.q_start = ss.q_start; s.q_end = ss.q_end; s.rev = jr;
        chain[n++] = s;
    }
    sort_chain(chain, n);
    PrimaryInfo pi; pi.aln_idx = i; pi.qname_id = a.qname_id[i]; pi.l_seq = a.l_seq[i]; pi.read_len = wk.read_len;
    Junction junc[SVIM_MAX_SEGMENTS]; Tandem tand[SVIM_MAX_SEGMENTS];
    if (n >= 2) analyze_chain(chain, n, pi, p, ct, out, wk.ord_sig, wk.ord_twin, err, junc, tand);
    if (err & CH_NO_READLEN) atomicAdd(cnt + CNT_NO_READLEN, 1u);
    if (err & CH_DATA_ERROR) atomicAdd(cnt + CNT_DATA_ERR, 1u);
}

// ---- large-read pass: reads with more than SVIM_MAX_SEGMENTS segments (the reference has no limit, SVIM_inter.py:24-49) --------------
// k_chain_big_caps sizes each listed read's arrays (an SA entry is at least 13 characters: "c,1,+,1M,0,0;"), the host lays them out in
// one scratch buffer, k_segment_chain_big runs the same chain logic as the fast kernels with its arrays in that buffer, one thread
// per read.  Rare by construction; the NO_READLEN skips of the fast attempt are counted here instead (the fast kernel returned early).
__global__ void k_chain_big_caps(DevSoa a, const ChainWork* work, const uint32_t* big_list, uint32_t n_big, const uint32_t* grp, const uint32_t* mem_off,
                                 uint32_t* caps) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_big) return;
    const uint32_t i = work[big_list[k]].aln_idx;
    caps[k] = grp ? (mem_off[grp[i] + 1] - mem_off[grp[i]]) + 2u : a.sa_len[i] / 13u + 3u;
}

struct BigScratch { Seg* seg; Junction* junc; Tandem* tand; const uint64_t* off; };

__global__ void __launch_bounds__(64) k_segment_chain_big(DevSoa a, ChainParams p, ContigTable ct, const ChainWork* work, const uint32_t* big_list, uint32_t n_big,
                                                           const uint32_t* grp, const uint32_t* mem_off, const uint32_t* mem_idx, const SegSum* segsum,
                                                           BigScratch sc, SigQueue qm, SigQueue qt, uint32_t* cnt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_big) return;
    const ChainWork wk = work[big_list[k]];
    const uint32_t i = wk.aln_idx;
    DeviceEmitter out{qm, qt, cnt + CNT_OVERFLOW};
    const int cap = (int)(sc.off[k + 1] - sc.off[k]);
    Seg* chain = sc.seg + sc.off[k]; Junction* junc = sc.junc + sc.off[k]; Tandem* tand = sc.tand + sc.off[k];
    int n = 0;
    uint32_t err = 0;
    const int32_t rev = (a.flag[i] & 0x10u) ? 1 : 0;
    if (rev && wk.read_len < 0) err |= CH_NO_READLEN;
    else { Seg s; s.tid = a.tid[i]; s.ref_start = a.pos[i]; s.ref_end = wk.ref_end; s.q_start = wk.q_start; s.q_end = wk.q_end; s.rev = rev; chain[n++] = s; }
    if (grp) {
        const uint32_t g = grp[i];
        for (uint32_t m = mem_off[g]; m < mem_off[g + 1] && n < cap; ++m) {
            const uint32_t j = mem_idx[m];
            const SegSum ss = segsum[j];
            const int32_t jr = (a.flag[j] & 0x10u) ? 1 : 0;
            if (jr && ss.read_len < 0) { err |= CH_NO_READLEN; continue; }
            Seg s; s.tid = a.tid[j]; s.ref_start = a.pos[j]; s.ref_end = ss.ref_end; s.q_start = ss.q_start; s.q_end = ss.q_end; s.rev = jr;
            chain[n++] = s;
        }
    } else n = parse_sa_segments(a.sa + a.sa_off[i], (int)a.sa_len[i], ct, p, a.l_seq[i], chain, n, err, cap);
    if (err & CH_TOO_MANY) { atomicAdd(cnt + CNT_DATA_ERR, 1u); return; }      // cannot happen: cap bounds the entry count
    sort_chain(chain, n);
    PrimaryInfo pi; pi.aln_idx = i; pi.qname_id = a.qname_id[i]; pi.l_seq = a.l_seq[i]; pi.read_len = wk.read_len;
    if (n >= 2) analyze_chain(chain, n, pi, p, ct, out, wk.ord_sig, wk.ord_twin, err, junc, tand);
    if (err & CH_BAD_FIELDS) atomicAdd(cnt + CNT_BAD_FIELDS, 1u);
    if (err & CH_NO_READLEN) atomicAdd(cnt + CNT_NO_READLEN, 1u);
    if (err & CH_DATA_ERROR) atomicAdd(cnt + CNT_DATA_ERR, 1u);
}

// keys for restoring the reference's emission order: (record index | read group, ordinal)
__global__ void k_sig_keys(const svim_sig* recs, uint32_t n, const uint32_t* grp, uint64_t* keys, uint32_t* vals) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t idx = recs[k].aln_idx;
    const uint32_t hi = (grp && idx != 0xffffffffu) ? grp[idx] : idx;     // holes (aln_idx = ~0) sort to the end
    keys[k] = recs[k].type == 0xff ? ~0ull : (((uint64_t)hi << 32) | recs[k].ordinal);
    vals[k] = k;
}

__global__ void k_sig_gather(const svim_sig* src, const uint32_t* order, uint32_t n, svim_sig* dst, uint64_t* ins_len) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    svim_sig s = src[order[k]];
    dst[k] = s;
    ins_len[k] = (s.type == SVIM_INS) ? s.seq_len : 0;
}

// one warp per record; INS records copy SEQ[src_off : src_off+len) (4-bit) to ASCII
__global__ void __launch_bounds__(256) k_ins_gather(DevSoa a, svim_sig* recs, const uint64_t* blob_off, uint32_t n, uint8_t* blob) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    if (recs[w].type != SVIM_INS) return;
    const uint64_t src = recs[w].seq_off; const uint32_t len = recs[w].seq_len;
    const uint64_t dst = blob_off[w];
    const uint8_t* seq = a.seq + a.seq_off[recs[w].aln_idx];
    for (uint32_t k = lane; k < len; k += 32) {
        uint64_t q = src + k;
        uint8_t b = seq[q >> 1];
        uint8_t code = (q & 1) ? (b & 15) : (b >> 4);
        blob[dst + k] = (uint8_t)"=ACMGRSVTWYHKDBN"[code];
    }
    __syncwarp();
    if (lane == 0) recs[w].seq_off = dst;
}

// lazy-SEQ variant: the packed bases of every insertion were staged contiguously by the host (collect_host path);
// stage_off[w] = byte offset of the record's first packed byte, parity = first nibble inside that byte
__global__ void __launch_bounds__(256) k_ins_gather_staged(const uint8_t* staged, const uint64_t* stage_off, svim_sig* recs, const uint64_t* blob_off,
                                                            uint32_t n, uint8_t* blob) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    if (recs[w].type != SVIM_INS) return;
    const uint64_t par = recs[w].seq_off & 1ull; const uint32_t len = recs[w].seq_len;
    const uint64_t dst = blob_off[w];
    const uint8_t* seq = staged + stage_off[w];
    for (uint32_t k = lane; k < len; k += 32) {
        const uint64_t q = par + k;
        const uint8_t b = seq[q >> 1];
        const uint8_t code = (q & 1) ? (b & 15) : (b >> 4);
        blob[dst + k] = (uint8_t)"=ACMGRSVTWYHKDBN"[code];
    }
    __syncwarp();
    if (lane == 0) recs[w].seq_off = dst;
}

// Host side of the lazy-SEQ path: copy only the packed bytes the gather kernel will read into a pinned staging
// buffer (a few threads: the ranges are scattered over the whole SEQ blob), upload them, run the staged gather.
static int collect_gather_lazy(svimgpu_ctx* ctx, SigSet& set, uint32_t n, const uint64_t* d_ins_off) {
    cudaStream_t st = ctx->stream;
    std::vector<svim_sig>& h = ctx->h_lazy_recs;
    h.resize(n);
    SVIM_CUDA(cudaMemcpyAsync(h.data(), set.recs.p, (size_t)n * sizeof(svim_sig), cudaMemcpyDeviceToHost, st));
    SVIM_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t>& off = ctx->h_lazy_off;
    off.assign(n, 0);
    uint64_t total = 0;
    for (uint32_t k = 0; k < n; ++k) {
        if (h[k].type != SVIM_INS || h[k].seq_len == 0) continue;
        off[k] = total;
        total += ((h[k].seq_off & 1ull) + h[k].seq_len + 1) >> 1;
    }
    if (total + 16 > ctx->h_stage_cap) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
        size_t want = (size_t)total + total / 4 + 4096;
        SVIM_CUDA(cudaMallocHost((void**)&ctx->h_stage, want));
        ctx->h_stage_cap = want;
    }
    uint8_t* stage = ctx->h_stage;
    const uint8_t* hseq = ctx->h_seq; const uint64_t* hoff = ctx->h_seq_off;
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t k = lo; k < hi; ++k) {
            if (h[k].type != SVIM_INS || h[k].seq_len == 0) continue;
            const uint64_t first = h[k].seq_off >> 1;
            const uint64_t nb = ((h[k].seq_off & 1ull) + h[k].seq_len + 1) >> 1;
            memcpy(stage + off[k], hseq + hoff[h[k].aln_idx - ctx->lazy_aln_base] + first, nb);
        }
    };
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 4096 || nt == 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, (uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt));
        for (auto& x : th) x.join();
    }
    SVIM_CUDA(ctx->d_stage.ensure((size_t)total + 16)); SVIM_CUDA(ctx->d_stage_off.ensure((size_t)n * 8));
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_stage.p, stage, (size_t)total, cudaMemcpyHostToDevice, st));
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_stage_off.p, off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    const uint32_t blocks = (uint32_t)(((uint64_t)n * 32 + 255) / 256);
    { ctx->launches++; k_ins_gather_staged<<<blocks, 256, 0, st>>>(ctx->d_stage.as<uint8_t>(), ctx->d_stage_off.as<uint64_t>(), set.recs.as<svim_sig>(), d_ins_off, n,
                                                 set.ins.as<uint8_t>()); }
    SVIM_CUDA(cudaStreamSynchronize(st));   // the host vectors are reused by the next call
    return 0;
}

static int collect_sort_queue(svimgpu_ctx* ctx, int which, uint32_t n_reserved, uint32_t n_holes) {
    // queue[which] (arbitrary order, with holes) -> sets[which].recs in emission order + INS blob
    SigSet& set = ctx->sets[which];
    const uint32_t n_all = n_reserved, n = n_reserved - n_holes;
    set.n = n; set.ins_bytes = 0; set.segmented = false;
    if (n == 0) return 0;
    SVIM_CUDA(set.recs.ensure((size_t)n * sizeof(svim_sig)));
    SVIM_CUDA(ctx->d_keys[0].ensure((size_t)n_all * 8)); SVIM_CUDA(ctx->d_keys[1].ensure((size_t)n_all * 8));
    SVIM_CUDA(ctx->d_vals[0].ensure((size_t)n_all * 4)); SVIM_CUDA(ctx->d_vals[1].ensure((size_t)n_all * 4));
    SVIM_CUDA(ctx->d_scan.ensure((size_t)(n + 1) * 8 * 2));
    cudaStream_t st = ctx->stream;
    const svim_sig* q = ctx->d_queue[which].as<svim_sig>();
    { ctx->launches++; k_sig_keys<<<(n_all + 255) / 256, 256, 0, st>>>(q, n_all, ctx->qs_mode ? ctx->d_qs_grp.as<uint32_t>() : nullptr, ctx->d_keys[0].as<uint64_t>(), ctx->d_vals[0].as<uint32_t>()); }
    cub::DoubleBuffer<uint64_t> dk(ctx->d_keys[0].as<uint64_t>(), ctx->d_keys[1].as<uint64_t>());
    cub::DoubleBuffer<uint32_t> dv(ctx->d_vals[0].as<uint32_t>(), ctx->d_vals[1].as<uint32_t>());
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)n_all, 0, 64, st);
    SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
    SVIM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->d_sort_tmp.p, tmp, dk, dv, (int)n_all, 0, 64, st));
    uint64_t* ins_len = ctx->d_scan.as<uint64_t>();
    uint64_t* ins_off = ins_len + (n + 1);
    { ctx->launches++; k_sig_gather<<<(n + 255) / 256, 256, 0, st>>>(q, dv.Current(), n, set.recs.as<svim_sig>(), ins_len); }
    size_t tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, ins_len, ins_off, (int)n + 1, st);
    SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp2));
    SVIM_CUDA(cudaMemsetAsync(ins_len + n, 0, 8, st));
    SVIM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, tmp2, ins_len, ins_off, (int)n + 1, st));
    uint64_t total = 0;
    SVIM_CUDA(cudaMemcpyAsync(&total, ins_off + n, 8, cudaMemcpyDeviceToHost, st));
    SVIM_CUDA(cudaStreamSynchronize(st));
    set.ins_bytes = (int64_t)total;
    SVIM_CUDA(set.ins.ensure((size_t)total + 16));
    if (total > 0) {
        if (ctx->lazy_seq) { int rc = collect_gather_lazy(ctx, set, n, ins_off); if (rc) return rc; }
        else {
            uint32_t blocks = (uint32_t)(((uint64_t)n * 32 + 255) / 256);
            { ctx->launches++; k_ins_gather<<<blocks, 256, 0, st>>>(ctx->soa, set.recs.as<svim_sig>(), ins_off, n, set.ins.as<uint8_t>()); }
        }
    }
    SVIM_CUDA(cudaGetLastError());
    return 0;
}

static ChainParams make_chain_params(const svim_params& p) {
    ChainParams c;
    c.min_sv = p.min_sv_size; c.max_sv = p.max_sv_size; c.tol_g = p.segment_gap_tolerance; c.tol_o = p.segment_overlap_tolerance;
    c.min_mapq = p.min_mapq; c.all_bnds = p.all_bnds;
    return c;
}

static int collect_run(svimgpu_ctx* ctx, svim_collect_stats* stats) {
    if (!ctx->have_soa) { ctx->set_error(SVIMGPU_ERR_STATE, "no alignments uploaded"); return SVIMGPU_ERR_STATE; }
    if (ctx->n_contigs == 0) { ctx->set_error(SVIMGPU_ERR_STATE, "svimgpu_set_contigs not called"); return SVIMGPU_ERR_STATE; }
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->soa.n;
    if (n >= (int64_t)1 << 32) { ctx->set_error(SVIMGPU_ERR_LIMIT, "more than 2^32 records in one batch"); return SVIMGPU_ERR_LIMIT; }
    // every resident scan warp may sit on up to SCAN_CHUNK-1 reserved-but-unused slots
    uint32_t cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(1 << 16, 2 * n + 1024) + (int64_t)148 * 12 * 8 * SCAN_CHUNK, 0x7fffffff);
    SVIM_CUDA(ctx->d_counters.ensure(CNT_N * 4));
    uint32_t h_cnt[CNT_N];
    ChainParams cp = make_chain_params(ctx->params);
    ContigTable ct{ctx->n_contigs, ctx->d_names.as<char>(), ctx->d_name_off.as<int32_t>(), ctx->d_rank.as<int32_t>()};
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, ctx->device);
    for (int attempt = 0; attempt < 6; ++attempt) {
        SVIM_CUDA(ctx->d_queue[0].ensure((size_t)cap * sizeof(svim_sig)));
        SVIM_CUDA(ctx->d_queue[1].ensure((size_t)(ctx->params.all_bnds ? cap : 16) * sizeof(svim_sig)));
        SVIM_CUDA(ctx->d_work.ensure((size_t)(n + 1) * sizeof(ChainWork)));
        SVIM_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, CNT_N * 4, st));
        SigQueue qm{ctx->d_queue[0].as<svim_sig>(), ctx->d_counters.as<uint32_t>() + CNT_MAIN, cap};
        SigQueue qt{ctx->d_queue[1].as<svim_sig>(), ctx->d_counters.as<uint32_t>() + CNT_TWIN, ctx->params.all_bnds ? cap : 16};
        {
            StageTimer t(ctx, T_SCAN);
            if (n > 0) {
                QsView qsv{nullptr, nullptr, nullptr};
                if (ctx->qs_mode) qsv = QsView{ctx->d_qs_info.as<uint32_t>(), ctx->d_qs_grp.as<uint32_t>(), ctx->d_qs_segsum.as<SegSum>()};
                const int variant = ctx->qs_mode ? 0 : ctx->scan_variant;   // the query-sorted mode lives in the default kernel only
                const uint32_t thresh = cp.min_sv <= 0 ? 0u : (cp.min_sv >= (1 << 28) ? 0xffffffffu : ((uint32_t)cp.min_sv << 4));
#define SCAN_S_LAUNCH(MB, QSF, LDN)                                                                                                     \
    { ctx->launches++; k_cigar_scan_s<MB, QSF, LDN><<<dev_sms * MB * 2, 256, 0, st>>>(ctx->soa, cp, qm, qt, ctx->d_work.as<ChainWork>(), \
                                                       (uint32_t)n + 1, ctx->d_counters.as<uint32_t>(), qsv, ctx->scan_chunks, thresh); }
#define SCAN_LEGACY_LAUNCH(UN, MB)                                                                                                      \
    { ctx->launches++; k_cigar_scan<UN, MB><<<dev_sms * 8, 256, 0, st>>>(ctx->soa, cp, qm, qt, ctx->d_work.as<ChainWork>(), (uint32_t)n + 1, \
                                                                          ctx->d_counters.as<uint32_t>()); }
                if (ctx->qs_mode) SCAN_S_LAUNCH(4, true, 4)                  // query-sorted instantiation of the default kernel
                else if (variant == 0) SCAN_S_LAUNCH(4, false, 4)            // DEFAULT: 4 x 128-bit loads per lane, 4 CTAs/SM (64 registers, no spills in the loop)
                else if (variant == 1) {                                     // experiment: cp.async.bulk ring (TMA engine)
                    const size_t smem = (size_t)SCAN_BULK_WARPS * SCAN_STAGES * SCAN_BLOCK_U4 * 16 + (size_t)SCAN_BULK_WARPS * SCAN_STAGES * 8;
                    SVIM_CUDA(cudaFuncSetAttribute(k_cigar_scan_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    { ctx->launches++; k_cigar_scan_bulk<<<dev_sms * 3, 32 * SCAN_BULK_WARPS, smem, st>>>(ctx->soa, cp, qm, qt, ctx->d_work.as<ChainWork>(), (uint32_t)n + 1,
                                                                  ctx->d_counters.as<uint32_t>()); }
                }
                else if (variant == 4) SCAN_LEGACY_LAUNCH(4, 3)              // the round-1 v2..v5 kernel (80 registers, plain LOP3/IADD classification)
                else if (variant == 7) SCAN_S_LAUNCH(5, false, 4)            // experiments: occupancy x loads in flight
                else if (variant == 8) SCAN_S_LAUNCH(6, false, 4)
                else if (variant == 11) SCAN_S_LAUNCH(6, false, 2)
                else if (variant == 12) SCAN_S_LAUNCH(8, false, 2)
                else if (variant == 13) SCAN_S_LAUNCH(5, false, 3)
                else if (variant == 14) SCAN_S_LAUNCH(6, false, 3)
                else if (variant == 15) SCAN_S_LAUNCH(3, false, 4)
                else SCAN_LEGACY_LAUNCH(4, 4)                                // variant 10: the legacy kernel at 4 CTAs/SM
#undef SCAN_S_LAUNCH
#undef SCAN_LEGACY_LAUNCH
            }
        }
        SVIM_CUDA(cudaGetLastError());
        SVIM_CUDA(cudaMemcpyAsync(h_cnt, ctx->d_counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaStreamSynchronize(st));
        uint32_t n_work = h_cnt[CNT_WORK];
        {
            StageTimer t(ctx, T_CHAIN);
            SVIM_CUDA(ctx->d_big_list.ensure((size_t)(n_work + 1) * 4));
            if (n_work > 0 && ctx->qs_mode)
                { ctx->launches++; k_segment_chain_qs<<<(n_work + 127) / 128, 128, 0, st>>>(ctx->soa, cp, ct, ctx->d_work.as<ChainWork>(), n_work,
                                                                        ctx->d_qs_grp.as<uint32_t>(), ctx->d_qs_mem_off.as<uint32_t>(), ctx->d_qs_mem_idx.as<uint32_t>(),
                                                                        ctx->d_qs_segsum.as<SegSum>(), qm, qt, ctx->d_counters.as<uint32_t>(), ctx->d_big_list.as<uint32_t>()); }
            else if (n_work > 0)
                { ctx->launches++; k_segment_chain<<<(n_work + 127) / 128, 128, 0, st>>>(ctx->soa, cp, ct, ctx->d_work.as<ChainWork>(), n_work, qm, qt,
                                                                     ctx->d_counters.as<uint32_t>(), ctx->d_big_list.as<uint32_t>()); }
        }
        SVIM_CUDA(cudaGetLastError());
        SVIM_CUDA(cudaMemcpyAsync(h_cnt, ctx->d_counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaStreamSynchronize(st));
        if (const uint32_t n_big = h_cnt[CNT_TOO_MANY]) {          // reads with more segments than the fast path's arrays hold
            StageTimer t(ctx, T_CHAIN);
            const uint32_t* grp = ctx->qs_mode ? ctx->d_qs_grp.as<uint32_t>() : nullptr;
            SVIM_CUDA(ctx->d_big_caps.ensure((size_t)(n_big + 2) * 12));
            uint32_t* d_caps = ctx->d_big_caps.as<uint32_t>();
            { ctx->launches++; k_chain_big_caps<<<(n_big + 127) / 128, 128, 0, st>>>(ctx->soa, ctx->d_work.as<ChainWork>(), ctx->d_big_list.as<uint32_t>(), n_big, grp,
                                                                     ctx->d_qs_mem_off.as<uint32_t>(), d_caps); }
            std::vector<uint32_t> caps(n_big); std::vector<uint64_t> off(n_big + 1, 0);
            SVIM_CUDA(cudaMemcpyAsync(caps.data(), d_caps, (size_t)n_big * 4, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            for (uint32_t k = 0; k < n_big; ++k) off[k + 1] = off[k] + caps[k];
            uint64_t* d_off = (uint64_t*)(d_caps + ((n_big + 2) & ~1u));
            SVIM_CUDA(cudaMemcpyAsync(d_off, off.data(), (size_t)(n_big + 1) * 8, cudaMemcpyHostToDevice, st));
            const size_t tot = (size_t)off[n_big];
            SVIM_CUDA(ctx->d_big_scratch.ensure(tot * (sizeof(Seg) + sizeof(Junction) + sizeof(Tandem)) + 64));
            BigScratch sc; sc.seg = ctx->d_big_scratch.as<Seg>(); sc.junc = (Junction*)(sc.seg + tot); sc.tand = (Tandem*)(sc.junc + tot); sc.off = d_off;
            { ctx->launches++; k_segment_chain_big<<<(n_big + 63) / 64, 64, 0, st>>>(ctx->soa, cp, ct, ctx->d_work.as<ChainWork>(), ctx->d_big_list.as<uint32_t>(), n_big, grp,
                                                                      ctx->d_qs_mem_off.as<uint32_t>(), ctx->d_qs_mem_idx.as<uint32_t>(), ctx->d_qs_segsum.as<SegSum>(), sc, qm, qt,
                                                                      ctx->d_counters.as<uint32_t>()); }
            SVIM_CUDA(cudaGetLastError());
            SVIM_CUDA(cudaMemcpyAsync(h_cnt, ctx->d_counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
        }
        if (!h_cnt[CNT_OVERFLOW]) break;
        if (attempt == 5 || cap == 0x7fffffff) { ctx->set_error(SVIMGPU_ERR_LIMIT, "signature queue overflow"); return SVIMGPU_ERR_LIMIT; }
        cap = (uint32_t)std::min<uint64_t>((uint64_t)std::max(h_cnt[CNT_MAIN], h_cnt[CNT_TWIN]) + 1024, 0x7fffffffull);
    }
    {
        StageTimer t(ctx, T_SORTBACK);
        int rc = collect_sort_queue(ctx, 0, h_cnt[CNT_MAIN], h_cnt[CNT_HOLES_MAIN]); if (rc) return rc;
        rc = collect_sort_queue(ctx, 1, h_cnt[CNT_TWIN], h_cnt[CNT_HOLES_TWIN]); if (rc) return rc;
    }
    SVIM_CUDA(cudaStreamSynchronize(st));
    svim_collect_stats& s = ctx->cstats;
    s.n_signatures = ctx->sets[0].n; s.n_twin_signatures = ctx->sets[1].n;
    s.ins_bytes = ctx->sets[0].ins_bytes; s.twin_ins_bytes = ctx->sets[1].ins_bytes;
    s.n_sa_bad_fields = h_cnt[CNT_BAD_FIELDS]; s.n_no_read_length = h_cnt[CNT_NO_READLEN];
    s.n_primaries = h_cnt[CNT_PRIMARIES]; s.n_data_errors = h_cnt[CNT_DATA_ERR];
    ctx->collected = true;
    if (stats) *stats = s;
    return 0;
}

// Word-level pieces of the bit-parallel edit distance (Myers / Hyyro) and the banded wavefront built on them.
// Everything here is SVIM_HD: tests/hostcheck compiles it with g++ and replays the G-lane wavefront on the host,
// the kernels in myers.cu run the same functions one lane per thread.
//
// Banding (Ukkonen): with m >= n and a distance bound k >= m-n, an alignment of cost <= k never leaves the
// diagonals  -a <= row-col <= b,  a = (k-(m-n))/2,  b = (m-n)+a.  A DP restricted to that band computes upper
// bounds everywhere and exact values along every path of cost <= k, so   result <= k  =>  result is exact,
// result > k  =>  the true distance is > k  (the pair is recomputed without a band).  edlib, which the reference
// calls (SVIM_clustering.py:45), does the same with block granularity and band doubling.
//
// Wavefront layout: the pattern is cut into word groups of WPL 64-row words; group q runs text column c at step
// c+q, so a group reads its upstream neighbour's horizontal delta of the same column one step later.  A pair is
// owned by G lanes and lane l runs groups l, l+G, l+2G, ...: when the band has left a group the lane reloads the
// next one (rotating window).  That is legal as long as  a + b + 64*WPL <= G*(64*WPL + 1)  (myers_band_fits):
// the lane count follows the band width, not the pattern length.
//   * The pattern is padded with virtual rows at the TOP (Eq = 0, vertical delta 0, horizontal delta +1: they carry
//     the boundary D[0][j] = j), so the last pattern row is bit 63 of the last word of the last group.
//   * A group that enters the band at column cs > 0 starts from Pv = ~0 (every cell one more than the cell above),
//     an upper bound; a group whose upstream neighbour has left the band takes horizontal delta +1 at its top.
//   * distance = m + sum over groups q of the bottom-row horizontal deltas of columns [cs_q, cs_{q+1}).
#pragma once
#include "common.cuh"

#ifdef SVIM_HOST_ONLY
static inline uint32_t mb_imad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
static inline uint32_t mb_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline uint32_t mb_funnel_l1(uint32_t lo, uint32_t hi) { return (hi << 1) | (lo >> 31); }
static inline void mb_add64(uint32_t al, uint32_t bl, uint32_t ah, uint32_t bh, uint32_t& sl, uint32_t& sh) {
    const uint64_t s = (((uint64_t)ah << 32) | al) + (((uint64_t)bh << 32) | bl);
    sl = (uint32_t)s; sh = (uint32_t)(s >> 32);
}
static inline int mb_popcll(uint64_t x) { return __builtin_popcountll(x); }
static inline int mb_popc(uint32_t x) { return __builtin_popcount(x); }
template <int LUT> static inline uint32_t mb_lop3(uint32_t a, uint32_t b, uint32_t c) {      // truth table: a = 0xF0, b = 0xCC, c = 0xAA
    uint32_t r = 0;
    if (LUT & 0x80) r |= a & b & c;
    if (LUT & 0x40) r |= a & b & ~c;
    if (LUT & 0x20) r |= a & ~b & c;
    if (LUT & 0x10) r |= a & ~b & ~c;
    if (LUT & 0x08) r |= ~a & b & c;
    if (LUT & 0x04) r |= ~a & b & ~c;
    if (LUT & 0x02) r |= ~a & ~b & c;
    if (LUT & 0x01) r |= ~a & ~b & ~c;
    return r;
}
static inline const uint8_t* mb_ptr_inc(const uint8_t* p, uint32_t) { return p + 1; }
#else
__device__ __forceinline__ uint32_t mb_imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t mb_umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ uint32_t mb_funnel_l1(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 1); }
__device__ __forceinline__ void mb_add64(uint32_t al, uint32_t bl, uint32_t ah, uint32_t bh, uint32_t& sl, uint32_t& sh) {
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=r"(sl), "=r"(sh) : "r"(al), "r"(bl), "r"(ah), "r"(bh));
}
__device__ __forceinline__ int mb_popcll(uint64_t x) { return __popcll(x); }
__device__ __forceinline__ int mb_popc(uint32_t x) { return __popc(x); }
template <int LUT> __device__ __forceinline__ uint32_t mb_lop3(uint32_t a, uint32_t b, uint32_t c) {   // one LOP3, as written
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
// p + 1 as IMAD.WIDE (FMA pipe; `one` is a kernel argument the compiler cannot fold)
__device__ __forceinline__ const uint8_t* mb_ptr_inc(const uint8_t* p, uint32_t one) {
    unsigned long long q = (unsigned long long)p;
    asm("mad.wide.u32 %0, %1, %1, %0;" : "+l"(q) : "r"(one));
    return (const uint8_t*)q;
}
#endif

// ---- pair bins: a group of G lanes owns one pair with WPL words per lane ------------------------------------------
#define MYERS_BINS 10

struct MyersBin { int G, WPL, capW; };
SVIM_HD MyersBin myers_bin_spec(int b) {
    switch (b) {
        case 0: return {4, 1, 4};   case 1: return {4, 2, 8};   case 2: return {4, 3, 12};  case 3: return {4, 4, 16};
        case 4: return {8, 3, 24};  case 5: return {8, 4, 32};  case 6: return {16, 3, 48}; case 7: return {16, 4, 64};
        case 8: return {32, 3, 96}; default: return {32, 4, 1 << 30};   // unbanded bin 9 strip-mines beyond 128 words
    }
}
SVIM_HD int myers_bin_of(int64_t m) {
    const int64_t W = (m + 63) >> 6;
    return W <= 4 ? 0 : W <= 8 ? 1 : W <= 12 ? 2 : W <= 16 ? 3 : W <= 24 ? 4 : W <= 32 ? 5 : W <= 48 ? 6 : W <= 64 ? 7 : W <= 96 ? 8 : 9;
}

// distance bound of the banded pass: k = m*num/1024 + add; -1 = no banded pass (policy off, or k < m-n <= distance)
SVIM_HD int64_t myers_band_k(int64_t m, int64_t n, int32_t num, int32_t add) {
    if (num <= 0) return -1;
    const int64_t k = ((m * (int64_t)num) >> 10) + add;
    return k < m - n ? -1 : k;
}
SVIM_HD bool myers_band_fits(int64_t m, int64_t n, int64_t k, int G, int WPL) {
    const int64_t a = (k - (m - n)) / 2, b = (m - n) + a;
    return a + b + 64 * WPL <= (int64_t)G * (64 * WPL + 1);
}
// smallest shape whose rotating window holds the band, if it is smaller than the shape the whole pattern needs; else -1
SVIM_HD int myers_band_bin(int64_t m, int64_t n, int32_t num, int32_t add) {
    const int64_t k = myers_band_k(m, n, num, add);
    if (k < 0 || n <= 0) return -1;
    const int full = myers_bin_of(m);
    for (int b = 0; b < full; ++b) {
        const MyersBin s = myers_bin_spec(b);
        if (myers_band_fits(m, n, k, s.G, s.WPL)) return b;
    }
    return -1;
}

// ---- three-plane word in 32-bit halves (ALU-pipe formulation) -------------------------------------------------------
struct Word32 { uint32_t p0l, p0h, p1l, p1h, p2l, p2h, pvl, pvh, mvl, mvh; };

// one word step; e = hin + 1 in {0,1,2}; returns hout + 1
template <int NP>
SVIM_D uint32_t word_step(Word32& w, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t e) {
    const uint32_t hneg = 1u >> e, hpos = e >> 1;
    const uint32_t eql = NP == 2 ? ~((w.p0l ^ m0) | (w.p1l ^ m1)) : ~((w.p0l ^ m0) | (w.p1l ^ m1) | (w.p2l ^ m2));
    const uint32_t eqh = NP == 2 ? ~((w.p0h ^ m0) | (w.p1h ^ m1)) : ~((w.p0h ^ m0) | (w.p1h ^ m1) | (w.p2h ^ m2));
    const uint32_t xvl = eql | w.mvl, xvh = eqh | w.mvh;
    const uint32_t el = eql | hneg;
    uint32_t sl, sh;
    mb_add64(el & w.pvl, w.pvl, eqh & w.pvh, w.pvh, sl, sh);
    const uint32_t xhl = (sl ^ w.pvl) | el, xhh = (sh ^ w.pvh) | eqh;
    const uint32_t phl = w.mvl | ~(xhl | w.pvl), phh = w.mvh | ~(xhh | w.pvh);
    const uint32_t mhl = w.pvl & xhl, mhh = w.pvh & xhh;
    const uint32_t eout = 1u + (phh >> 31) - (mhh >> 31);
    const uint32_t phl2 = (phl << 1) | hpos, phh2 = mb_funnel_l1(phl, phh);
    const uint32_t mhl2 = (mhl << 1) | hneg, mhh2 = mb_funnel_l1(mhl, mhh);
    w.pvl = mhl2 | ~(xvl | phl2); w.pvh = mhh2 | ~(xvh | phh2);
    w.mvl = phl2 & xvl; w.mvh = phh2 & xvh;
    return eout;
}

// ---- two-plane (A/C/G/T) words: FMA-pipe formulation ----------------------------------------------------------------
// The ALU pipe (LOP3/SHF/IADD3, one warp instruction every two cycles per SM sub-partition) bounds the plain
// formulation while the FMA pipe idles.  Here every 64-row word is two independent 32-row blocks chained through
// their horizontal deltas, and whatever can be phrased as a multiply-add runs as IMAD on the FMA pipe:
//   Eq      = P0 + b0*(P1-P0) + b1*(P2-P0) + b0b1*(P3-P2-P1+P0)   (b1 b0 = text symbol code; 3 IMAD, no LOP3)
//   s       = (Eq|hn) & Pv + Pv                               (IMAD with multiplicand `one`)
//   Ph<<1|hp = Ph*two + hp,  Mh<<1|hn = Mh*two + hn           (IMAD; bit 0 is free after the shift)
// leaving 10 ALU-pipe and 6 FMA-pipe instructions per 32 cells instead of ~16 ALU.
struct WordQ { uint32_t q0[2], d1[2], d2[2], d3[2], pv[2], mv[2]; };

// from the two code bit-planes of 64 rows; `real` = rows that exist (a missing row matches nothing)
SVIM_D void wordq_from_planes(WordQ& w, uint64_t pl0, uint64_t pl1, uint64_t real, uint64_t pv0) {
    const uint64_t P0 = ~pl0 & ~pl1 & real, P1 = pl0 & ~pl1 & real, P2 = ~pl0 & pl1 & real, P3 = pl0 & pl1 & real;
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
        const uint32_t p0 = (uint32_t)(P0 >> (32 * h));
        const uint32_t p1 = (uint32_t)(P1 >> (32 * h)), p2 = (uint32_t)(P2 >> (32 * h)), p3 = (uint32_t)(P3 >> (32 * h));
        w.q0[h] = p0; w.d1[h] = p1 - p0; w.d2[h] = p2 - p0; w.d3[h] = p3 - p2 - p1 + p0;
        w.pv[h] = (uint32_t)(pv0 >> (32 * h)); w.mv[h] = 0u;
    }
}

// one 32-row block step; hp/hn = incoming horizontal delta (+1 / -1 flags), replaced by the outgoing one
template <bool HI>
SVIM_D void block_step(uint32_t q0, uint32_t d1, uint32_t d2, uint32_t d3, uint32_t& pv_io, uint32_t& mv_io,
                        uint32_t c1, uint32_t c2, uint32_t c3, uint32_t one, uint32_t two, uint32_t& hp, uint32_t& hn) {
    const uint32_t pv = pv_io, mv = mv_io;
    const uint32_t eq = mb_imad(c3, d3, mb_imad(c2, d2, mb_imad(c1, d1, q0)));
    const uint32_t xv = eq | mv;
    const uint32_t el = eq | hn;
    const uint32_t s = mb_imad(el & pv, one, pv);
    const uint32_t xh = (s ^ pv) | el;
    const uint32_t ph = mv | ~(xh | pv);
    const uint32_t mh = pv & xh;
    const uint32_t ph2 = mb_imad(ph, two, hp), mh2 = mb_imad(mh, two, hn);
    if (HI) { hp = mb_umulhi(ph, two); hn = mb_umulhi(mh, two); }   // top bit via IMAD.HI (FMA pipe)
    else { hp = ph >> 31; hn = mh >> 31; }
    pv_io = mh2 | ~(xv | ph2);
    mv_io = ph2 & xv;
}

// ---- banded wavefront ---------------------------------------------------------------------------------------------
struct BandGeom {
    int32_t n;      // text columns
    int32_t NG;     // word groups of the padded pattern
    int32_t pad;    // virtual rows above row 0 (the pattern ends on a group boundary)
    int32_t a, b;   // band: -a <= row - col <= b
};

SVIM_HD BandGeom band_geom(int64_t m, int64_t n, int64_t k, int WPL) {
    BandGeom ge;
    ge.n = (int32_t)n;
    ge.NG = (int32_t)((m + 64 * WPL - 1) / (64 * WPL));
    ge.pad = (int32_t)((int64_t)ge.NG * 64 * WPL - m);
    ge.a = (int32_t)((k - (m - n)) / 2);
    ge.b = (int32_t)((m - n) + ge.a);
    return ge;
}

// word w of the padded pattern (bit t <-> row 64w + t - pad): out[0..1] = bit-planes of the 2-bit symbol codes, out[2] = rows that exist
SVIM_HD void band_build_word(const uint8_t* pat, int64_t m, int32_t pad, int64_t w, uint64_t* out) {
    uint64_t p0 = 0, p1 = 0, real = 0;
    const int64_t r0 = 64 * w - pad;
    for (int t = 0; t < 64; ++t) {
        const int64_t r = r0 + t;
        if (r >= 0 && r < m) {
            const uint64_t code = pat[r];
            p0 |= (code & 1ull) << t; p1 |= ((code >> 1) & 1ull) << t; real |= 1ull << t;
        }
    }
    out[0] = p0; out[1] = p1; out[2] = real;
}

template <int WPL>
struct BandLane {
    WordQ w[WPL];
    int32_t g;            // word group this lane is on
    int32_t s_lo, s_hi;   // wavefront steps on which that group is inside the band (group q runs column c at step c + q)
    int32_t s_top;        // steps >= s_top have no upstream neighbour inside the band
    int32_t s_acc;        // the group's bottom-row deltas count on steps < s_acc
    int32_t score;
    const uint8_t* tp;    // text symbol code of this step's column
};

template <int WPL>
SVIM_D void band_load_group(BandLane<WPL>& L, const BandGeom& ge, const uint64_t* planes) {
    const int32_t base = 64 * WPL * L.g - ge.pad;                 // pattern row of the group's first bit
    const int32_t lo = base - ge.b, hi = base + 64 * WPL - 1 + ge.a;
    L.s_lo = (lo > 0 ? lo : 0) + L.g;
    L.s_hi = (hi < ge.n - 1 ? hi : ge.n - 1) + L.g;
    L.s_top = L.g == 0 ? INT32_MIN : base + ge.a + L.g;
    const int32_t nxt = base + 64 * WPL - ge.b;
    L.s_acc = ((L.g + 1 < ge.NG) ? (nxt > 0 ? nxt : 0) : ge.n) + L.g;
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
    for (int k = 0; k < WPL; ++k) {
        const uint64_t* p = planes + 3 * ((int64_t)L.g * WPL + k);
        wordq_from_planes(L.w[k], p[0], p[1], p[2], p[2]);
    }
}

template <int WPL>
SVIM_D void band_lane_init(BandLane<WPL>& L, const BandGeom& ge, const uint64_t* planes, const uint8_t* txt, int gl, bool valid) {
    L.g = valid ? gl : ge.NG; L.score = 0; L.s_lo = L.s_hi = INT32_MAX; L.s_top = 0; L.s_acc = 0;
    L.tp = txt - gl;                   // column of step s = s - g; never read outside [s_lo, s_hi]
    if (L.g < ge.NG) band_load_group(L, ge, planes);
    else {
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
        for (int k = 0; k < WPL; ++k)
            for (int h = 0; h < 2; ++h) { L.w[k].q0[h] = L.w[k].d1[h] = L.w[k].d2[h] = L.w[k].d3[h] = 0; L.w[k].pv[h] = L.w[k].mv[h] = 0; }
    }
}

// One wavefront step of one lane.  recv = the upstream lane's packet of the previous step, e_out = this lane's
// packet (horizontal delta + 1 at the bottom of its group), kept when the lane idles.  The kernel is bound by the
// ALU pipe, so the bookkeeping stays small: windows are kept in step space (two compares decide rotate / active),
// the text pointer advances on the FMA pipe.
template <int G, int WPL, bool HI>
SVIM_D uint32_t band_lane_step(BandLane<WPL>& L, const BandGeom& ge, const uint64_t* planes, int32_t s, uint32_t recv,
                                uint32_t e_out, uint32_t one, uint32_t two) {
    if (s > L.s_hi) {                  // the band has left this group: rotate to the lane's next one (s_hi = INT32_MAX once the lane is done)
        L.g += G; L.tp -= G;
        if (L.g < ge.NG) band_load_group(L, ge, planes);
        else L.s_lo = L.s_hi = INT32_MAX;
    }
    if (s >= L.s_lo) {
        const uint32_t code = *L.tp;
        const uint32_t c1 = code & 1u, c2 = code >> 1, c3 = c1 & c2;      // b0, b1, b0b1 (codes 0..3)
        const uint32_t e = s >= L.s_top ? 2u : recv;
        uint32_t hp = e >> 1, hn = 1u >> e;
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
        for (int k = 0; k < WPL; ++k) {
            block_step<HI>(L.w[k].q0[0], L.w[k].d1[0], L.w[k].d2[0], L.w[k].d3[0], L.w[k].pv[0], L.w[k].mv[0], c1, c2, c3, one, two, hp, hn);
            block_step<HI>(L.w[k].q0[1], L.w[k].d1[1], L.w[k].d2[1], L.w[k].d3[1], L.w[k].pv[1], L.w[k].mv[1], c1, c2, c3, one, two, hp, hn);
        }
        e_out = 1u + hp - hn;
        if (s < L.s_acc) L.score += (int32_t)hp - (int32_t)hn;
    }
    L.tp = mb_ptr_inc(L.tp, one);
    return e_out;
}

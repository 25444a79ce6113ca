// Thread-per-pair banded edit distance ("TPP"): the time-dominant kernel of CLUSTER (compute_haplotype_edit_distance +
// edlib.align, SVIM_clustering.py:32-45).
//
// The G-lane wavefront of myers_band.cuh spends ~1/4 of its ALU-pipe slots on lane bookkeeping, idles lanes while the
// wavefront fills and drains, and rounds the band up to 64*WPL rows.  Here ONE THREAD owns one pair:
//   * the Ukkonen band  -a <= row-col <= b  is a sliding window of B 32-row blocks whose vertical deltas (Pv, Mv) live in
//     2*B registers; a column is B dependent block steps in program order, no shuffles, no per-step lane logic;
//   * the match masks Eq[block][symbol] of the window sit in shared memory, laid out [block][symbol][lane] so the 32 lanes
//     of a warp (32 different pairs, different symbols) never conflict: one LDS per block step on the otherwise idle
//     load/store pipe replaces the three IMADs of the bilinear Eq of myers_band.cuh and frees four registers per block;
//   * the window slides down one block every 32 columns.  Every thread delays its first column by phase = (-a) mod 32, so
//     all 32 pairs of a warp slide at the same loop iteration (chunk boundaries of the shifted column index s = j+phase);
//     rows above row 0 are virtual (Eq = 0, vertical delta 0: they carry the boundary D[0][j] = j downwards), rows past
//     the pattern are virtual too (Eq = 0) and are subtracted again at the end;
//   * the distance is tracked along the bottom row of the window: + its horizontal delta per column (the carry that
//     leaves the last block), + 32 per slide (a block enters with all vertical deltas +1);
//   * result <= k  =>  exact;  result > k  =>  the true distance is > k and the pair is recomputed without a band
//     (what edlib does with band doubling).  Patterns that fit B blocks whole run unbanded in the same code (a < 0).
// Block step: Hyyro's D0 form, 7 LOP3 + 2 SHF on the ALU pipe, 3 IMAD on the FMA pipe (carry add and both shifts), 1 LDS.
//
// Everything above the kernel is SVIM_HD: tests/hostcheck replays tpp_thread on the host against a plain DP.
#pragma once
#include "myers_band.cuh"

#ifndef TPP_AHEAD
#define TPP_AHEAD 4      // block steps between the load of a match mask and its use
#endif
#define TPP_BUCKETS 18
#define TPP_MAX_B 28
SVIM_HD int tpp_bucket_B(int q) {
    switch (q) {
        case 0: return 2;  case 1: return 3;  case 2: return 4;  case 3: return 5;  case 4: return 6;  case 5: return 7;
        case 6: return 8;  case 7: return 9;  case 8: return 10; case 9: return 12; case 10: return 14; case 11: return 16;
        case 12: return 18; case 13: return 20; case 14: return 22; case 15: return 24; case 16: return 26; default: return 28;
    }
}
SVIM_HD int tpp_bucket_of(int B) {   // smallest bucket that holds B blocks, -1 if none
    if (B > TPP_MAX_B) return -1;
    if (B <= 2) return 0;
    if (B <= 10) return B - 2;
    return 8 + (B - 10 + 1) / 2;
}

// window blocks of the banded run: 32*B >= a + b + 32 (the window top is the band top rounded down to a block)
SVIM_HD int64_t tpp_blocks_band(int64_t m, int64_t n, int64_t k) {
    const int64_t a = (k - (m - n)) / 2, b = (m - n) + a;
    return (a + b + 63) / 32;
}
SVIM_HD int64_t tpp_blocks_full(int64_t m) { return (m + 31) / 32; }

struct TppPlan { int32_t B; int32_t a; };   // a < 0: unbanded
// how the pair runs under the band policy (num, add): banded when that needs fewer blocks than the whole pattern
SVIM_HD TppPlan tpp_plan(int64_t m, int64_t n, int32_t num, int32_t add) {
    TppPlan p; p.a = -1;
    int64_t B = tpp_blocks_full(m);
    const int64_t k = myers_band_k(m, n, num, add);
    if (k >= 0 && n > 0) {
        const int64_t Bb = tpp_blocks_band(m, n, k);
        if (Bb < B) { B = Bb; p.a = (int32_t)((k - (m - n)) / 2); }
    }
    p.B = B > 0x7fffffff ? 0x7fffffff : (int32_t)B;
    return p;
}

// one 32-row block step, D0 form.  hp/hn: incoming horizontal delta flags (+1 / -1) at the block's top, replaced by the
// outgoing ones at its bottom.  `one`, `two` are opaque 1 and 2 (multiplicands that keep add and shifts on the FMA pipe).
SVIM_D void tpp_block(uint32_t eq, uint32_t& vp_io, uint32_t& vn_io, uint32_t& hp, uint32_t& hn, uint32_t one, uint32_t two) {
    const uint32_t vp = vp_io, vn = vn_io;
    const uint32_t t = mb_lop3<0xA8>(eq, hn, vp);            // (eq | hn) & vp
    const uint32_t x = mb_lop3<0xFE>(eq, hn, vn);            // eq | hn | vn          (off the carry chain)
    const uint32_t s = mb_imad(t, one, vp);
    const uint32_t d0 = mb_lop3<0xBE>(s, vp, x);             // (s ^ vp) | x
    const uint32_t ph = mb_lop3<0xF1>(vn, d0, vp);           // vn | ~(d0 | vp)
    const uint32_t mh = d0 & vp;
    const uint32_t ph2 = mb_imad(ph, two, hp), mh2 = mb_imad(mh, two, hn);
    hp = ph >> 31; hn = mh >> 31;                          // (as IMAD.HI on the FMA pipe this measured slower: 25.0 vs 23.2 ms on config2)
    vp_io = mb_lop3<0xF1>(mh2, d0, ph2);                      // mh2 | ~(d0 | ph2)
    vn_io = d0 & ph2;
}

// bits of a block whose absolute row (row0 + bit) is >= lim
SVIM_D uint32_t tpp_mask_from(int32_t row0, int32_t lim) {
    const int32_t d = lim - row0;
    return d <= 0 ? 0xffffffffu : (d >= 32 ? 0u : (0xffffffffu << d));
}

template <int B, class Eq>
SVIM_D void tpp_load_column(const Eq& eq, uint32_t sym, uint32_t (&E)[B]) {
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
    for (int i = 0; i < B; ++i) E[i] = eq.get(i, sym);
}

// One pair, one thread.  Eq: window store (put / get / shift_up), Peq: match masks of pattern block b (zeros outside the
// pattern), Txt: text symbols at shifted column index s = j + phase, pre-scaled for Eq::get (byte(s), word(s) = 4 bytes).
template <int B, class Eq, class Peq, class Txt>
SVIM_D int32_t tpp_thread(int32_t m, int32_t n, int32_t a, Eq& eq, const Peq& peq, const Txt& txt, uint32_t one, uint32_t two) {
    const bool banded = a >= 0;
    const int32_t phase = banded ? ((-a) & 31) : 0;
    const int32_t A = banded ? a + phase : 0;            // multiple of 32: the window top of chunk c is row 32c - A
    uint32_t vp[B], vn[B];
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
    for (int i = 0; i < B; ++i) {
        vp[i] = tpp_mask_from(32 * i - A, 0); vn[i] = 0u;
        uint32_t v[4]; peq.block(i - (A >> 5), v); eq.put(i, v);
    }
    int32_t score = 32 * B - A;
    const int32_t s_end = phase + n;
    // global loads are issued one unit ahead of their use: the masks of the block that enters at the next slide, and the next word
    // of text symbols (the scratch is padded, reading one word past the text is harmless)
    uint32_t v_next[4]; peq.block(1 - (A >> 5) + B - 1, v_next);
    uint32_t w_next = txt.word((phase + 3) & ~3);
    for (int32_t c = 0; 32 * c < s_end; ++c) {
        if (c > 0 && banded) {                            // slide: the top block leaves, a block of fresh rows enters
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
            for (int i = 0; i + 1 < B; ++i) { vp[i] = vp[i + 1]; vn[i] = vn[i + 1]; }
            vp[B - 1] = 0xffffffffu; vn[B - 1] = 0u;
            eq.shift_up();
            eq.put(B - 1, v_next);
            peq.block(c + 1 - (A >> 5) + B - 1, v_next);
            score += 32;
        }
        int32_t s = 32 * c > phase ? 32 * c : phase;
        const int32_t s_hi = 32 * c + 32 < s_end ? 32 * c + 32 : s_end;
        while (s < s_hi) {
            if ((s & 3) == 0 && s + 4 <= s_hi) {
                // four columns from one text word = 4*B block steps in program order.  The match masks are loaded TPP_AHEAD block
                // steps before their use (a small ring of registers), so the shared-memory latency stays off the carry chain.
                const uint32_t w = w_next;
                w_next = txt.word(s + 4);
                uint32_t ring[TPP_AHEAD];
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
                for (int t = 0; t < TPP_AHEAD; ++t) ring[t] = eq.get(t % B, (w >> (8 * ((t / B) & 3))) & 0xffu);
                uint32_t hp = 1u, hn = 0u;
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
                for (int t = 0; t < 4 * B; ++t) {
                    const uint32_t e = ring[t % TPP_AHEAD];
                    if (t + TPP_AHEAD < 4 * B) ring[t % TPP_AHEAD] = eq.get((t + TPP_AHEAD) % B, (w >> (8 * ((t + TPP_AHEAD) / B))) & 0xffu);
                    tpp_block(e, vp[t % B], vn[t % B], hp, hn, one, two);
                    if (t % B == B - 1) { score += (int32_t)hp - (int32_t)hn; hp = 1u; hn = 0u; }
                }
                s += 4;
            } else {
                const uint32_t sym = txt.byte(s);
                uint32_t hp = 1u, hn = 0u;
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
                for (int i = 0; i < B; ++i) tpp_block(eq.get(i, sym), vp[i], vn[i], hp, hn, one, two);
                score += (int32_t)hp - (int32_t)hn;
                ++s;
            }
        }
    }
    // the tracked value is D at the window's bottom row; walk the vertical deltas back up to the pattern's last row
    const int32_t cl = s_end > 0 ? (s_end - 1) >> 5 : 0;
    const int32_t top = banded ? 32 * cl - A : 0;
#ifndef SVIM_HOST_ONLY
#pragma unroll
#endif
    for (int i = 0; i < B; ++i) {
        const uint32_t below = tpp_mask_from(top + 32 * i, m);
        score -= mb_popc(vp[i] & below) - mb_popc(vn[i] & below);
    }
    return score;
}

// match masks of 32 pattern rows from their symbol codes (0..3); rows >= cnt do not exist
SVIM_HD void tpp_masks_from_codes(const uint8_t* codes, int cnt, uint32_t v[4]) {
    v[0] = v[1] = v[2] = v[3] = 0u;
    for (int r = 0; r < cnt; ++r) v[codes[r] & 3] |= 1u << r;
}

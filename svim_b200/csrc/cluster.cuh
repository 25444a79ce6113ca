// Scalar pieces of CLUSTER shared by the kernels (and tests/hostcheck):
// span_position_distance (SVIM_clustering.py:47-96), the dendrogram post-processing of
// scipy linkage/fcluster (call sites :170-171), and cluster consolidation (:183-303).
// All floating point is FP64 in the reference's operation order; compile with
// -fmad=false so no contraction changes a rounding.
#pragma once
#include "common.cuh"

struct ClusterParams {
    double partition_max_distance, pos_norm, edit_norm, cluster_max_distance;
};

// condensed index of (i,j), i<j, n points — scipy condensed_index
SVIM_HD int cidx(int n, int i, int j) { return n * i - (i * (i + 1)) / 2 + (j - i - 1); }
SVIM_HD int cidx_any(int n, int i, int j) { return i < j ? cidx(n, i, j) : cidx(n, j, i); }

// Fields of one signature as the distance function reads them.
struct SigView {
    double start, end, dpos;
    uint32_t read_id;
    uint8_t dirs;
};

// span_position_distance; `ed` = haplotype edit distance (INS, only read when the gate passes).
// *err set when the reference would raise ZeroDivisionError.
SVIM_HD double spd(int type, const SigView& a, const SigView& b, const ClusterParams& p, double ed, int* err) {
    if (type == SVIM_BND) {                                           // :87-94
        double d1 = fabs(a.start - b.start), d2 = fabs(a.dpos - b.dpos);
        if (a.dirs == b.dirs) return (d1 + d2) / 3000.0;
        return 99999.0;
    }
    double span1 = a.end - a.start, span2 = b.end - b.start;
    double mx = span1 > span2 ? span1 : span2;
    if (type == SVIM_INS) {                                           // :64-77
        double pd = fabs(a.start - b.start) / p.pos_norm;
        if (pd > 2.0 * p.cluster_max_distance) {
            if (mx == 0.0) { *err = 1; return 0.0; }
            return pd + fabs(span1 - span2) / mx;
        }
        if (mx == 0.0) { *err = 1; return 0.0; }
        return pd + ed / mx / p.edit_norm;
    }
    double c1 = floor((a.start + a.end) / 2.0), c2 = floor((b.start + b.end) / 2.0);
    double pd = fabs(c1 - c2) / p.pos_norm;
    if (mx == 0.0) { *err = 1; return 0.0; }
    double sd = fabs(span1 - span2) / mx;
    if (type == SVIM_DUP_INT || type == SVIM_DUP_INT_CAND) {         // :78-86 and :110-119 (candidates)
        double pdd = fabs(a.dpos - b.dpos) / p.pos_norm;
        return pd + pdd + sd;
    }
    return pd + sd;                                                   // DEL / DUP_TAN / INV :48-63
}

SVIM_HD bool ins_gate_needs_ed(const SigView& a, const SigView& b, const ClusterParams& p) {
    return !(fabs(a.start - b.start) / p.pos_norm > 2.0 * p.cluster_max_distance);
}

// partition gap test of form_partitions (:23) with the per-type downstream_distance_to
SVIM_HD bool gap_exceeds(int type, double a_start, double a_end, double a_dpos, double b_start, double b_dpos, double max_distance) {
    double g;
    if (type == SVIM_INS) g = b_start - a_start;
    else if (type == SVIM_DUP_INT) g = b_dpos - a_dpos;
    else g = b_start - a_end;
    if (g < 0.0) g = 0.0;
    return g > max_distance;
}

// order-preserving map double -> uint64 for radix sorting
SVIM_HD uint64_t double_key(double d) {
    d = d + 0.0;   // -0.0 -> +0.0
    uint64_t b;
    memcpy(&b, &d, 8);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// ---- dendrogram post-processing (sequential, m <= 100) -----------------------------------------
// Z rows come out of nn-chain unsorted: (x, y, d, size) with x<y cluster REPRESENTATIVE indices.
// 1. stable sort by d   2. union-find relabel   3. max-dist   4. DFS numbering.
struct LinkScratch {
    int* zx; int* zy; double* zd;          // m-1 rows (in: unsorted; out: sorted+relabelled)
    int* order;                            // m-1
    int* parent;                           // 2m-1
    double* md;                            // m-1
    int* stack;                            // m
    unsigned char* visited;                // 2m-1
};

SVIM_HD int fcluster_from_chain(int m, const LinkScratch& s, const int* ux, const int* uy, const double* ud, double t, int* T) {
    const int r = m - 1;
    // stable insertion sort of row indices by distance (scipy: argsort kind='mergesort')
    for (int i = 0; i < r; ++i) {
        int j = i - 1; double d = ud[i];
        while (j >= 0 && ud[s.order[j]] > d) { s.order[j + 1] = s.order[j]; --j; }
        s.order[j + 1] = i;
    }
    for (int i = 0; i < 2 * m - 1; ++i) s.parent[i] = i;
    for (int k = 0; k < r; ++k) {
        int o = s.order[k];
        int a = ux[o], b = uy[o];
        while (s.parent[a] != a) a = s.parent[a];
        while (s.parent[b] != b) b = s.parent[b];
        // path compression is an optimisation only; roots are what matter
        s.zx[k] = a < b ? a : b; s.zy[k] = a < b ? b : a; s.zd[k] = ud[o];
        s.parent[a] = m + k; s.parent[b] = m + k;
    }
    for (int k = 0; k < r; ++k) {
        double v = s.zd[k];
        if (s.zx[k] >= m && s.md[s.zx[k] - m] > v) v = s.md[s.zx[k] - m];
        if (s.zy[k] >= m && s.md[s.zy[k] - m] > v) v = s.md[s.zy[k] - m];
        s.md[k] = v;
    }
    for (int i = 0; i < 2 * m - 1; ++i) s.visited[i] = 0;
    int sp = 0, leader = -1, ncl = 0;
    s.stack[sp++] = 2 * m - 2;
    while (sp > 0) {
        int root = s.stack[sp - 1];
        int k = root - m;
        int left = s.zx[k], right = s.zy[k];
        if (leader == -1 && s.md[k] <= t) { leader = root; ++ncl; }
        if (left >= m && !s.visited[left]) { s.visited[left] = 1; s.stack[sp++] = left; continue; }
        if (right >= m && !s.visited[right]) { s.visited[right] = 1; s.stack[sp++] = right; continue; }
        if (left < m) { if (leader == -1) ++ncl; T[left] = ncl; }
        if (right < m) { if (leader == -1) ++ncl; T[right] = ncl; }
        if (leader == root) leader = -1;
        --sp;
    }
    return ncl;
}

// ---- consolidation ----------------------------------------------------------------------------------
// Correctly rounded sqrt(N / D) for exact non-negative integers (statistics.stdev goes through exact
// rationals and _float_sqrt_of_frac).  N < 2^62, 0 < D < 2^20.
// sign of  N/D - (m2 * 2^e)^2  for integers N >= 0, D > 0, m2 < 2^54, e <= 0 (exact, 128-bit)
SVIM_HD int cmp_ratio_sq(uint64_t N, uint64_t D, uint64_t m2, int e) {
    const unsigned __int128 rhs = (unsigned __int128)m2 * m2 * D;      // < 2^108 * 2^20
    const int sh = -2 * e;                                               // compare N * 2^sh with rhs
    if (N == 0) return rhs == 0 ? 0 : -1;
    if (sh >= 128) return 1;
    if (sh > 0 && (((unsigned __int128)N) >> (128 - sh)) != 0) return 1;
    const unsigned __int128 lhs = ((unsigned __int128)N) << sh;
    return lhs < rhs ? -1 : (lhs > rhs ? 1 : 0);
}

// Correctly rounded sqrt(N / D) for exact integers (statistics.stdev goes through exact rationals and
// _float_sqrt_of_frac): start from the FP64 estimate and move to a neighbour if the exact value lies beyond
// the midpoint.  N < 2^62, 0 < D < 2^20.
SVIM_HD double sqrt_ratio(int64_t N, int64_t D) {
    if (N <= 0) return 0.0;
    double r = sqrt((double)N / (double)D);
    for (int it = 0; it < 2; ++it) {
        uint64_t bits; memcpy(&bits, &r, 8);
        const int E = (int)((bits >> 52) & 0x7ff);
        if (E == 0 || E == 0x7ff) return r;
        const uint64_t mant = (bits & 0xfffffffffffffull) | (1ull << 52);
        const int ex = E - 1075;                     // r = mant * 2^ex
        if (ex - 1 > 0) return r;                     // outside the range this path is used for
        const int up = cmp_ratio_sq((uint64_t)N, (uint64_t)D, 2 * mant + 1, ex - 1);     // value vs midpoint above
        if (up > 0 || (up == 0 && (mant & 1))) { bits += 1; memcpy(&r, &bits, 8); continue; }
        if (mant == (1ull << 52)) return r;           // below a power of two the spacing halves; the estimate is never that far off
        const int dn = cmp_ratio_sq((uint64_t)N, (uint64_t)D, 2 * mant - 1, ex - 1);     // value vs midpoint below
        if (dn < 0 || (dn == 0 && (mant & 1))) { bits -= 1; memcpy(&r, &bits, 8); continue; }
        return r;
    }
    return r;
}

// sample standard deviation of v[0..n) (n >= 2).  Integral inputs (and half-integral, scale=2) take an
// exact integer path; anything else (float coordinates through the Python seam) a two-pass FP64 path.
SVIM_HD double stdev_values(const double* v, int n, int stride) {
    bool exact = true;
    double v0 = v[0];
    for (int i = 0; i < n; ++i) { double x = v[i * stride] * 2.0; if (x != floor(x) || fabs(x) > 9.0e15 || fabs(x - 2.0 * v0) > 2.0e7) exact = false; }
    if (exact) {
        int64_t s1 = 0, s2 = 0;
        for (int i = 0; i < n; ++i) { int64_t y = (int64_t)(v[i * stride] * 2.0 - v0 * 2.0); s1 += y; s2 += y * y; }
        // var = (n*s2 - s1^2) / (n (n-1)) / 4
        int64_t N = (int64_t)n * s2 - s1 * s1;
        return sqrt_ratio(N, (int64_t)n * (n - 1) * 4);
    }
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += v[i * stride];
    mean /= n;
    double ss = 0.0;
    for (int i = 0; i < n; ++i) { double d = v[i * stride] - mean; ss += d * d; }
    return sqrt(ss / (n - 1));
}

SVIM_HD double py_min1(double x) { return x < 1.0 ? x : 1.0; }   // min(1, x)

// calculate_score (:183-211); has_std == 0 <=> std_* is None
SVIM_HD double cluster_score(int n_eff, int has_std, double std_span, double std_pos, double span) {
    double sds = 0.0, pds = 0.0;
    if (has_std) { sds = 1.0 - py_min1(std_span / span); pds = 1.0 - py_min1(std_pos / span); }
    double num = (double)(n_eff < 80 ? n_eff : 80);
    return num + sds * (num / 8.0) + pds * (num / 8.0);
}

SVIM_HD int64_t py_round_int(double x) { return (int64_t)rint(x); }   // int(round(x)): half to even

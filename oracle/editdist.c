/* ORACLE (test infrastructure, not product code).
 *
 * Global (Needleman-Wunsch, unit cost) Levenshtein distance between two byte
 * strings.  Stands in for `edlib.align(a, b)["editDistance"]` with edlib's
 * defaults mode="NW", task="distance", k=-1, no additional equalities
 * (reference call site: SVIM_clustering.py:45).  edlib itself is an un-vendored
 * third-party dependency (setup.py:41, unpinned); the integer result of a
 * global unit-cost alignment is implementation independent, so any exact
 * algorithm is a valid oracle.  No reference test pins this call ("parity
 * unpinned" for the third-party part; the two implementations below are
 * cross-checked against each other in tests/test_oracle_editdist.py).
 *
 *   oracle_editdist_dp     textbook Wagner-Fischer, two rows, O(n*m)
 *   oracle_editdist_myers  Myers/Hyyro bit-vector, 64-bit blocks, O(n*m/64)
 *                          (used as the CPU baseline: edlib is the same family)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

long oracle_editdist_dp(const unsigned char* a, long n, const unsigned char* b, long m) {
    if (n == 0) return m;
    if (m == 0) return n;
    long* prev = (long*)malloc(sizeof(long) * (size_t)(m + 1));
    long* cur = (long*)malloc(sizeof(long) * (size_t)(m + 1));
    for (long j = 0; j <= m; ++j) prev[j] = j;
    for (long i = 1; i <= n; ++i) {
        cur[0] = i;
        unsigned char ca = a[i - 1];
        for (long j = 1; j <= m; ++j) {
            long best = prev[j - 1] + (ca != b[j - 1]);
            long up = prev[j] + 1, left = cur[j - 1] + 1;
            if (up < best) best = up;
            if (left < best) best = left;
            cur[j] = best;
        }
        long* t = prev; prev = cur; cur = t;
    }
    long r = prev[m];
    free(prev); free(cur);
    return r;
}

/* Pattern = a (rows, length n), text = b (columns, length m).
 * Column-wise Myers with horizontal carry between 64-row blocks; the top
 * boundary D[0][j] = j is a +1 horizontal delta into block 0. */
long oracle_editdist_myers(const unsigned char* a, long n, const unsigned char* b, long m) {
    if (n == 0) return m;
    if (m == 0) return n;
    long W = (n + 63) / 64;
    uint64_t* peq = (uint64_t*)calloc((size_t)(256 * W), sizeof(uint64_t));
    uint64_t* pv = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)W);
    uint64_t* mv = (uint64_t*)calloc((size_t)W, sizeof(uint64_t));
    for (long i = 0; i < n; ++i) peq[(long)a[i] * W + (i >> 6)] |= 1ull << (i & 63);
    for (long w = 0; w < W; ++w) pv[w] = ~0ull;
    long score = n;
    int last_bits = (int)((n - 1) & 63);
    for (long j = 0; j < m; ++j) {
        const uint64_t* eqrow = peq + (long)b[j] * W;
        int hin = 1; /* D[0][j]-D[0][j-1] = +1 */
        for (long w = 0; w < W; ++w) {
            uint64_t Eq = eqrow[w], Pv = pv[w], Mv = mv[w];
            uint64_t hin_neg = (hin < 0) ? 1ull : 0ull;
            uint64_t Xv = Eq | Mv;
            Eq |= hin_neg;
            uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            int top = (w == W - 1) ? last_bits : 63;
            int hout = (int)((Ph >> top) & 1) - (int)((Mh >> top) & 1);
            Ph <<= 1; Mh <<= 1;
            if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
            pv[w] = Mh | ~(Xv | Ph);
            mv[w] = Ph & Xv;
            hin = hout;
        }
        score += hin;
    }
    free(peq); free(pv); free(mv);
    return score;
}

// Test-only host build of the SVIM_HD (host+device) logic headers, so the branchy per-read /
// per-cluster code that the kernels execute can be fuzzed against the oracle without a GPU.
// Not part of the product: libsvimgpu.so never calls this.
#define SVIM_HOST_ONLY
#include <vector>
#include <cstring>
#include "../../svim_b200/csrc/collect.cuh"
#include "../../svim_b200/csrc/cluster.cuh"

struct VecEmitter {
    std::vector<svim_sig> m, t;
    void sig(const svim_sig& s) { m.push_back(s); }
    void twin(const svim_sig& s) { t.push_back(s); }
};

extern "C" {

// One primary record + SA tag through parse_sa_segments / sort_chain / analyze_chain.
int hc_chain(int32_t tid, int32_t pos, uint32_t flag, int64_t l_seq, const uint32_t* cigar, int32_t n_cigar, const uint8_t* sa, int32_t sa_len,
             int32_t n_contigs, const char* names, const int32_t* name_off, const int32_t* rank,
             int64_t min_sv, int64_t max_sv, int64_t tol_g, int64_t tol_o, int32_t min_mapq, int32_t all_bnds,
             svim_sig* out_main, int32_t* n_main, svim_sig* out_twin, int32_t* n_twin, int32_t cap, uint32_t* err_out) {
    ContigTable ct{n_contigs, names, name_off, rank};
    ChainParams p{min_sv, max_sv, tol_g, tol_o, min_mapq, all_bnds};
    CigarSummary cs; cigsum_init(cs);
    for (int k = 0; k < n_cigar; ++k) cigsum_add(cs, cigar[k] & 15u, cigar[k] >> 4);
    Seg chain[SVIM_MAX_SEGMENTS]; int n = 0; uint32_t err = 0;
    Seg s; int64_t rl; s.tid = tid;
    const int rev = (flag & 0x10u) ? 1 : 0;
    cigsum_finish(cs, l_seq, pos, rev, s, rl);
    VecEmitter out;
    if (cs.hard == 0) {
        if (rev && rl < 0) err |= CH_NO_READLEN; else chain[n++] = s;
        n = parse_sa_segments(sa, sa_len, ct, p, l_seq, chain, n, err);
    } else if (!(rev && rl < 0)) chain[n++] = s;
    sort_chain(chain, n);
    PrimaryInfo pi{0, 0, l_seq, rl};
    Junction junc[SVIM_MAX_SEGMENTS]; Tandem tand[SVIM_MAX_SEGMENTS];
    if (n >= 2) analyze_chain(chain, n, pi, p, ct, out, 0x80000000u, 0x80000000u, err, junc, tand);
    *n_main = (int32_t)out.m.size(); *n_twin = (int32_t)out.t.size();
    for (size_t k = 0; k < out.m.size() && (int)k < cap; ++k) out_main[k] = out.m[k];
    for (size_t k = 0; k < out.t.size() && (int)k < cap; ++k) out_twin[k] = out.t[k];
    *err_out = err;
    return 0;
}

double hc_spd(int type, const double* a, const double* b, uint32_t dirs_a, uint32_t dirs_b, double pos_norm, double edit_norm, double cmd, double ed, int* err) {
    SigView va{a[0], a[1], a[2], 0, (uint8_t)dirs_a}, vb{b[0], b[1], b[2], 1, (uint8_t)dirs_b};
    ClusterParams cp{1000.0, pos_norm, edit_norm, cmd};
    return spd(type, va, vb, cp, ed, err);
}

// unsorted nn-chain rows -> flat clusters (the sequential half of the linkage kernel)
int hc_fcluster(int m, const int* ux, const int* uy, const double* ud, double t, int* T) {
    std::vector<int> zx(m), zy(m), order(m), parent(2 * m), stack(m + 1);
    std::vector<double> zd(m), md(m);
    std::vector<unsigned char> vis(2 * m);
    LinkScratch s{zx.data(), zy.data(), zd.data(), order.data(), parent.data(), md.data(), stack.data(), vis.data()};
    return fcluster_from_chain(m, s, ux, uy, ud, t, T);
}

double hc_stdev(const double* v, int n) { return stdev_values(v, n, 1); }
double hc_score(int n_eff, int has, double a, double b, double span) { return cluster_score(n_eff, has, a, b, span); }
long long hc_round(double x) { return py_round_int(x); }
void hc_py_slice(long long a, long long n, long long L, long long* lo, long long* hi) { int64_t l, h; py_slice(a, n, L, l, h); *lo = l; *hi = h; }

}

// ---- banded Myers wavefront (myers_band.cuh): replay the G lanes of one pair step by step --------------------------
#include "../../svim_b200/csrc/myers_band.cuh"

template <int G, int WPL, bool HI>
static long long band_emulate(const uint8_t* pat, long long m, const uint8_t* txt, long long n, long long k) {
    const BandGeom ge = band_geom(m, n, k, WPL);
    std::vector<uint64_t> planes((size_t)3 * ge.NG * WPL);
    for (long long w = 0; w < (long long)ge.NG * WPL; ++w) band_build_word(pat, m, ge.pad, w, planes.data() + 3 * w);
    BandLane<WPL> L[G];
    uint32_t e_out[G], recv[G];
    for (int l = 0; l < G; ++l) { band_lane_init(L[l], ge, planes.data(), txt, l, true); e_out[l] = 0; }
    const int32_t steps = ge.n + ge.NG - 1;
    for (int32_t s = 0; s < steps; ++s) {
        for (int l = 0; l < G; ++l) recv[l] = e_out[(l - 1) & (G - 1)];
        for (int l = 0; l < G; ++l) e_out[l] = band_lane_step<G, WPL, HI>(L[l], ge, planes.data(), s, recv[l], e_out[l], 1u, 2u);
    }
    long long d = m;
    for (int l = 0; l < G; ++l) d += L[l].score;
    return d;
}

extern "C" {
// pat/txt: symbol codes 0..3, m >= n >= 1.  Returns the banded result, -1 if the band does not fit the shape, -2 on bad arguments.
long long hc_myers_banded(const uint8_t* pat, long long m, const uint8_t* txt, long long n, long long k, int bin, int hi) {
    if (m < n || n < 1 || k < m - n || bin < 0 || bin >= MYERS_BINS) return -2;
    const MyersBin sp = myers_bin_spec(bin);
    if (!myers_band_fits(m, n, k, sp.G, sp.WPL)) return -1;
#define HC_BAND(G_, W_) return hi ? band_emulate<G_, W_, true>(pat, m, txt, n, k) : band_emulate<G_, W_, false>(pat, m, txt, n, k);
    switch (bin) {
        case 0: HC_BAND(4, 1) case 1: HC_BAND(4, 2) case 2: HC_BAND(4, 3) case 3: HC_BAND(4, 4) case 4: HC_BAND(8, 3)
        case 5: HC_BAND(8, 4) case 6: HC_BAND(16, 3) case 7: HC_BAND(16, 4) case 8: HC_BAND(32, 3) default: HC_BAND(32, 4)
    }
#undef HC_BAND
}
int hc_myers_band_bin(long long m, long long n, int num, int add) { return myers_band_bin(m, n, num, add); }
long long hc_myers_band_k(long long m, long long n, int num, int add) { return myers_band_k(m, n, num, add); }
}

// ---- thread-per-pair banded Myers (myers_tpp.cuh): the per-thread program of k_myers_tpp, run on the host ------------------
#include "../../svim_b200/csrc/myers_tpp.cuh"

template <int B> struct HostEq {
    uint32_t v[B][4];
    void put(int slot, const uint32_t x[4]) { for (int c = 0; c < 4; ++c) v[slot][c] = x[c]; }
    uint32_t get(int slot, uint32_t sym) const { return v[slot][sym >> 5]; }          // symbols arrive pre-scaled by 32
    void shift_up() { for (int i = 0; i + 1 < B; ++i) for (int c = 0; c < 4; ++c) v[i][c] = v[i + 1][c]; }
};
struct HostPeq {
    const uint8_t* pat; long long m;
    void block(int b, uint32_t v[4]) const {
        v[0] = v[1] = v[2] = v[3] = 0u;
        if (b < 0 || 32ll * b >= m) return;
        const long long left = m - 32ll * b;
        tpp_masks_from_codes(pat + 32ll * b, (int)(left < 32 ? left : 32), v);
    }
};
struct HostTxt {
    std::vector<uint8_t> t;     // index s = j + phase, symbol * 32
    uint32_t byte(int s) const { return t[s]; }
    uint32_t word(int s) const { return (uint32_t)t[s] | ((uint32_t)t[s + 1] << 8) | ((uint32_t)t[s + 2] << 16) | ((uint32_t)t[s + 3] << 24); }
};
template <int B>
static long long tpp_emulate(const uint8_t* pat, long long m, const uint8_t* txt, long long n, int a) {
    HostEq<B> eq; HostPeq peq{pat, m}; HostTxt tx;
    const int phase = a >= 0 ? ((-a) & 31) : 0;
    tx.t.assign((size_t)(phase + n + 8), 0);
    for (long long j = 0; j < n; ++j) tx.t[(size_t)(phase + j)] = (uint8_t)(txt[j] * 32);
    return tpp_thread<B>((int32_t)m, (int32_t)n, a, eq, peq, tx, 1u, 2u) ;
}
extern "C" {
// pat/txt: symbol codes 0..3, m >= n >= 1.  k < 0: unbanded.  Runs in the smallest bucket that holds the pair (or `force_B` blocks);
// returns the result (exact iff unbanded or <= k), -1 if no bucket holds it, -2 on bad arguments.
long long hc_myers_tpp(const uint8_t* pat, long long m, const uint8_t* txt, long long n, long long k, int force_B) {
    if (m < n || n < 1 || (k >= 0 && k < m - n)) return -2;
    int a = -1; long long B = tpp_blocks_full(m);
    if (k >= 0) { B = tpp_blocks_band(m, n, k); a = (int)((k - (m - n)) / 2); }
    if (force_B > 0) { if (force_B < B) return -1; B = force_B; }
    const int q = tpp_bucket_of((int)(B > 1000 ? 1000 : B));
    if (q < 0) return -1;
    switch (tpp_bucket_B(q)) {
#define HC_TPP(B_) case B_: return tpp_emulate<B_>(pat, m, txt, n, a);
        HC_TPP(2) HC_TPP(3) HC_TPP(4) HC_TPP(5) HC_TPP(6) HC_TPP(7) HC_TPP(8) HC_TPP(9) HC_TPP(10) HC_TPP(12) HC_TPP(14) HC_TPP(16)
        HC_TPP(18) HC_TPP(20) HC_TPP(22) HC_TPP(24) HC_TPP(26) HC_TPP(28)
#undef HC_TPP
    }
    return -1;
}
int hc_tpp_plan(long long m, long long n, int num, int add, int* a_out) { const TppPlan p = tpp_plan(m, n, num, add); *a_out = p.a; return p.B; }
}

// ---- host+device core of the on-GPU BAM decoder (csrc/bgzf_core.cuh): raw DEFLATE of a BGZF payload, record-start search ----
#include "../../svim_b200/csrc/bgzf_core.cuh"
extern "C" {
int hc_bgzf_inflate(const uint8_t* src, unsigned clen, uint8_t* dst, unsigned ulen) {
    std::vector<uint32_t> tab(BGZF_TABLE_WORDS);
    return bgzf_inflate_block(src, clen, dst, ulen, tab.data());
}
unsigned long long hc_bam_find_record_start(const uint8_t* data, unsigned long long size, unsigned long long from, int n_ref, int depth) {
    return bam_find_record_start(data, size, from, n_ref, depth);
}
}

// The parallel record-boundary scheme of DESIGN.md §11, replayed serially: every chunk finds a speculative record start on its
// own, chains to the next chunk's territory, and the scheme is accepted iff every chain lands on the next chunk's start.
// Returns the total number of records, -1 if some chain did not land (the GPU path then repairs from that chunk serially),
// -2 on a truncated stream.  starts_out (optional) receives each chunk's start.
extern "C" long long hc_bam_chunked_starts(const uint8_t* data, unsigned long long size, unsigned long long first_record, unsigned long long chunk,
                                           int n_ref, int depth, unsigned long long* starts_out) {
    const unsigned long long n_chunks = (size - first_record + chunk - 1) / chunk;
    std::vector<unsigned long long> st(n_chunks + 1), en(n_chunks);
    std::vector<uint32_t> cnt(n_chunks);
    for (unsigned long long c = 0; c < n_chunks; ++c)          // independent per chunk (one thread each on the GPU)
        st[c] = c == 0 ? first_record : bam_find_record_start(data, size, first_record + c * chunk, n_ref, depth);
    st[n_chunks] = size;
    for (unsigned long long c = 0; c < n_chunks; ++c) {        // independent per chunk
        const unsigned long long limit = c + 1 < n_chunks ? first_record + (c + 1) * chunk : size;
        cnt[c] = 0;
        en[c] = st[c] >= limit ? st[c] : bam_chain(data, size, st[c], limit, &cnt[c]);
        if (en[c] == ~0ull) return -2;
    }
    // chunk 0 starts on a true record start, and a chain that starts on a true start ends on the first true start of the next
    // territory: by induction every chunk is right iff each chain lands on the next chunk's own guess
    long long total = 0;
    for (unsigned long long c = 0; c < n_chunks; ++c) {
        if (en[c] != st[c + 1]) return -1;
        total += cnt[c];
        if (starts_out) starts_out[c] = st[c];
    }
    return total;
}

// ---- svim_aln_soa.cigar8 expansion: the warp loop of k_expand_cigar8 (collect.cu) replayed with 32 sequential "lanes" ----
extern "C" int hc_expand_cigar8(const uint8_t* src, unsigned long long nb, unsigned long long nc, uint32_t* dst) {
    unsigned long long out = 0; uint32_t carry = 0, err = 0;
    for (unsigned long long base = 0; base < nb; base += 512) {
        uint32_t w[32][4], cnt[32], pre[32], tail[32];
        for (int lane = 0; lane < 32; ++lane) {
            const unsigned long long at = base + (unsigned long long)lane * 16;
            for (int q = 0; q < 4; ++q) w[lane][q] = 0x0F0F0F0Fu;
            if (at < nb) memcpy(w[lane], src + at, 16);
            cnt[lane] = 16u - c8_ext_bytes(w[lane]);
            tail[lane] = c8_tail(w[lane]);
        }
        uint32_t run = 0;
        for (int lane = 0; lane < 32; ++lane) { run += cnt[lane]; pre[lane] = run; }
        for (int lane = 0; lane < 32; ++lane) {
            const uint32_t init = lane == 0 ? carry : tail[lane - 1];
            const uint32_t got = c8_decode_lane(w[lane], init, out + pre[lane] - cnt[lane], nc, dst, &err);
            if (got != cnt[lane]) return -2;          // the scan's count and the decoder's must agree
        }
        out += pre[31]; carry = tail[31];
    }
    return (err || out != nc || carry) ? 1 : 0;
}

// k_expand_cigar8_staged: same round, words staged in a 512-word area and written out in 32-word rows aligned to the destination
extern "C" int hc_expand_cigar8_staged(const uint8_t* src, unsigned long long nb, unsigned long long nc, uint32_t* dst, unsigned dst_word_phase) {
    unsigned long long out = 0; uint32_t carry = 0, err = 0;
    std::vector<uint32_t> stage(512);
    for (unsigned long long base = 0; base < nb; base += 512) {
        uint32_t w[32][4], cnt[32], pre[32], tail[32];
        for (int lane = 0; lane < 32; ++lane) {
            const unsigned long long at = base + (unsigned long long)lane * 16;
            for (int q = 0; q < 4; ++q) w[lane][q] = 0x0F0F0F0Fu;
            if (at < nb) memcpy(w[lane], src + at, 16);
            cnt[lane] = 16u - c8_ext_bytes(w[lane]);
            tail[lane] = c8_tail(w[lane]);
        }
        uint32_t run = 0;
        for (int lane = 0; lane < 32; ++lane) { run += cnt[lane]; pre[lane] = run; }
        for (uint32_t& x : stage) x = 0xABABABABu;
        for (int lane = 0; lane < 32; ++lane) {
            const uint32_t init = lane == 0 ? carry : tail[lane - 1];
            c8_decode_lane(w[lane], init, pre[lane] - cnt[lane], 512, stage.data(), &err);
        }
        const uint32_t total = pre[31];
        const int shift = (int)((dst_word_phase + out) & 31u);
        for (int lane = 0; lane < 32; ++lane)
            for (int k = lane - shift; k < (int)total; k += 32)
                if (k >= 0 && out + (unsigned long long)k < nc) dst[out + (unsigned long long)k] = stage[k];
        out += total; carry = tail[31];
    }
    return (err || out != nc || carry) ? 1 : 0;
}


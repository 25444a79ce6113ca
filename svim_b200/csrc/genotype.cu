// GENOTYPE on the resident record buffer: genotype(candidates, bam, type, options), SVIM_genotyping.py:34-93.
//
// The reference asks pysam for the alignments of a window around every candidate (bam.fetch, :49) and walks them one by
// one.  Here the coordinate-sorted record rows are already in HBM (uploaded for COLLECT), so a region fetch is two binary
// searches and the walk is one warp per candidate over 32 records at a time:
//
//   k_ref_end      bam_endpos of every record (pysam reference_end): one warp per record streams the BAM-encoded CIGAR with
//                  128-bit loads and sums the lengths of M/D/N/=/X ops.  HBM-bound, 4 B per CIGAR op.
//   k_row_bounds   first/last row of every contig + a sortedness check of (tid, pos) (fetch needs an indexed, i.e.
//                  coordinate-sorted, file: svim:93-98).
//   scan by key    running maximum of bam_endpos inside each contig: the first row whose running maximum exceeds the window
//                  start is exactly the first record htslib's iterator can return, however long the alignments before it are.
//   k_genotype     one warp per candidate: rows [first, lower_bound(pos >= stop)) in file order, 32 per step; skip variant
//                  reads (sorted id list, binary search), unmapped / secondary / low-MAPQ rows (:63-66); the first 500
//                  remaining rows (:57) are tested against the locus (:69-76); supporting read ids go to shared memory and
//                  are counted distinct (a set of read names, :53); then the genotype decision (:78-93).
#pragma once
#include <cub/cub.cuh>
#include "ctx.cuh"

struct GenoMax {
    __host__ __device__ __forceinline__ int32_t operator()(const int32_t& a, const int32_t& b) const { return a > b ? a : b; }
};

// bam_endpos (htslib): pos + reference length for mapped records with a CIGAR (a zero length counts as 1), else pos + 1.
__global__ void __launch_bounds__(256) k_ref_end(DevSoa s, int32_t* __restrict__ ref_end) {
    const uint32_t lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < s.n; r += n_warps) {
        const uint32_t n_ops = s.n_cigar[r];
        uint32_t acc = 0;
        if (n_ops && !(s.flag[r] & 0x4)) {
            const uint4* v = reinterpret_cast<const uint4*>(s.cigar + s.cigar_off[r]);
            const uint32_t n_vec = (n_ops + 3) >> 2;          // records are padded to 16 bytes with zero words (0M)
            for (uint32_t k = lane; k < n_vec; k += 128) {
                uint4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) w[u] = (k + 32 * u < n_vec) ? __ldcs(v + k + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc += (w[u].x >> 4) & (0u - ((SVIM_MASK_REF_TRUE >> (w[u].x & 15)) & 1u));
                    acc += (w[u].y >> 4) & (0u - ((SVIM_MASK_REF_TRUE >> (w[u].y & 15)) & 1u));
                    acc += (w[u].z >> 4) & (0u - ((SVIM_MASK_REF_TRUE >> (w[u].z & 15)) & 1u));
                    acc += (w[u].w >> 4) & (0u - ((SVIM_MASK_REF_TRUE >> (w[u].w & 15)) & 1u));
                }
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane == 0) ref_end[r] = s.pos[r] + (int32_t)(acc ? acc : 1u);
    }
}

// rows of contig t are [lo[t], hi[t]); bad[0] counts adjacent rows out of (tid, pos) order (tid -1 = unplaced sorts last)
__global__ void k_row_bounds(const int32_t* __restrict__ tid, const int32_t* __restrict__ pos, int64_t n, int32_t n_contigs,
                             int32_t* __restrict__ lo, int32_t* __restrict__ hi, uint32_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t t = tid[i];
    if (i == 0) { if (t >= 0 && t < n_contigs) lo[t] = 0; }
    else {
        const int32_t tp = tid[i - 1];
        if ((uint32_t)tp > (uint32_t)t || (tp == t && t >= 0 && pos[i - 1] > pos[i])) atomicAdd(bad, 1u);
        if (tp != t) {
            if (t >= 0 && t < n_contigs) lo[t] = (int32_t)i;
            if (tp >= 0 && tp < n_contigs) hi[tp] = (int32_t)i;
        }
    }
    if (i == n - 1 && t >= 0 && t < n_contigs) hi[t] = (int32_t)n;
}

struct GenoArgs {
    DevSoa s;
    const int32_t* ref_end; const int32_t* run_max;     // bam_endpos, running maximum inside the contig
    const int32_t* row_lo; const int32_t* row_hi;
    const svim_geno_cand* cand; int64_t n_cand;
    const uint32_t* variant_ids;
    const int64_t* contig_len;
    svim_geno_result* out;
    svim_geno_params p;
    int32_t ins_like;                                    // INS / DUP_INT: end := start (:45), test :74-76
};

#define GENO_MAX_ALN 500          // `while aln_no < 500`, SVIM_genotyping.py:57
#define GENO_WARPS 8

__global__ void __launch_bounds__(GENO_WARPS * 32) k_genotype(GenoArgs a) {
    __shared__ uint32_t s_ids[GENO_WARPS][GENO_MAX_ALN + 12];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * GENO_WARPS + wib;
    if (c >= a.n_cand) return;
    const svim_geno_cand cd = a.cand[c];
    svim_geno_result res;
    res.support_fraction = nan(""); res.ref_reads = 0; res.alt_reads = (int32_t)cd.n_variant_reads; res.genotype = 3; res.status = 0;
    res.pad = 0; res.n_fetched = 0;
    const int64_t start = cd.start, end = a.ins_like ? cd.start : cd.end;
    const int64_t clen = a.contig_len[cd.tid];
    const int64_t fs = start - 1000 > 0 ? start - 1000 : 0, fe = end + 1000 < clen ? end + 1000 : clen;     // :49
    if (fs > fe) {                                        // pysam: ValueError "invalid coordinates"
        res.status = 2;
        if (lane == 0) a.out[c] = res;
        return;
    }
    // first row whose running-maximum end exceeds fs, last = first row with pos >= fe
    int32_t lo = a.row_lo[cd.tid], hi = a.row_hi[cd.tid];
    int32_t first, last;
    { int32_t l = lo, h = hi; while (l < h) { int32_t m = l + ((h - l) >> 1); if ((int64_t)a.run_max[m] > fs) h = m; else l = m + 1; } first = l; }
    { int32_t l = first, h = hi; while (l < h) { int32_t m = l + ((h - l) >> 1); if ((int64_t)a.s.pos[m] >= fe) h = m; else l = m + 1; } last = l; }
    // thresholds of :69-76 — (end - start) / 2 is a float in the reference; every operand is exact in FP64
    const double d_start = (double)start, d_end = (double)end;
    const double min_overlap = fmin((d_end - d_start) / 2.0, 2000.0);
    const double thr_a = d_end - min_overlap, thr_b = d_end + 100.0, thr_c = d_start - 100.0, thr_d = d_start + min_overlap;
    const uint32_t* var = a.variant_ids + cd.variant_off;
    const uint32_t n_var = cd.n_variant_reads;
    uint32_t counted = 0, n_sup = 0, n_fetched = 0, err = 0;
    for (int32_t base = first; base < last && counted < GENO_MAX_ALN; base += 32) {
        const int32_t i = base + (int32_t)lane;
        bool ok = false; int32_t rs = 0, re = 0; uint32_t qid = 0, ncig = 0;
        if (i < last) {
            re = a.ref_end[i];
            if ((int64_t)re > fs) {                       // htslib iterator: pos < stop (by `last`) and endpos > start
                ok = true;
                rs = a.s.pos[i]; qid = a.s.qname_id[i]; ncig = a.s.n_cigar[i];
                const uint32_t fl = a.s.flag[i];
                const bool filtered = (fl & 0x104u) || (int32_t)a.s.mapq[i] < a.p.min_mapq;     // :65
                uint32_t l = 0, h = n_var;                                                       // :63
                while (l < h) { uint32_t m = (l + h) >> 1; if (var[m] < qid) l = m + 1; else h = m; }
                const bool is_variant = l < n_var && var[l] == qid;
                n_fetched += 1;
                ok = !filtered && !is_variant;
            }
        }
        const uint32_t m_ok = __ballot_sync(0xffffffffu, ok);
        const uint32_t rank = counted + __popc(m_ok & ((1u << lane) - 1u));
        bool sup = false;
        if (ok && rank < GENO_MAX_ALN) {
            const double d_rs = (double)rs, d_re = (double)re;
            if (a.ins_like) {
                const bool left = d_rs < thr_c;                                                  // :75
                if (left && ncig == 0) err = 1;          // reference_end is None: the comparison raises TypeError
                sup = left && d_re > thr_b;
            } else {
                const bool a1 = d_rs < thr_a, c1 = d_rs < thr_c;                                  // :71-72
                if (ncig == 0 && (a1 || c1)) err = 1;
                sup = (a1 && d_re > thr_b) || (c1 && d_re > thr_d);
            }
        }
        const uint32_t m_sup = __ballot_sync(0xffffffffu, sup);
        if (sup) s_ids[wib][n_sup + __popc(m_sup & ((1u << lane) - 1u))] = qid;
        n_sup += __popc(m_sup);
        counted += __popc(m_ok);
    }
    __syncwarp();
    // distinct read ids (set of query names, :53/:73/:76)
    uint32_t firsts = 0;
    for (uint32_t j = lane; j < n_sup; j += 32) {
        const uint32_t v = s_ids[wib][j];
        bool seen = false;
        for (uint32_t k = 0; k < j && !seen; ++k) seen = s_ids[wib][k] == v;
        firsts += seen ? 0u : 1u;
    }
    for (int o = 16; o > 0; o >>= 1) {
        firsts += __shfl_xor_sync(0xffffffffu, firsts, o);
        n_fetched += __shfl_xor_sync(0xffffffffu, n_fetched, o);
        err |= __shfl_xor_sync(0xffffffffu, err, o);
    }
    if (lane != 0) return;
    const int64_t n_ref = firsts, tot = (int64_t)n_var + n_ref;
    res.ref_reads = (int32_t)n_ref; res.n_fetched = n_fetched; res.status = (uint8_t)err;
    if (tot >= a.p.minimum_depth) {                       // :78-88
        if (tot == 0) res.status = 3;                     // ZeroDivisionError (minimum_depth <= 0)
        else {
            const double f = (double)n_var / (double)tot;
            res.support_fraction = f;
            res.genotype = f >= a.p.homozygous_threshold ? 0 : (f >= a.p.heterozygous_threshold && f < a.p.homozygous_threshold) ? 1
                           : f < a.p.heterozygous_threshold ? 2 : 3;
        }
    } else if (tot > 0) {                                 // :89-91
        res.support_fraction = (double)n_var / (double)tot;
    }
    a.out[c] = res;
}

// bam_endpos, contig row ranges and the running maximum for the resident rows; cached until the next upload.
static int genotype_prepare(svimgpu_ctx* ctx, int32_t n_contigs) {
    if (ctx->geno_ready && ctx->geno_contigs == n_contigs) return 0;
    const int64_t n = ctx->soa.n;
    cudaStream_t st = ctx->stream;
    StageTimer t(ctx, T_GENO_PREP);
    SVIM_CUDA(ctx->d_geno_end.ensure((size_t)(n + 1) * 4)); SVIM_CUDA(ctx->d_geno_max.ensure((size_t)(n + 1) * 4));
    SVIM_CUDA(ctx->d_geno_rows.ensure((size_t)(2 * n_contigs + 2) * 4));
    SVIM_CUDA(cudaMemsetAsync(ctx->d_geno_rows.p, 0, (size_t)(2 * n_contigs + 2) * 4, st));
    int32_t* lo = ctx->d_geno_rows.as<int32_t>(); int32_t* hi = lo + n_contigs; uint32_t* bad = (uint32_t*)(hi + n_contigs);
    if (n) {
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        { ctx->launches++; k_ref_end<<<sms * 8, 256, 0, st>>>(ctx->soa, ctx->d_geno_end.as<int32_t>()); }
        { ctx->launches++; k_row_bounds<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(ctx->soa.tid, ctx->soa.pos, n, n_contigs, lo, hi, bad); }
        size_t tmp = 0;
        cub::DeviceScan::InclusiveScanByKey(nullptr, tmp, ctx->soa.tid, ctx->d_geno_end.as<int32_t>(), ctx->d_geno_max.as<int32_t>(), GenoMax(), (int)n,
                                            cub::Equality(), st);
        SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
        SVIM_CUDA(cub::DeviceScan::InclusiveScanByKey(ctx->d_sort_tmp.p, tmp, ctx->soa.tid, ctx->d_geno_end.as<int32_t>(), ctx->d_geno_max.as<int32_t>(),
                                                      GenoMax(), (int)n, cub::Equality(), st));
    }
    uint32_t h_bad = 0;
    SVIM_CUDA(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    SVIM_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        ctx->set_error(SVIMGPU_ERR_STATE, "records are not coordinate-sorted (%u inversions): region fetch needs a sorted, indexed file", h_bad);
        return SVIMGPU_ERR_STATE;
    }
    ctx->geno_ready = true; ctx->geno_contigs = n_contigs;
    return 0;
}

static int genotype_run(svimgpu_ctx* ctx, int32_t type, const svim_geno_params* gp, int64_t n, const svim_geno_cand* cands, const uint32_t* variant_ids,
                        int64_t n_variant_ids, const int64_t* contig_lengths, int32_t n_contigs, svim_geno_result* out) {
    if (!ctx->rows_resident) { ctx->set_error(SVIMGPU_ERR_STATE, "no alignment records on the device (upload or collect first)"); return SVIMGPU_ERR_STATE; }
    if (type != SVIM_DEL && type != SVIM_INV && type != SVIM_INS && type != SVIM_DUP_INT) {
        ctx->set_error(SVIMGPU_ERR_ARG, "genotype: type must be DEL, INV, INS or DUP_INT (svim:161-170)"); return SVIMGPU_ERR_ARG;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (cands[i].tid < 0 || cands[i].tid >= n_contigs) { ctx->set_error(SVIMGPU_ERR_ARG, "genotype: candidate %lld has no valid contig", (long long)i); return SVIMGPU_ERR_ARG; }
        if (cands[i].variant_off + cands[i].n_variant_reads > (uint64_t)n_variant_ids) { ctx->set_error(SVIMGPU_ERR_ARG, "genotype: variant read list out of range"); return SVIMGPU_ERR_ARG; }
    }
    int rc = genotype_prepare(ctx, n_contigs);
    if (rc) return rc;
    if (n == 0) return 0;
    cudaStream_t st = ctx->stream;
    StageTimer t(ctx, T_GENO);
    SVIM_CUDA(ctx->d_geno_cand.ensure((size_t)n * sizeof(svim_geno_cand))); SVIM_CUDA(ctx->d_geno_out.ensure((size_t)n * sizeof(svim_geno_result)));
    SVIM_CUDA(ctx->d_geno_var.ensure((size_t)n_variant_ids * 4 + 4)); SVIM_CUDA(ctx->d_geno_clen.ensure((size_t)n_contigs * 8));
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_geno_cand.p, cands, (size_t)n * sizeof(svim_geno_cand), cudaMemcpyHostToDevice, st));
    if (n_variant_ids) SVIM_CUDA(cudaMemcpyAsync(ctx->d_geno_var.p, variant_ids, (size_t)n_variant_ids * 4, cudaMemcpyHostToDevice, st));
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_geno_clen.p, contig_lengths, (size_t)n_contigs * 8, cudaMemcpyHostToDevice, st));
    GenoArgs a;
    a.s = ctx->soa; a.ref_end = ctx->d_geno_end.as<int32_t>(); a.run_max = ctx->d_geno_max.as<int32_t>();
    a.row_lo = ctx->d_geno_rows.as<int32_t>(); a.row_hi = a.row_lo + n_contigs;
    a.cand = ctx->d_geno_cand.as<svim_geno_cand>(); a.n_cand = n; a.variant_ids = ctx->d_geno_var.as<uint32_t>();
    a.contig_len = ctx->d_geno_clen.as<int64_t>(); a.out = ctx->d_geno_out.as<svim_geno_result>(); a.p = *gp;
    a.ins_like = (type == SVIM_INS || type == SVIM_DUP_INT) ? 1 : 0;
    { ctx->launches++; k_genotype<<<(uint32_t)((n + GENO_WARPS - 1) / GENO_WARPS), GENO_WARPS * 32, 0, st>>>(a); }
    SVIM_CUDA(cudaGetLastError());
    SVIM_CUDA(cudaMemcpyAsync(out, ctx->d_geno_out.p, (size_t)n * sizeof(svim_geno_result), cudaMemcpyDeviceToHost, st));
    SVIM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- cut&paste search: flag_cutpaste_candidates (SVIM_merging.py:12-29) -----------------------------------------------------
// For every DUP_INT cluster the closest deletion cluster under span_position_distance_clusters (SVIM_clustering.py:99-107):
// O(#DUP_INT x #DEL) in the reference (a Python list + sort per cluster); here one warp per DUP_INT cluster strides over the
// deletion intervals, FP64 in the reference's operation order (-fmad=false), first minimum wins like the stable sort at :20.
__global__ void __launch_bounds__(256) k_closest_source(const int64_t* __restrict__ a_start, const int64_t* __restrict__ a_end, int64_t n_a,
                                                        const int64_t* __restrict__ b_start, const int64_t* __restrict__ b_end, int64_t n_b,
                                                        double normalizer, int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                                        uint32_t* __restrict__ zero_div) {
    const uint32_t lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= n_a) return;
    const int64_t s2 = a_start[q], e2 = a_end[q];
    const int64_t span2 = e2 - s2, c2 = (s2 + e2) >> 1;              // Python // on ints: floor
    double best = INFINITY; int64_t best_j = INT64_MAX; uint32_t zd = 0;
    for (int64_t j = lane; j < n_b; j += 32) {
        const int64_t s1 = b_start[j], e1 = b_end[j];
        const int64_t span1 = e1 - s1, c1 = (s1 + e1) >> 1;
        const int64_t mx = span1 > span2 ? span1 : span2;
        if (mx == 0) { zd = 1; continue; }
        const int64_t dc = c1 > c2 ? c1 - c2 : c2 - c1, ds = span1 > span2 ? span1 - span2 : span2 - span1;
        const double d = (double)dc / normalizer + (double)ds / (double)mx;
        if (d < best) { best = d; best_j = j; }                       // ascending j per lane: strict < keeps the first minimum
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        zd |= __shfl_xor_sync(0xffffffffu, zd, o);
        if (od < best || (od == best && oj < best_j)) { best = od; best_j = oj; }
    }
    if (lane == 0) { out_idx[q] = best_j == INT64_MAX ? -1 : best_j; out_dist[q] = best; if (zd) atomicOr(zero_div, 1u); }
}

static int closest_source_run(svimgpu_ctx* ctx, int64_t n_a, const int64_t* a_start, const int64_t* a_end, int64_t n_b, const int64_t* b_start,
                              const int64_t* b_end, double normalizer, int64_t* out_idx, double* out_dist) {
    cudaStream_t st = ctx->stream;
    StageTimer t(ctx, T_CUTPASTE);
    const size_t na = (size_t)n_a, nb = (size_t)n_b;
    SVIM_CUDA(ctx->d_geno_cand.ensure((2 * na + 2 * nb) * 8 + 64)); SVIM_CUDA(ctx->d_geno_out.ensure(na * 16 + 64));
    int64_t* d_as = ctx->d_geno_cand.as<int64_t>(); int64_t* d_ae = d_as + na; int64_t* d_bs = d_ae + na; int64_t* d_be = d_bs + nb;
    int64_t* d_idx = ctx->d_geno_out.as<int64_t>(); double* d_dist = (double*)(d_idx + na); uint32_t* d_zd = (uint32_t*)(d_dist + na);
    SVIM_CUDA(cudaMemcpyAsync(d_as, a_start, na * 8, cudaMemcpyHostToDevice, st)); SVIM_CUDA(cudaMemcpyAsync(d_ae, a_end, na * 8, cudaMemcpyHostToDevice, st));
    if (nb) { SVIM_CUDA(cudaMemcpyAsync(d_bs, b_start, nb * 8, cudaMemcpyHostToDevice, st)); SVIM_CUDA(cudaMemcpyAsync(d_be, b_end, nb * 8, cudaMemcpyHostToDevice, st)); }
    SVIM_CUDA(cudaMemsetAsync(d_zd, 0, 4, st));
    { ctx->launches++; k_closest_source<<<(uint32_t)((na * 32 + 255) / 256), 256, 0, st>>>(d_as, d_ae, n_a, d_bs, d_be, n_b, normalizer, d_idx, d_dist, d_zd); }
    SVIM_CUDA(cudaGetLastError());
    uint32_t zd = 0;
    SVIM_CUDA(cudaMemcpyAsync(out_idx, d_idx, na * 8, cudaMemcpyDeviceToHost, st)); SVIM_CUDA(cudaMemcpyAsync(out_dist, d_dist, na * 8, cudaMemcpyDeviceToHost, st));
    SVIM_CUDA(cudaMemcpyAsync(&zd, d_zd, 4, cudaMemcpyDeviceToHost, st));
    SVIM_CUDA(cudaStreamSynchronize(st));
    if (zd) { ctx->set_error(SVIMGPU_ERR_DATA, "two zero-length source intervals: the reference divides by max(span1, span2) = 0"); return SVIMGPU_ERR_DATA; }
    return 0;
}

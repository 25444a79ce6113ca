// C ABI (include/svimgpu.h) over the COLLECT / CLUSTER kernels.  Single translation unit.
#include <cuda_runtime.h>
#include <nccl.h>
#include <algorithm>
#include <string>
#include <vector>
#include "ctx.cuh"
#include "collect.cu"
#include "cluster.cu"
#include "nccl.cu"
#include "genotype.cu"
#include "bam.cu"

const char* const k_stage_names[T_N] = {
    "h2d_alignments", "cigar_scan", "segment_chain", "sort_back+ins_gather", "ins_gather", "collect_d2h", "sig_to_csig", "key_sort",
    "partition", "host_sampling", "ins_pair_list", "myers_edit_distance", "linkage", "consolidate", "final_order", "cluster_d2h", "nccl_exchange", "genotype_prepare", "genotype", "closest_deletion",
    "bam_h2d+inflate", "bam_record_bounds", "bam_rows", "bam_fill", "bam_read_names", "ins_peer_fetch"};

extern "C" {

const char* svimgpu_version(void) { return "svimgpu 0.1 (sm_100a)"; }

int svimgpu_create(svimgpu_ctx** out, int device, const svim_params* params) {
    if (!out || !params) return SVIMGPU_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "svimgpu_create: no CUDA device (%s); this library has no CPU path\n", cudaGetErrorString(e));
        return SVIMGPU_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return SVIMGPU_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SVIMGPU_ERR_CUDA;
    svimgpu_ctx* ctx = new svimgpu_ctx();
    ctx->device = device; ctx->params = *params;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SVIMGPU_ERR_CUDA; }
    for (int i = 0; i < 2 * T_N; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaEventCreate(&ctx->user_ev[0]); cudaEventCreate(&ctx->user_ev[1]);
    for (int i = 0; i < SVIM_AUX_STREAMS; ++i) cudaStreamCreateWithFlags(&ctx->aux_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i <= SVIM_AUX_STREAMS; ++i) cudaEventCreateWithFlags(&ctx->aux_ev[i], cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_collect_done, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&ctx->ev_host_copy[i], cudaEventDisableTiming);
    timings_begin(ctx);
    memset(&ctx->cstats, 0, sizeof(ctx->cstats)); memset(&ctx->clstats, 0, sizeof(ctx->clstats));
    myers_init_symcode();
    if (const char* v = getenv("SVIM_SCAN_VARIANT")) ctx->scan_variant = atoi(v);
    if (const char* v = getenv("SVIM_SCAN_CHUNKS")) ctx->scan_chunks = atoi(v);
    if (const char* v = getenv("SVIM_MYERS_MODE")) ctx->myers_mode = atoi(v);
    if (const char* v = getenv("SVIM_MYERS_TPP")) ctx->myers_tpp = atoi(v);
    if (const char* v = getenv("SVIM_MYERS_TRACE")) ctx->myers_trace = atoi(v);
    if (const char* v = getenv("SVIM_PEER_INS")) ctx->peer_ins = atoi(v) != 0;
    if (const char* v = getenv("SVIM_EXPAND8_STAGED")) ctx->expand8_staged = atoi(v) != 0;
    if (const char* v = getenv("SVIM_MYERS_BAND")) { int num = 0, add = 24; if (sscanf(v, "%d,%d", &num, &add) >= 1) { ctx->myers_band_num = num; ctx->myers_band_add = add; } }
    *out = ctx;
    return 0;
}

void svimgpu_destroy(svimgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_teardown(ctx);
    DevBuf* bufs[] = {&ctx->d_names, &ctx->d_name_off, &ctx->d_rank, &ctx->d_rank_to_tid, &ctx->d_genome, &ctx->d_genome_off, &ctx->d_counters,
                      &ctx->d_queue[0], &ctx->d_queue[1], &ctx->d_work, &ctx->d_sort_tmp, &ctx->d_keys[0], &ctx->d_keys[1], &ctx->d_vals[0],
                      &ctx->d_vals[1], &ctx->d_scan, &ctx->sets[0].recs, &ctx->sets[0].ins, &ctx->sets[1].recs, &ctx->sets[1].ins, &ctx->d_csig,
                      &ctx->d_csig_sorted, &ctx->d_cins, &ctx->d_order, &ctx->d_head, &ctx->d_partid, &ctx->d_part_off, &ctx->d_samp_off,
                      &ctx->d_samp_idx, &ctx->d_labels, &ctx->d_part_ncl, &ctx->d_part_nkept, &ctx->d_part_stats, &ctx->d_plist,
                      &ctx->d_user_rank_to_tid, &ctx->d_cl_off, &ctx->d_mem_off, &ctx->d_clusters, &ctx->d_clusters_sorted,
                      &ctx->d_members, &ctx->d_pair_off, &ctx->d_pair_ed, &ctx->d_pairs, &ctx->d_ckeys[0], &ctx->d_ckeys[1], &ctx->d_cvals[0],
                      &ctx->d_cvals[1], &ctx->d_xchg[0], &ctx->d_xchg[1], &ctx->d_xchg[2], &ctx->d_xchg[3], &ctx->d_xchg[4], &ctx->d_xchg[5], &ctx->d_xchg[6], &ctx->d_xchg[7],
                      &ctx->d_genome_codes, &ctx->d_ins_codes, &ctx->d_myers_trace, &ctx->d_big_list, &ctx->d_big_caps, &ctx->d_big_scratch, &ctx->d_cig16, &ctx->d_cig16_off, &ctx->d_cig16_err, &ctx->d_bam_names, &ctx->d_bam_name_off, &ctx->d_bam_rec_of_id, &ctx->d_pmeta, &ctx->d_ppref, &ctx->d_ptype, &ctx->d_hdr, &ctx->d_large_list, &ctx->d_picks,
                      &ctx->d_seg_tab, &ctx->d_shard_ins, &ctx->d_shard_len};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 14; ++i) ctx->d_soa[i].release();
    for (int i = 0; i < 48; ++i) ctx->d_myers_scratch[i].release();
    ctx->d_myers_ctl.release();
    ctx->h_hdr.release(); ctx->h_picks.release(); ctx->h_clusters.release(); ctx->h_members.release();
    ctx->d_stage.release(); ctx->d_stage_off.release();
    for (DevBuf* b : {&ctx->d_geno_end, &ctx->d_geno_max, &ctx->d_geno_rows, &ctx->d_geno_cand, &ctx->d_geno_out, &ctx->d_geno_var, &ctx->d_geno_clen}) b->release();
    ctx->d_qs_info.release(); ctx->d_qs_grp.release(); ctx->d_qs_segsum.release(); ctx->d_qs_mem_off.release(); ctx->d_qs_mem_idx.release();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) { if (ctx->h_out[i]) cudaFreeHost(ctx->h_out[i]); cudaEventDestroy(ctx->ev_host_copy[i]); }
    cudaEventDestroy(ctx->ev_collect_done); cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2 * T_N; ++i) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < SVIM_AUX_STREAMS; ++i) cudaStreamDestroy(ctx->aux_stream[i]);
    for (int i = 0; i <= SVIM_AUX_STREAMS; ++i) cudaEventDestroy(ctx->aux_ev[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int svimgpu_pci_bus_id(svimgpu_ctx* ctx, char* out, int32_t cap) {
    if (!ctx || !out || cap < 16) return SVIMGPU_ERR_ARG;
    SVIM_CUDA(cudaDeviceGetPCIBusId(out, cap, ctx->device));
    return 0;
}

const char* svimgpu_last_error(const svimgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int svimgpu_set_params(svimgpu_ctx* ctx, const svim_params* p) {
    if (!ctx || !p) return SVIMGPU_ERR_ARG;
    ctx->params = *p;
    return 0;
}

int svimgpu_set_contigs(svimgpu_ctx* ctx, int32_t n, const char* names, const int32_t* name_off) {
    if (!ctx || n <= 0 || !names || !name_off) return SVIMGPU_ERR_ARG;
    if (n >= (1 << 30)) { ctx->set_error(SVIMGPU_ERR_LIMIT, "too many contigs"); return SVIMGPU_ERR_LIMIT; }
    cudaSetDevice(ctx->device);
    std::vector<int32_t> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    auto name = [&](int i) { return std::string(names + name_off[i], names + name_off[i + 1]); };
    std::vector<std::string> nm(n);
    for (int i = 0; i < n; ++i) nm[i] = name(i);
    // Python str ordering == byte order for ASCII contig names
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return nm[a] < nm[b]; });
    ctx->h_rank.assign(n, 0); ctx->h_rank_to_tid.assign(n, 0);
    int r = -1;
    for (int k = 0; k < n; ++k) {
        if (k == 0 || nm[idx[k]] != nm[idx[k - 1]]) { ++r; ctx->h_rank_to_tid[r] = idx[k]; }
        ctx->h_rank[idx[k]] = r;   // equal names share a rank
    }
    ctx->n_contigs = n;
    SVIM_CUDA(ctx->d_names.ensure((size_t)name_off[n] + 1)); SVIM_CUDA(ctx->d_name_off.ensure((size_t)(n + 1) * 4));
    SVIM_CUDA(ctx->d_rank.ensure((size_t)n * 4)); SVIM_CUDA(ctx->d_rank_to_tid.ensure((size_t)n * 4));
    SVIM_CUDA(cudaMemcpy(ctx->d_names.p, names, (size_t)name_off[n], cudaMemcpyHostToDevice));
    SVIM_CUDA(cudaMemcpy(ctx->d_name_off.p, name_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice));
    SVIM_CUDA(cudaMemcpy(ctx->d_rank.p, ctx->h_rank.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    SVIM_CUDA(cudaMemcpy(ctx->d_rank_to_tid.p, ctx->h_rank_to_tid.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    return 0;
}

int svimgpu_set_genome(svimgpu_ctx* ctx, int32_t n, const int64_t* offsets, const uint8_t* bytes) {
    if (!ctx || n <= 0 || !offsets || !bytes) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    SVIM_CUDA(ctx->d_genome.ensure((size_t)offsets[n] + 16)); SVIM_CUDA(ctx->d_genome_off.ensure((size_t)(n + 1) * 8));
    SVIM_CUDA(cudaMemcpy(ctx->d_genome.p, bytes, (size_t)offsets[n], cudaMemcpyHostToDevice));
    SVIM_CUDA(cudaMemcpy(ctx->d_genome_off.p, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice));
    ctx->genome_bytes = offsets[n]; ctx->genome_contigs = n;
    // symbol-code image for the thread-per-pair edit-distance kernels (myers.cu, k_tpp_encode): once per genome
    SVIM_CUDA(ctx->d_genome_codes.ensure((size_t)offsets[n] + 16));
    if (offsets[n] > 0) { ctx->launches++; k_tpp_encode<<<(unsigned)(((size_t)offsets[n] + 16 * 256 - 1) / (16 * 256)), 256, 0, ctx->stream>>>(ctx->d_genome.as<uint8_t>(), offsets[n], ctx->d_genome_codes.as<uint8_t>()); }
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int svimgpu_pin_host(void* p, int64_t bytes) { return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : SVIMGPU_ERR_CUDA; }
int svimgpu_unpin_host(void* p) { return cudaHostUnregister(p) == cudaSuccess ? 0 : SVIMGPU_ERR_CUDA; }

static int upload_alignments(svimgpu_ctx* ctx, const svim_aln_soa* s, bool with_seq);

int svimgpu_upload_alignments(svimgpu_ctx* ctx, const svim_aln_soa* s) { return upload_alignments(ctx, s, true); }

static int upload_alignments(svimgpu_ctx* ctx, const svim_aln_soa* s, bool with_seq) {
    if (!ctx || !s || s->n_aln < 0) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    timings_begin(ctx);
    const int64_t n = s->n_aln;
    const void* src[14] = {s->tid, s->pos, s->flag, s->mapq, s->n_cigar, s->cigar_off, s->l_seq, s->seq_off, s->sa_off, s->sa_len, s->qname_id,
                           s->cigar, s->seq, s->sa};
    const size_t bytes[14] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * 2, (size_t)n, (size_t)n * 4, (size_t)n * 8, (size_t)n * 4, (size_t)n * 8,
                              (size_t)n * 8, (size_t)n * 4, (size_t)n * 4, (size_t)s->cigar_words * 4, (size_t)s->seq_bytes, (size_t)s->sa_bytes};
    {
        StageTimer t(ctx, T_H2D);
        const bool packed8 = s->cigar8 != nullptr && s->cigar8_off != nullptr;
        const bool packed = !packed8 && s->cigar16 != nullptr && s->cigar16_off != nullptr;
        for (int i = 0; i < 14; ++i) {
            if (i == 12 && !with_seq) continue;   // SEQ blob stays on the host (lazy path)
            SVIM_CUDA(ctx->d_soa[i].ensure(bytes[i] + 64));
            if (i == 11 && (packed || packed8)) continue;      // CIGAR crosses as a packed stream below
            if (bytes[i]) SVIM_CUDA(cudaMemcpyAsync(ctx->d_soa[i].p, src[i], bytes[i], cudaMemcpyHostToDevice, ctx->stream));
        }
        if (packed || packed8) {
            // a half / a quarter of the PCIe bytes: upload the packed stream, expand it to BAM's uint32 words in HBM (the expansion
            // writes 4 B per operation at HBM speed, ~25x the PCIe rate the copy just ran at)
            const size_t pbytes = packed8 ? (size_t)s->cigar8_bytes : (size_t)s->cigar16_words * 2;
            const void* psrc = packed8 ? (const void*)s->cigar8 : (const void*)s->cigar16;
            const uint64_t* poff = packed8 ? s->cigar8_off : s->cigar16_off;
            SVIM_CUDA(ctx->d_cig16.ensure(pbytes + 64)); SVIM_CUDA(ctx->d_cig16_off.ensure((size_t)(n + 1) * 8));
            SVIM_CUDA(ctx->d_cig16_err.ensure(16)); SVIM_CUDA(cudaMemsetAsync(ctx->d_cig16_err.p, 0, 4, ctx->stream));
            SVIM_CUDA(cudaMemcpyAsync(ctx->d_cig16_off.p, poff, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            if (n > 0) {
                // the stream crosses in slices of whole records on the copy stream while the slices already there are expanded on the
                // main stream: the expansion (4 B written per operation) hides behind the PCIe copy instead of following it
                int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
                const size_t unit = packed8 ? 1 : 2;                      // bytes per offset unit
                const int max_slices = SVIM_AUX_STREAMS;                  // one event per slice (aux_ev)
                const uint64_t total_units = poff[n];
                const uint64_t per = std::max<uint64_t>((total_units + max_slices - 1) / max_slices, (uint64_t)(4u << 20) / unit);
                int64_t r0 = 0; int slice = 0;
                while (r0 < n) {
                    int64_t r1;
                    if (slice == max_slices - 1) r1 = n;
                    else { const uint64_t want = poff[r0] + per; r1 = std::upper_bound(poff + r0, poff + n + 1, want) - poff; if (r1 <= r0) r1 = r0 + 1; if (r1 > n) r1 = n; }
                    const size_t b0 = (size_t)poff[r0] * unit, b1 = (size_t)poff[r1] * unit;
                    if (b1 > b0) SVIM_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_cig16.p + b0, (const uint8_t*)psrc + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->copy_stream));
                    SVIM_CUDA(cudaEventRecord(ctx->aux_ev[slice], ctx->copy_stream));
                    SVIM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[slice], 0));
                    const int64_t nr = r1 - r0;
                    const int blocks = (int)std::min<int64_t>((int64_t)sms * 8, (nr + 7) / 8);
                    ctx->launches++;
                    if (packed8 && ctx->expand8_staged) k_expand_cigar8_staged<<<blocks, 256, 0, ctx->stream>>>(ctx->d_cig16.as<uint8_t>(), ctx->d_cig16_off.as<uint64_t>() + r0, ctx->d_soa[4].as<uint32_t>() + r0,
                                                                                  ctx->d_soa[5].as<uint64_t>() + r0, nr, ctx->d_soa[11].as<uint32_t>(), ctx->d_cig16_err.as<uint32_t>());
                    else if (packed8) k_expand_cigar8<<<blocks, 256, 0, ctx->stream>>>(ctx->d_cig16.as<uint8_t>(), ctx->d_cig16_off.as<uint64_t>() + r0, ctx->d_soa[4].as<uint32_t>() + r0,
                                                                                  ctx->d_soa[5].as<uint64_t>() + r0, nr, ctx->d_soa[11].as<uint32_t>(), ctx->d_cig16_err.as<uint32_t>());
                    else k_expand_cigar16<<<blocks, 256, 0, ctx->stream>>>(ctx->d_cig16.as<uint16_t>(), ctx->d_cig16_off.as<uint64_t>() + r0, ctx->d_soa[4].as<uint32_t>() + r0,
                                                                           ctx->d_soa[5].as<uint64_t>() + r0, nr, ctx->d_soa[11].as<uint32_t>(), ctx->d_cig16_err.as<uint32_t>());
                    r0 = r1; ++slice;
                }
                SVIM_CUDA(cudaGetLastError());
            }
        }
    }
    DevSoa& d = ctx->soa;
    d.n = n;
    d.tid = ctx->d_soa[0].as<int32_t>(); d.pos = ctx->d_soa[1].as<int32_t>(); d.flag = ctx->d_soa[2].as<uint16_t>(); d.mapq = ctx->d_soa[3].as<uint8_t>();
    d.n_cigar = ctx->d_soa[4].as<uint32_t>(); d.cigar_off = ctx->d_soa[5].as<uint64_t>(); d.l_seq = ctx->d_soa[6].as<int32_t>();
    d.seq_off = ctx->d_soa[7].as<uint64_t>(); d.sa_off = ctx->d_soa[8].as<uint64_t>(); d.sa_len = ctx->d_soa[9].as<uint32_t>();
    d.qname_id = ctx->d_soa[10].as<uint32_t>(); d.cigar = ctx->d_soa[11].as<uint32_t>(); d.seq = ctx->d_soa[12].as<uint8_t>(); d.sa = ctx->d_soa[13].as<uint8_t>();
    ctx->cigar_words = s->cigar_words; ctx->seq_bytes = s->seq_bytes; ctx->sa_bytes = s->sa_bytes;
    uint32_t bad16 = 0;
    const bool any_packed = (s->cigar16 && s->cigar16_off) || (s->cigar8 && s->cigar8_off);
    if (any_packed && n > 0) SVIM_CUDA(cudaMemcpyAsync(&bad16, ctx->d_cig16_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad16) { ctx->have_soa = false; ctx->set_error(SVIMGPU_ERR_ARG, "cigar8 / cigar16: a record's packed stream does not hold n_cigar operations"); return SVIMGPU_ERR_ARG; }
    ctx->have_soa = true; ctx->collected = false;
    ctx->rows_resident = true; ctx->geno_ready = false;
    ctx->lazy_seq = !with_seq; ctx->h_seq = with_seq ? nullptr : s->seq; ctx->h_seq_off = with_seq ? nullptr : s->seq_off; ctx->lazy_aln_base = 0;
    timings_end(ctx);
    return 0;
}

int svimgpu_decode_bam(svimgpu_ctx* ctx, const uint8_t* file, int64_t file_bytes, const svim_bgzf_block* blocks, int64_t n_blocks, int64_t first_record,
                       int32_t n_ref, svim_bam_info* info) {
    if (!ctx || (file_bytes > 0 && !file) || n_blocks < 0 || (n_blocks > 0 && !blocks) || first_record < 0) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->host_copy[0] || ctx->host_copy[1]) { cudaStreamSynchronize(ctx->copy_stream); ctx->host_copy[0] = ctx->host_copy[1] = false; }
    ctx->have_soa = false; ctx->rows_resident = false;
    static_assert(sizeof(svim_bgzf_block) == sizeof(BgBlock), "block table layout");
    timings_begin(ctx);
    const int rc = bam_decode_run(ctx, file, file_bytes, (const BgBlock*)blocks, n_blocks, first_record, n_ref, info);
    timings_end(ctx);
    return rc;
}

int svimgpu_fetch_bam_names(svimgpu_ctx* ctx, uint8_t* names, uint64_t* name_off, uint32_t* rec_of_id, uint32_t* qname_id) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    if (!ctx->rows_resident || !ctx->d_bam_names.p) { ctx->set_error(SVIMGPU_ERR_STATE, "no BAM-decoded record buffer is resident"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->soa.n;
    if (names && ctx->bam_names_bytes) SVIM_CUDA(cudaMemcpyAsync(names, ctx->d_bam_names.p, (size_t)ctx->bam_names_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (name_off && n) SVIM_CUDA(cudaMemcpyAsync(name_off, ctx->d_bam_name_off.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (rec_of_id && ctx->bam_n_names) SVIM_CUDA(cudaMemcpyAsync(rec_of_id, ctx->d_bam_rec_of_id.p, (size_t)ctx->bam_n_names * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (qname_id && n) SVIM_CUDA(cudaMemcpyAsync(qname_id, ctx->d_soa[10].p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// the resident record buffer back on the host (tests: the GPU decoder against the host decoder; callers that want an AlignmentBatch)
int svimgpu_download_alignments(svimgpu_ctx* ctx, int32_t* tid, int32_t* pos, uint16_t* flag, uint8_t* mapq, uint32_t* n_cigar, uint64_t* cigar_off, int32_t* l_seq,
                                uint64_t* seq_off, uint64_t* sa_off, uint32_t* sa_len, uint32_t* qname_id, uint32_t* cigar, uint8_t* seq, uint8_t* sa) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    if (!ctx->rows_resident || !ctx->have_soa) { ctx->set_error(SVIMGPU_ERR_STATE, "no record buffer is resident"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->soa.n;
    void* dst[14] = {tid, pos, flag, mapq, n_cigar, cigar_off, l_seq, seq_off, sa_off, sa_len, qname_id, cigar, seq, sa};
    const size_t bytes[14] = {n * 4, n * 4, n * 2, n, n * 4, n * 8, n * 4, n * 8, n * 8, n * 4, n * 4, (size_t)ctx->cigar_words * 4, (size_t)ctx->seq_bytes, (size_t)ctx->sa_bytes};
    for (int i = 0; i < 14; ++i) if (dst[i] && bytes[i]) SVIM_CUDA(cudaMemcpyAsync(dst[i], ctx->d_soa[i].p, bytes[i], cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int svimgpu_download_cigar(svimgpu_ctx* ctx, uint32_t* out, int64_t words) {
    if (!ctx || !out || words < 0) return SVIMGPU_ERR_ARG;
    if (!ctx->rows_resident || words > ctx->cigar_words) { ctx->set_error(SVIMGPU_ERR_STATE, "no resident record buffer of that size"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    if (words) SVIM_CUDA(cudaMemcpy(out, ctx->d_soa[11].p, (size_t)words * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int svimgpu_collect(svimgpu_ctx* ctx, svim_collect_stats* stats) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    // a host copy of the previous lists may still be in flight on the copy stream: the lists are about to be rewritten
    if (ctx->host_copy[0] || ctx->host_copy[1]) { cudaStreamSynchronize(ctx->copy_stream); ctx->host_copy[0] = ctx->host_copy[1] = false; }
    ctx->host_copy_pending = false;
    double h2d = ctx->ms[T_H2D];
    timings_begin(ctx);
    int rc = collect_run(ctx, stats);
    timings_end(ctx);
    ctx->ms[T_H2D] = h2d;
    return rc;
}

static size_t host_copy_ins_off(const SigSet& set) { return ((size_t)set.n * sizeof(svim_sig) + 255) & ~(size_t)255; }

// INS blob of a list -> host.  A segmented list (after the exchange) is read piece by piece from the ranks that hold it: the copy
// engine pulls a peer's piece over NVLink and writes it out over this GPU's PCIe link.  The peers must not have started their next
// collect (include/svimgpu.h, svimgpu_exchange_signatures).
static int copy_ins_to_host(svimgpu_ctx* ctx, const SigSet& set, uint8_t* dst, cudaStream_t s) {
    if (!set.segmented) {
        if (set.ins_bytes) SVIM_CUDA(cudaMemcpyAsync(dst, set.ins.p, (size_t)set.ins_bytes, cudaMemcpyDeviceToHost, s));
        return 0;
    }
    for (size_t r = 0; r + 1 < set.seg_base.size(); ++r) {
        const int64_t len = set.seg_base[r + 1] - set.seg_base[r];
        if (len <= 0) continue;
        if (!set.seg_ptr[r]) { ctx->set_error(SVIMGPU_ERR_STATE, "the insertion bytes of rank %d are not mapped", (int)r); return SVIMGPU_ERR_STATE; }
        SVIM_CUDA(cudaMemcpyAsync(dst + set.seg_base[r], set.seg_ptr[r], (size_t)len, cudaMemcpyDefault, s));
    }
    return 0;
}

// Start copying the collected lists to pinned host memory on the copy stream; CLUSTER (ctx->stream) runs meanwhile.
static int start_host_copy(svimgpu_ctx* ctx) {
    SVIM_CUDA(cudaEventRecord(ctx->ev_collect_done, ctx->stream));
    SVIM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_collect_done, 0));
    for (int w = 0; w < 2; ++w) {
        SigSet& set = ctx->sets[w];
        ctx->host_copy[w] = false;
        if (w == 1 && !ctx->params.all_bnds) continue;
        const size_t ins_off = host_copy_ins_off(set), need = ins_off + (size_t)set.ins_bytes + 256;
        if (need > ctx->h_out_cap[w]) {
            if (ctx->h_out[w]) cudaFreeHost(ctx->h_out[w]);
            ctx->h_out[w] = nullptr; ctx->h_out_cap[w] = 0;
            const size_t want = need + need / 4;
            SVIM_CUDA(cudaMallocHost((void**)&ctx->h_out[w], want));
            ctx->h_out_cap[w] = want;
        }
        if (set.n) SVIM_CUDA(cudaMemcpyAsync(ctx->h_out[w], set.recs.p, (size_t)set.n * sizeof(svim_sig), cudaMemcpyDeviceToHost, ctx->copy_stream));
        ctx->host_copy_has_ins[w] = !set.segmented || ctx->mirror_gathered_ins;
        if (ctx->host_copy_has_ins[w]) { int rc = copy_ins_to_host(ctx, set, ctx->h_out[w] + ins_off, ctx->copy_stream); if (rc) return rc; }
        SVIM_CUDA(cudaEventRecord(ctx->ev_host_copy[w], ctx->copy_stream));
        ctx->host_copy[w] = true;
    }
    return 0;
}

int svimgpu_collect_host(svimgpu_ctx* ctx, const svim_aln_soa* soa, svim_collect_stats* stats) {
    // SEQ is only read where an insertion is emitted: keep it on the host and upload just those bytes
    int rc = upload_alignments(ctx, soa, false);
    if (rc) return rc;
    rc = svimgpu_collect(ctx, stats);
    ctx->have_soa = false;      // the host SEQ pointers must not outlive this call
    ctx->lazy_seq = false; ctx->h_seq = nullptr; ctx->h_seq_off = nullptr;
    // one process per GPU: the lists are about to be replaced by the gathered ones, which svimgpu_exchange_signatures mirrors instead
    if (!rc) { if (ctx->nccl_comm) { ctx->host_copy[0] = ctx->host_copy[1] = false; ctx->host_copy_pending = true; } else rc = start_host_copy(ctx); }
    return rc;
}

int svimgpu_signatures_host(svimgpu_ctx* ctx, int which, const svim_sig** sigs, const uint8_t** ins) {
    if (!ctx || which < 0 || which > 1 || !sigs || !ins) return SVIMGPU_ERR_ARG;
    *sigs = nullptr; *ins = nullptr;
    if (!ctx->collected || !ctx->host_copy[which]) return 0;     // no host copy (resident collect, or the lists were exchanged): use svimgpu_fetch_signatures
    cudaSetDevice(ctx->device);
    SVIM_CUDA(cudaEventSynchronize(ctx->ev_host_copy[which]));
    *sigs = (const svim_sig*)ctx->h_out[which];
    *ins = ctx->host_copy_has_ins[which] ? ctx->h_out[which] + host_copy_ins_off(ctx->sets[which]) : nullptr;
    return 0;
}

int svimgpu_mirror_gathered_ins(svimgpu_ctx* ctx, int with_ins) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    ctx->mirror_gathered_ins = with_ins != 0;
    return 0;
}

int svimgpu_collect_host_querysorted(svimgpu_ctx* ctx, const svim_aln_soa* soa, svim_collect_stats* stats) {
    if (!ctx || !soa) return SVIMGPU_ERR_ARG;
    int rc = upload_alignments(ctx, soa, false);
    if (rc) return rc;
    // bam_iterator (SVIM_COLLECT.py:8-41): runs of consecutive records with the same read name; a read is analysed when it has
    // exactly one non-secondary non-supplementary record, mapped, MAPQ >= min_mapq (:108); its good supplementary records
    // (mapped, MAPQ >= min_mapq, :112) are both CIGAR-analysed and used as segments.
    const int64_t n = soa->n_aln;
    std::vector<uint32_t> info((size_t)n, 0), grp((size_t)n, 0), mem_off(1, 0), mem_idx;
    uint32_t g = 0;
    const int32_t min_mapq = ctx->params.min_mapq;
    for (int64_t i = 0; i < n;) {
        int64_t j = i, prim = -1; int n_prim = 0;
        while (j < n && soa->qname_id[j] == soa->qname_id[i]) {
            const uint16_t f = soa->flag[j];
            if (!(f & 0x100) && !(f & 0x800)) { ++n_prim; prim = j; }
            ++j;
        }
        const size_t first_mem = mem_idx.size();
        if (n_prim == 1 && !(soa->flag[prim] & 0x4) && (int32_t)soa->mapq[prim] >= min_mapq) {
            uint32_t slot = 1;
            for (int64_t k = i; k < j; ++k) {
                const uint16_t f = soa->flag[k];
                if ((f & 0x800) && !(f & 0x100) && !(f & 0x4) && (int32_t)soa->mapq[k] >= min_mapq) {
                    if (slot >= 0xFFF) { ctx->set_error(SVIMGPU_ERR_LIMIT, "more than 4094 supplementary records for one read"); return SVIMGPU_ERR_LIMIT; }
                    info[k] = 2u | (slot << 3); grp[k] = g; mem_idx.push_back((uint32_t)k); ++slot;
                }
            }
            const bool chain = mem_idx.size() > first_mem;
            info[prim] = 1u; grp[prim] = g;
            if (chain) { info[prim] |= 4u; for (size_t m = first_mem; m < mem_idx.size(); ++m) info[mem_idx[m]] |= 4u; }
        }
        mem_off.push_back((uint32_t)mem_idx.size());
        ++g; i = j;
    }
    cudaStream_t st = ctx->stream;
    SVIM_CUDA(ctx->d_qs_info.ensure((size_t)n * 4 + 4)); SVIM_CUDA(ctx->d_qs_grp.ensure((size_t)n * 4 + 4));
    SVIM_CUDA(ctx->d_qs_segsum.ensure((size_t)(n + 1) * sizeof(SegSum)));
    SVIM_CUDA(ctx->d_qs_mem_off.ensure(mem_off.size() * 4)); SVIM_CUDA(ctx->d_qs_mem_idx.ensure(mem_idx.size() * 4 + 4));
    if (n) {
        SVIM_CUDA(cudaMemcpyAsync(ctx->d_qs_info.p, info.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        SVIM_CUDA(cudaMemcpyAsync(ctx->d_qs_grp.p, grp.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_qs_mem_off.p, mem_off.data(), mem_off.size() * 4, cudaMemcpyHostToDevice, st));
    if (!mem_idx.empty()) SVIM_CUDA(cudaMemcpyAsync(ctx->d_qs_mem_idx.p, mem_idx.data(), mem_idx.size() * 4, cudaMemcpyHostToDevice, st));
    SVIM_CUDA(cudaStreamSynchronize(st));
    ctx->qs_mode = true;
    rc = svimgpu_collect(ctx, stats);
    ctx->qs_mode = false;
    ctx->have_soa = false; ctx->lazy_seq = false; ctx->h_seq = nullptr; ctx->h_seq_off = nullptr;
    if (!rc) { if (ctx->nccl_comm) { ctx->host_copy[0] = ctx->host_copy[1] = false; ctx->host_copy_pending = true; } else rc = start_host_copy(ctx); }
    return rc;
}

int svimgpu_fetch_signatures(svimgpu_ctx* ctx, int which, svim_sig* out_sigs, uint8_t* out_ins) {
    if (!ctx || which < 0 || which > 1) return SVIMGPU_ERR_ARG;
    if (!ctx->collected) { ctx->set_error(SVIMGPU_ERR_STATE, "collect has not run"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    SigSet& set = ctx->sets[which];
    if (set.n && out_sigs) SVIM_CUDA(cudaMemcpyAsync(out_sigs, set.recs.p, (size_t)set.n * sizeof(svim_sig), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ins) { int rc = copy_ins_to_host(ctx, set, out_ins, ctx->stream); if (rc) return rc; }
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

__global__ void k_max_ins_len(const svim_csig* c, uint32_t n, unsigned long long* out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = (k < n && c[k].type == SVIM_INS) ? c[k].seq_len : 0;
    for (int o = 16; o > 0; o >>= 1) { unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

static int finish_csig(svimgpu_ctx* ctx) {
    unsigned long long h = 0;
    SVIM_CUDA(ctx->d_part_stats.ensure(64 * 4));
    SVIM_CUDA(cudaMemsetAsync(ctx->d_part_stats.p, 0, 8, ctx->stream));
    if (ctx->n_csig) { ctx->launches++; k_max_ins_len<<<(uint32_t)((ctx->n_csig + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_csig.as<svim_csig>(), (uint32_t)ctx->n_csig,
                                                                                                   (unsigned long long*)ctx->d_part_stats.p); }
    SVIM_CUDA(cudaMemcpyAsync(&h, ctx->d_part_stats.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->cluster_max_ins_len = (int64_t)h;
    ctx->have_csig = true; ctx->clustered = false;
    return 0;
}

int svimgpu_use_collected(svimgpu_ctx* ctx, int which) {
    if (!ctx || which < 0 || which > 1) return SVIMGPU_ERR_ARG;
    if (!ctx->collected) { ctx->set_error(SVIMGPU_ERR_STATE, "collect has not run"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    timings_begin(ctx);
    SigSet& set = ctx->sets[which];
    ctx->n_csig = set.n;
    SVIM_CUDA(ctx->d_csig.ensure((size_t)(set.n + 1) * sizeof(svim_csig)));
    {
        StageTimer t(ctx, T_CSIG);
        if (set.n) { ctx->launches++; k_sig_to_csig<<<(uint32_t)((set.n + 255) / 256), 256, 0, ctx->stream>>>(set.recs.as<svim_sig>(), (uint32_t)set.n, ctx->d_rank.as<int32_t>(),
                                                                                           ctx->d_csig.as<svim_csig>()); }
    }
    if (set.segmented) { ctx->cluster_ins = nullptr; ctx->cluster_ins_bytes = 0; ctx->cluster_seg = which; }     // cluster_run pulls what its partitions need
    else { ctx->cluster_ins = set.ins.as<uint8_t>(); ctx->cluster_ins_bytes = set.ins_bytes; ctx->cluster_seg = -1; }
    ctx->cluster_rank_to_tid = ctx->d_rank_to_tid.as<int32_t>(); ctx->cluster_n_ranks = ctx->n_contigs;
    return finish_csig(ctx);
}

int svimgpu_set_signatures(svimgpu_ctx* ctx, int64_t n, const svim_csig* sigs, const uint8_t* ins_blob, int64_t ins_bytes,
                           const int32_t* rank_to_tid, int32_t n_ranks) {
    if (!ctx || n < 0 || (n > 0 && !sigs)) return SVIMGPU_ERR_ARG;
    if (n >= ((int64_t)1 << 31)) { ctx->set_error(SVIMGPU_ERR_LIMIT, "too many signatures"); return SVIMGPU_ERR_LIMIT; }
    cudaSetDevice(ctx->device);
    timings_begin(ctx);
    ctx->n_csig = n;
    SVIM_CUDA(ctx->d_csig.ensure((size_t)(n + 1) * sizeof(svim_csig)));
    if (n) SVIM_CUDA(cudaMemcpyAsync(ctx->d_csig.p, sigs, (size_t)n * sizeof(svim_csig), cudaMemcpyHostToDevice, ctx->stream));
    SVIM_CUDA(ctx->d_cins.ensure((size_t)ins_bytes + 16));
    if (ins_bytes && ins_blob) SVIM_CUDA(cudaMemcpyAsync(ctx->d_cins.p, ins_blob, (size_t)ins_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->cluster_ins = ctx->d_cins.as<uint8_t>(); ctx->cluster_ins_bytes = ins_bytes; ctx->cluster_seg = -1;
    ctx->cluster_rank_to_tid = nullptr; ctx->cluster_n_ranks = 0;
    if (rank_to_tid && n_ranks > 0) {
        SVIM_CUDA(ctx->d_user_rank_to_tid.ensure((size_t)n_ranks * 4));
        SVIM_CUDA(cudaMemcpyAsync(ctx->d_user_rank_to_tid.p, rank_to_tid, (size_t)n_ranks * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->cluster_rank_to_tid = ctx->d_user_rank_to_tid.as<int32_t>(); ctx->cluster_n_ranks = n_ranks;
    }
    return finish_csig(ctx);
}

int svimgpu_cluster(svimgpu_ctx* ctx, svim_cluster_stats* stats) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = cluster_run(ctx, stats, 0, 1);
    timings_end(ctx);
    return rc;
}

int svimgpu_partition(svimgpu_ctx* ctx, int64_t* n_partitions) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = cluster_run(ctx, nullptr, 0, 1, true);
    timings_end(ctx);
    if (!rc && n_partitions) *n_partitions = ctx->n_partitions;
    return rc;
}

int svimgpu_cluster_sharded(svimgpu_ctx* ctx, svim_cluster_stats* stats) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    int rc = cluster_run(ctx, stats, ctx->rank, ctx->nranks);
    timings_end(ctx);
    return rc;
}

int svimgpu_fetch_clusters(svimgpu_ctx* ctx, svim_cluster* clusters, uint32_t* members) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    if (!ctx->clustered) { ctx->set_error(SVIMGPU_ERR_STATE, "cluster has not run"); return SVIMGPU_ERR_STATE; }
    if (clusters && ctx->n_clusters_host) memcpy(clusters, ctx->h_clusters.p, (size_t)ctx->n_clusters_host * sizeof(svim_cluster));
    if (members && ctx->n_members_host) memcpy(members, ctx->h_members.p, (size_t)ctx->n_members_host * 4);
    return 0;
}

int svimgpu_clusters_host(svimgpu_ctx* ctx, const svim_cluster** clusters, const uint32_t** members) {
    if (!ctx || !clusters || !members) return SVIMGPU_ERR_ARG;
    if (!ctx->clustered) { ctx->set_error(SVIMGPU_ERR_STATE, "cluster has not run"); return SVIMGPU_ERR_STATE; }
    *clusters = ctx->n_clusters_host ? ctx->h_clusters.as<svim_cluster>() : nullptr;
    *members = ctx->n_members_host ? ctx->h_members.as<uint32_t>() : nullptr;
    return 0;
}

int svimgpu_fetch_partitions(svimgpu_ctx* ctx, int64_t* n_partitions, uint32_t* order, uint32_t* part_off) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    if (!ctx->clustered) { ctx->set_error(SVIMGPU_ERR_STATE, "cluster has not run"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    if (n_partitions) *n_partitions = ctx->n_partitions;
    if (order && ctx->n_csig) SVIM_CUDA(cudaMemcpy(order, ctx->d_order.p, (size_t)ctx->n_csig * 4, cudaMemcpyDeviceToHost));
    if (part_off) {
        if (ctx->n_partitions) SVIM_CUDA(cudaMemcpy(part_off, ctx->d_part_off.p, (size_t)ctx->n_partitions * 4, cudaMemcpyDeviceToHost));
        part_off[ctx->n_partitions] = (uint32_t)ctx->n_csig;
    }
    return 0;
}

// ---- micro entry points ------------------------------------------------------------------------
int svimgpu_genotype(svimgpu_ctx* ctx, int32_t type, const svim_geno_params* gp, int64_t n, const svim_geno_cand* cands, const uint32_t* variant_qname_ids,
                     int64_t n_variant_ids, const int64_t* contig_lengths, int32_t n_contigs, svim_geno_result* out) {
    if (!ctx || !gp || n < 0 || n_contigs <= 0 || !contig_lengths || (n && (!cands || !out)) || (n_variant_ids && !variant_qname_ids)) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    timings_begin(ctx);
    int rc = genotype_run(ctx, type, gp, n, cands, variant_qname_ids, n_variant_ids, contig_lengths, n_contigs, out);
    timings_end(ctx);
    return rc;
}

int svimgpu_closest_source(svimgpu_ctx* ctx, int64_t n_a, const int64_t* a_start, const int64_t* a_end, int64_t n_b, const int64_t* b_start,
                           const int64_t* b_end, double position_distance_normalizer, int64_t* out_index, double* out_distance) {
    if (!ctx || n_a < 0 || n_b < 0 || (n_a && (!a_start || !a_end || !out_index || !out_distance)) || (n_b && (!b_start || !b_end))) return SVIMGPU_ERR_ARG;
    if (n_a == 0) return 0;
    cudaSetDevice(ctx->device);
    timings_begin(ctx);
    int rc = closest_source_run(ctx, n_a, a_start, a_end, n_b, b_start, b_end, position_distance_normalizer, out_index, out_distance);
    timings_end(ctx);
    return rc;
}

int svimgpu_cigar_indel(svimgpu_ctx* ctx, const uint32_t* cigar, int64_t n, int32_t min_len, int64_t* out, int64_t out_cap, int64_t* n_out) {
    if (!ctx || n < 0 || !n_out) return SVIMGPU_ERR_ARG;
    // one synthetic record at position 0 on contig 0 through the real scan kernel
    svim_params saved = ctx->params;
    ctx->params.min_sv_size = min_len; ctx->params.min_mapq = 0; ctx->params.all_bnds = 0;
    int32_t tid = 0, pos = 0, l_seq = 0; uint16_t flag = 0; uint8_t mapq = 60; uint32_t nc = (uint32_t)n, sal = 0, qid = 0; uint64_t zero = 0;
    std::vector<uint32_t> padded((size_t)((n + 3) & ~3ll) + 4, 0);
    if (n) memcpy(padded.data(), cigar, (size_t)n * 4);
    uint8_t dummy = 0;
    svim_aln_soa s{1, &tid, &pos, &flag, &mapq, &nc, &zero, &l_seq, &zero, &zero, &sal, &qid, padded.data(), (int64_t)padded.size(), &dummy, 0, &dummy, 0};
    int rc = 0;
    if (ctx->n_contigs == 0) { const char nm[] = "c"; int32_t off[2] = {0, 1}; rc = svimgpu_set_contigs(ctx, 1, nm, off); }
    svim_collect_stats st;
    if (!rc) rc = svimgpu_collect_host(ctx, &s, &st);
    ctx->params = saved;
    if (rc) return rc;
    std::vector<svim_sig> sigs((size_t)st.n_signatures);
    rc = svimgpu_fetch_signatures(ctx, 0, sigs.data(), nullptr);
    if (rc) return rc;
    *n_out = st.n_signatures;
    // pos_read is not part of a signature; recover it for INS from the clamp-free source offset is impossible with l_seq=0,
    // so report (pos_ref, -1, len, type); the Python test recomputes pos_read from the CIGAR prefix.
    for (int64_t k = 0; k < st.n_signatures && k < out_cap; ++k) {
        out[4 * k] = sigs[k].start; out[4 * k + 1] = -1; out[4 * k + 2] = sigs[k].end - sigs[k].start; out[4 * k + 3] = sigs[k].type == SVIM_INS ? 1 : 2;
    }
    return 0;
}

int svimgpu_edit_distance(svimgpu_ctx* ctx, int64_t n_pairs, const uint8_t* blob, const int64_t* a_off, const int32_t* a_len, const int64_t* b_off,
                          const int32_t* b_len, int32_t* out) {
    if (!ctx || n_pairs < 0) return SVIMGPU_ERR_ARG;
    if (n_pairs == 0) return 0;
    if (n_pairs >= 0x7fffffff) return SVIMGPU_ERR_LIMIT;
    cudaSetDevice(ctx->device);
    int64_t blob_bytes = 0, maxlen = 16;
    // same scheduling as the pipeline (k_ins_pairs): banded shape if the band policy gives a smaller one, else the pattern's bin
    std::vector<uint32_t> lists[MYERS_LISTS];
    MyersPlan pl; memset(&pl, 0, sizeof(pl));
    for (int64_t i = 0; i < n_pairs; ++i) {
        blob_bytes = std::max<int64_t>(blob_bytes, std::max(a_off[i] + a_len[i], b_off[i] + b_len[i]));
        const int64_t m = std::max(a_len[i], b_len[i]), n = std::min(a_len[i], b_len[i]);
        maxlen = std::max<int64_t>(maxlen, m);
        const int tq = ctx->myers_tpp ? tpp_bucket_of(tpp_plan(m, n, ctx->myers_band_num, ctx->myers_band_add).B) : -1;
        const int band = tq >= 0 ? -1 : myers_band_bin(m, n, ctx->myers_band_num, ctx->myers_band_add);
        if (tq >= 0) { lists[2 * MYERS_BINS + tq].push_back((uint32_t)i); pl.retry_cap[myers_bin_of(m)]++; }
        else if (band >= 0) { lists[MYERS_BINS + band].push_back((uint32_t)i); pl.retry_cap[myers_bin_of(m)]++; }
        else lists[myers_bin_of(m)].push_back((uint32_t)i);
    }
    maxlen = (maxlen + 15) & ~15ll;
    std::vector<uint32_t> flat; flat.reserve((size_t)n_pairs);
    for (int q = 0; q < MYERS_LISTS; ++q) { pl.off[q] = (uint32_t)flat.size(); pl.cnt[q] = (uint32_t)lists[q].size(); flat.insert(flat.end(), lists[q].begin(), lists[q].end()); }
    DevBuf d_blob, d_ao, d_al, d_bo, d_bl, d_out, d_ctl, d_list, d_fb, d_retry, d_misc, d_codes;
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    chk(d_blob.ensure((size_t)blob_bytes + 16)); chk(d_codes.ensure((size_t)blob_bytes + 16)); chk(d_ao.ensure((size_t)n_pairs * 8)); chk(d_al.ensure((size_t)n_pairs * 4)); chk(d_bo.ensure((size_t)n_pairs * 8));
    chk(d_bl.ensure((size_t)n_pairs * 4)); chk(d_out.ensure((size_t)n_pairs * 4)); chk(d_ctl.ensure(MYERS_CTL_N * 4)); chk(d_list.ensure((size_t)n_pairs * 4));
    chk(d_fb.ensure((size_t)n_pairs * sizeof(MyersWork))); chk(d_retry.ensure((size_t)n_pairs * sizeof(MyersWork))); chk(d_misc.ensure(64));
    uint32_t h_err = 0;
    if (e == cudaSuccess) {
        chk(cudaMemcpyAsync(d_blob.p, blob, (size_t)blob_bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (blob_bytes > 0) k_tpp_encode<<<(unsigned)(((size_t)blob_bytes + 16 * 256 - 1) / (16 * 256)), 256, 0, ctx->stream>>>(d_blob.as<uint8_t>(), blob_bytes, d_codes.as<uint8_t>());
        chk(cudaMemcpyAsync(d_ao.p, a_off, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, ctx->stream));
        chk(cudaMemcpyAsync(d_al.p, a_len, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
        chk(cudaMemcpyAsync(d_bo.p, b_off, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, ctx->stream));
        chk(cudaMemcpyAsync(d_bl.p, b_len, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
        chk(cudaMemcpyAsync(d_list.p, flat.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
        chk(cudaMemsetAsync(d_ctl.p, 0, MYERS_CTL_N * 4, ctx->stream)); chk(cudaMemsetAsync(d_misc.p, 0, 64, ctx->stream));
        MyersArgs ma; memset(&ma, 0, sizeof(ma));
        ma.ed_out = d_out.as<int32_t>(); ma.maxlen = maxlen; ma.fallback = d_fb.as<MyersWork>();
        ma.cells = (unsigned long long*)d_misc.p; ma.err = (uint32_t*)d_misc.p + 4;
        ma.band_num = ctx->myers_band_num; ma.band_add = ctx->myers_band_add;
        ma.str_codes = d_codes.as<uint8_t>(); ma.ins_base = d_blob.as<uint8_t>();
        StringPairs sp{d_blob.as<uint8_t>(), d_ao.as<int64_t>(), d_al.as<int32_t>(), d_bo.as<int64_t>(), d_bl.as<int32_t>(), nullptr};
        chk(myers_run_plan<true>(ctx, pl, ma, sp, nullptr, d_list.as<uint32_t>(), d_retry.as<MyersWork>(), d_ctl.as<uint32_t>(), maxlen, 148));
        chk(cudaMemcpyAsync(out, d_out.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        chk(cudaMemcpyAsync(&h_err, (uint32_t*)d_misc.p + 4, 4, cudaMemcpyDeviceToHost, ctx->stream));
        chk(cudaStreamSynchronize(ctx->stream));
    }
    DevBuf* all[] = {&d_blob, &d_ao, &d_al, &d_bo, &d_bl, &d_out, &d_ctl, &d_list, &d_fb, &d_retry, &d_misc, &d_codes};
    for (DevBuf* b : all) b->release();
    if (e != cudaSuccess) { ctx->set_error(SVIMGPU_ERR_CUDA, "edit_distance: %s", cudaGetErrorString(e)); return SVIMGPU_ERR_CUDA; }
    if (h_err) { ctx->set_error(SVIMGPU_ERR_LIMIT, "edit_distance: internal length bound exceeded (%u)", h_err); return SVIMGPU_ERR_LIMIT; }
    return 0;
}

// linkage + fcluster on an explicit condensed matrix, through the same device functions as k_linkage
__global__ void __launch_bounds__(32) k_linkage_raw(const double* condensed, int m, double t, double* Z, int32_t* T) {
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    const int lane = threadIdx.x;
    PartSmem s = carve(smem_raw2, m < 2 ? 2 : m);
    const int np = m * (m - 1) / 2;
    for (int q = lane; q < np; q += 32) s.D[q] = condensed[q];
    __syncwarp();
    nn_chain_warp(s, m, lane);
    if (lane == 0) {
        fcluster_from_chain(m, s.ls, s.ux, s.uy, s.ud, t, s.T);
        for (int k = 0; k < m; ++k) T[k] = s.T[k];
        // Z with scipy's size column: recompute sizes from the relabelled tree
        for (int k = 0; k < m - 1; ++k) {
            int a = s.ls.zx[k], b = s.ls.zy[k];
            double sa = a < m ? 1.0 : Z[4 * (a - m) + 3], sb = b < m ? 1.0 : Z[4 * (b - m) + 3];
            Z[4 * k] = a; Z[4 * k + 1] = b; Z[4 * k + 2] = s.ls.zd[k]; Z[4 * k + 3] = sa + sb;
        }
    }
}

int svimgpu_linkage_average(svimgpu_ctx* ctx, const double* condensed, int32_t m, double t, double* Z, int32_t* T) {
    if (!ctx || m < 2 || m > 100 || !condensed || !Z || !T) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    const size_t np = (size_t)m * (m - 1) / 2;
    DevBuf d_c, d_z, d_t;
    SVIM_CUDA(d_c.ensure(np * 8)); SVIM_CUDA(d_z.ensure((size_t)(m - 1) * 32)); SVIM_CUDA(d_t.ensure((size_t)m * 4));
    SVIM_CUDA(cudaMemcpyAsync(d_c.p, condensed, np * 8, cudaMemcpyHostToDevice, ctx->stream));
    size_t sm = part_smem_bytes(m);
    SVIM_CUDA(cudaFuncSetAttribute(k_linkage_raw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(sm, 1024)));
    { ctx->launches++; k_linkage_raw<<<1, 32, sm, ctx->stream>>>(d_c.as<double>(), m, t, d_z.as<double>(), d_t.as<int32_t>()); }
    SVIM_CUDA(cudaGetLastError());
    SVIM_CUDA(cudaMemcpyAsync(Z, d_z.p, (size_t)(m - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaMemcpyAsync(T, d_t.p, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    d_c.release(); d_z.release(); d_t.release();
    return 0;
}

int svimgpu_sample_indices(const int64_t* sizes, int64_t n_sizes, int32_t* out) {
    if (!sizes || !out) return SVIMGPU_ERR_ARG;
    PyRandom rng; rng.seed_int(1524);
    int64_t k = 0;
    for (int64_t i = 0; i < n_sizes; ++i)
        if (sizes[i] > 100) { rng.sample100((uint64_t)sizes[i], out + 100 * k); ++k; }
    return 0;
}

int svimgpu_last_timings(svimgpu_ctx* ctx, double* ms, int32_t cap, int32_t* n) {
    if (!ctx || !ms || !n) return SVIMGPU_ERR_ARG;
    *n = T_N;
    for (int i = 0; i < T_N && i < cap; ++i) ms[i] = ctx->ms[i];
    return 0;
}

int svimgpu_timer_start(svimgpu_ctx* ctx) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    SVIM_CUDA(cudaEventRecord(ctx->user_ev[0], ctx->stream));
    return 0;
}

int svimgpu_timer_stop(svimgpu_ctx* ctx, double* ms) {
    if (!ctx || !ms) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    SVIM_CUDA(cudaEventRecord(ctx->user_ev[1], ctx->stream));
    SVIM_CUDA(cudaEventSynchronize(ctx->user_ev[1]));
    float f = 0;
    SVIM_CUDA(cudaEventElapsedTime(&f, ctx->user_ev[0], ctx->user_ev[1]));
    *ms = f;
    return 0;
}

int64_t svimgpu_launch_count(svimgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

const char* svimgpu_timing_name(int32_t i) { return (i >= 0 && i < T_N) ? k_stage_names[i] : ""; }

}  // extern "C"

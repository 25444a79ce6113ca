// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The path shards twice: COLLECT over contiguous record ranges, CLUSTER over
// contiguous partition ranges.  Two variable-size all-gathers are the only collectives:
//   1. signature records (+ INS sequences) after COLLECT — signatures land at loci unrelated to
//      the emitting record's shard, so every rank needs the full list to form partitions;
//   2. cluster records (+ member indices) after consolidation.
// allgatherv = grouped ncclBroadcast, one per rank, into the rank's slice of the result.
#pragma once
#include <nccl.h>
#include "ctx.cuh"

#define SVIM_NCCL(call)                                                                                   \
    do {                                                                                                  \
        ncclResult_t _r = (call);                                                                         \
        if (_r != ncclSuccess) {                                                                          \
            ctx->set_error(SVIMGPU_ERR_NCCL, "%s failed: %s", #call, ncclGetErrorString(_r));             \
            return SVIMGPU_ERR_NCCL;                                                                      \
        }                                                                                                 \
    } while (0)

static void nccl_teardown(svimgpu_ctx* ctx) {
    if (ctx->nccl_comm) { ncclCommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
}

// all-gather `count` int64 values per rank (host in/out through a device bounce buffer)
static int nccl_allgather_i64(svimgpu_ctx* ctx, const int64_t* mine, int count, int64_t* all) {
    SVIM_CUDA(ctx->d_xchg[3].ensure((size_t)(ctx->nranks + 1) * count * 8));
    int64_t* d = ctx->d_xchg[3].as<int64_t>();
    SVIM_CUDA(cudaMemcpyAsync(d + (size_t)ctx->nranks * count, mine, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    SVIM_NCCL(ncclAllGather(d + (size_t)ctx->nranks * count, d, (size_t)count, ncclInt64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    SVIM_CUDA(cudaMemcpyAsync(all, d, (size_t)ctx->nranks * count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// dst[off[r] .. off[r]+bytes[r]) <- rank r's src, for every r
static int nccl_allgatherv_bytes(svimgpu_ctx* ctx, const void* src, const std::vector<int64_t>& bytes, const std::vector<int64_t>& off, uint8_t* dst) {
    SVIM_NCCL(ncclGroupStart());
    for (int r = 0; r < ctx->nranks; ++r) {
        if (bytes[r] == 0) continue;
        const void* s = (r == ctx->rank) ? src : (const void*)(dst + off[r]);
        SVIM_NCCL(ncclBroadcast(s, dst + off[r], (size_t)bytes[r], ncclUint8, r, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    }
    SVIM_NCCL(ncclGroupEnd());
    return 0;
}

__global__ void k_rebase_sigs(svim_sig* s, uint32_t n, uint32_t aln_add, uint64_t seq_add) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    s[k].aln_idx += aln_add;
    if (s[k].type == SVIM_INS) s[k].seq_off += seq_add;
}

__global__ void k_rebase_clusters(svim_cluster* c, uint32_t n, uint32_t mem_add) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) c[k].member_off += mem_add;
}

static int cluster_exchange(svimgpu_ctx* ctx, uint32_t* n_clusters, uint32_t* n_members) {
    if (!ctx->nccl_comm) { ctx->set_error(SVIMGPU_ERR_STATE, "svimgpu_comm_init not called"); return SVIMGPU_ERR_STATE; }
    StageTimer t(ctx, T_EXCHANGE);
    const int R = ctx->nranks;
    int64_t mine[2] = {(int64_t)*n_clusters, (int64_t)*n_members};
    std::vector<int64_t> all(2 * R);
    int rc = nccl_allgather_i64(ctx, mine, 2, all.data()); if (rc) return rc;
    std::vector<int64_t> cb(R), co(R), mb(R), mo(R);
    int64_t ct = 0, mt = 0;
    for (int r = 0; r < R; ++r) { co[r] = ct * (int64_t)sizeof(svim_cluster); cb[r] = all[2 * r] * (int64_t)sizeof(svim_cluster); mo[r] = mt * 4; mb[r] = all[2 * r + 1] * 4; ct += all[2 * r]; mt += all[2 * r + 1]; }
    // rebase this rank's member offsets to the global member array, then gather both arrays
    const uint32_t my_mem_base = (uint32_t)(mo[ctx->rank] / 4);
    if (*n_clusters) { ctx->launches++; k_rebase_clusters<<<(*n_clusters + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_clusters.as<svim_cluster>(), *n_clusters, my_mem_base); }
    SVIM_CUDA(ctx->d_xchg[0].ensure((size_t)(ct + 1) * sizeof(svim_cluster))); SVIM_CUDA(ctx->d_xchg[1].ensure((size_t)(mt + 1) * 4));
    rc = nccl_allgatherv_bytes(ctx, ctx->d_clusters.p, cb, co, ctx->d_xchg[0].as<uint8_t>()); if (rc) return rc;
    rc = nccl_allgatherv_bytes(ctx, ctx->d_members.p, mb, mo, ctx->d_xchg[1].as<uint8_t>()); if (rc) return rc;
    std::swap(ctx->d_clusters, ctx->d_xchg[0]); std::swap(ctx->d_members, ctx->d_xchg[1]);
    SVIM_CUDA(ctx->d_clusters_sorted.ensure((size_t)(ct + 1) * sizeof(svim_cluster)));
    *n_clusters = (uint32_t)ct; *n_members = (uint32_t)mt;
    return 0;
}

extern "C" {

static int start_host_copy(svimgpu_ctx* ctx);   // api.cu

int svimgpu_nccl_unique_id(uint8_t* id_bytes) {
    if (!id_bytes) return SVIMGPU_ERR_ARG;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return SVIMGPU_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id_bytes, &id, 128);
    return 0;
}

int svimgpu_comm_init(svimgpu_ctx* ctx, int nranks, int rank, const uint8_t* id_bytes) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || !id_bytes) return SVIMGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    nccl_teardown(ctx);
    ncclUniqueId id; memcpy(&id, id_bytes, 128);
    ncclComm_t comm;
    SVIM_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm; ctx->nranks = nranks; ctx->rank = rank;
    return 0;
}

int svimgpu_exchange_signatures(svimgpu_ctx* ctx, uint32_t aln_base, svim_collect_stats* stats) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    if (!ctx->collected) { ctx->set_error(SVIMGPU_ERR_STATE, "collect has not run"); return SVIMGPU_ERR_STATE; }
    if (!ctx->nccl_comm) { ctx->set_error(SVIMGPU_ERR_STATE, "svimgpu_comm_init not called"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    // the local lists are replaced by the global ones: a host copy in flight (collect_host) is dropped and restarted at the end
    const bool want_host = ctx->host_copy[0] || ctx->host_copy[1];
    if (want_host) { cudaStreamSynchronize(ctx->copy_stream); ctx->host_copy[0] = ctx->host_copy[1] = false; }
    timings_begin(ctx);
    const int R = ctx->nranks;
    {
        StageTimer t(ctx, T_EXCHANGE);
        int64_t mine[12] = {ctx->sets[0].n, ctx->sets[0].ins_bytes, ctx->sets[1].n, ctx->sets[1].ins_bytes, (int64_t)aln_base,
                            ctx->cstats.n_sa_bad_fields, ctx->cstats.n_no_read_length, ctx->cstats.n_primaries, ctx->cstats.n_data_errors, 0, 0, 0};
        std::vector<int64_t> all(12 * R);
        int rc = nccl_allgather_i64(ctx, mine, 12, all.data()); if (rc) return rc;
        for (int w = 0; w < 2; ++w) {
            SigSet& set = ctx->sets[w];
            std::vector<int64_t> rb(R), ro(R), ib(R), io(R);
            int64_t nt = 0, it = 0;
            for (int r = 0; r < R; ++r) {
                ro[r] = nt * (int64_t)sizeof(svim_sig); rb[r] = all[12 * r + 2 * w] * (int64_t)sizeof(svim_sig);
                io[r] = it; ib[r] = all[12 * r + 2 * w + 1];
                nt += all[12 * r + 2 * w]; it += all[12 * r + 2 * w + 1];
            }
            // make local records global before sending: record index += aln_base, INS offset += blob base
            if (set.n) { ctx->launches++; k_rebase_sigs<<<(uint32_t)((set.n + 255) / 256), 256, 0, ctx->stream>>>(set.recs.as<svim_sig>(), (uint32_t)set.n, aln_base, (uint64_t)io[ctx->rank]); }
            SVIM_CUDA(ctx->d_xchg[0].ensure((size_t)(nt + 1) * sizeof(svim_sig))); SVIM_CUDA(ctx->d_xchg[1].ensure((size_t)it + 16));
            rc = nccl_allgatherv_bytes(ctx, set.recs.p, rb, ro, ctx->d_xchg[0].as<uint8_t>()); if (rc) return rc;
            rc = nccl_allgatherv_bytes(ctx, set.ins.p, ib, io, ctx->d_xchg[1].as<uint8_t>()); if (rc) return rc;
            SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
            std::swap(set.recs, ctx->d_xchg[0]); std::swap(set.ins, ctx->d_xchg[1]);
            set.n = nt; set.ins_bytes = it;
        }
        svim_collect_stats& s = ctx->cstats;
        s.n_signatures = ctx->sets[0].n; s.ins_bytes = ctx->sets[0].ins_bytes; s.n_twin_signatures = ctx->sets[1].n; s.twin_ins_bytes = ctx->sets[1].ins_bytes;
        s.n_sa_bad_fields = s.n_no_read_length = s.n_primaries = s.n_data_errors = 0;
        for (int r = 0; r < R; ++r) { s.n_sa_bad_fields += all[12 * r + 5]; s.n_no_read_length += all[12 * r + 6]; s.n_primaries += all[12 * r + 7]; s.n_data_errors += all[12 * r + 8]; }
    }
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    timings_end(ctx);
    if (stats) *stats = ctx->cstats;
    return want_host ? start_host_copy(ctx) : 0;
}

int svimgpu_barrier_max(svimgpu_ctx* ctx, double* value) {
    if (!ctx || !value) return SVIMGPU_ERR_ARG;
    if (!ctx->nccl_comm) return 0;
    cudaSetDevice(ctx->device);
    SVIM_CUDA(ctx->d_xchg[2].ensure(16));
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_xchg[2].p, value, 8, cudaMemcpyHostToDevice, ctx->stream));
    SVIM_NCCL(ncclAllReduce(ctx->d_xchg[2].p, ctx->d_xchg[2].p, 1, ncclDouble, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    SVIM_CUDA(cudaMemcpyAsync(value, ctx->d_xchg[2].p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"

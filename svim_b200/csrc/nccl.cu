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
    for (int w = 0; w < 2; ++w) {
        for (PeerMap& m : ctx->peer_map[w]) if (m.p) cudaIpcCloseMemHandle(m.p);
        ctx->peer_map[w].clear();
        ctx->sets[w].segmented = false; ctx->sets[w].seg_ptr.clear(); ctx->sets[w].seg_base.clear();
    }
    ctx->peer_failed = false;
    if (ctx->nccl_comm) { ncclCommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
}

// all-gather `count` int64 values per rank (host in/out through a device bounce buffer)
static int nccl_allgather_i64(svimgpu_ctx* ctx, const int64_t* mine, int count, int64_t* all) {
    SVIM_CUDA(ctx->d_xchg[5].ensure((size_t)(ctx->nranks + 1) * count * 8));
    int64_t* d = ctx->d_xchg[5].as<int64_t>();
    SVIM_CUDA(cudaMemcpyAsync(d + (size_t)ctx->nranks * count, mine, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    SVIM_NCCL(ncclAllGather(d + (size_t)ctx->nranks * count, d, (size_t)count, ncclInt64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    SVIM_CUDA(cudaMemcpyAsync(all, d, (size_t)ctx->nranks * count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Variable-size all-gather of K byte arrays per rank with ONE collective: the K arrays of a rank are packed back to back into a
// send slot padded to the largest rank's total, one ncclAllGather moves all slots (NCCL's tuned ring / NVLS path instead of R
// grouped broadcasts), and device-to-device copies lay every array out contiguously in rank order.
//   bytes[r*K + k] : size of array k on rank r (from the count all-gather)     src[k] : this rank's arrays
//   dst[k]         : receives rank 0's array k, rank 1's array k, ...  (caller sized it)
static int nccl_allgatherv_packed(svimgpu_ctx* ctx, int K, const void* const* src, const int64_t* bytes, uint8_t* const* dst) {
    const int R = ctx->nranks;
    auto pad = [](int64_t x) { return (x + 15) & ~15ll; };
    int64_t slot = 16;
    for (int r = 0; r < R; ++r) { int64_t t = 0; for (int k = 0; k < K; ++k) t += pad(bytes[r * K + k]); slot = std::max(slot, t); }
    SVIM_CUDA(ctx->d_xchg[4].ensure((size_t)slot * (R + 1) + 64));
    uint8_t* recv = ctx->d_xchg[4].as<uint8_t>();
    uint8_t* send = recv + (size_t)slot * R;
    int64_t at = 0;
    for (int k = 0; k < K; ++k) {
        const int64_t b = bytes[ctx->rank * K + k];
        if (b) SVIM_CUDA(cudaMemcpyAsync(send + at, src[k], (size_t)b, cudaMemcpyDeviceToDevice, ctx->stream));
        at += pad(b);
    }
    SVIM_NCCL(ncclAllGather(send, recv, (size_t)slot, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    std::vector<int64_t> out_at(K, 0);
    for (int r = 0; r < R; ++r) {
        int64_t in_at = 0;
        for (int k = 0; k < K; ++k) {
            const int64_t b = bytes[r * K + k];
            if (b) SVIM_CUDA(cudaMemcpyAsync(dst[k] + out_at[k], recv + (size_t)slot * r + in_at, (size_t)b, cudaMemcpyDeviceToDevice, ctx->stream));
            in_at += pad(b); out_at[k] += b;
        }
    }
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

// all-gather-v of this rank's cluster records + member indices; `local_status` != 0 (this rank failed before the exchange) is
// agreed on first, so that either every rank enters the payload gather or none does
static int cluster_exchange(svimgpu_ctx* ctx, uint32_t* n_clusters, uint32_t* n_members, int local_status) {
    if (!ctx->nccl_comm) { if (local_status) return local_status; ctx->set_error(SVIMGPU_ERR_STATE, "svimgpu_comm_init not called"); return SVIMGPU_ERR_STATE; }
    StageTimer t(ctx, T_EXCHANGE);
    const int R = ctx->nranks;
    int64_t mine[3] = {(int64_t)*n_clusters, (int64_t)*n_members, (int64_t)local_status};
    std::vector<int64_t> all(3 * R);
    int rc = nccl_allgather_i64(ctx, mine, 3, all.data()); if (rc) return local_status ? local_status : rc;
    for (int r = 0; r < R; ++r)
        if (all[3 * r + 2]) {
            if (!local_status) ctx->set_error(SVIMGPU_ERR_PEER, "cluster: rank %d failed with status %lld; no rank keeps a result", r, (long long)all[3 * r + 2]);
            return local_status ? local_status : SVIMGPU_ERR_PEER;
        }
    std::vector<int64_t> bytes(2 * R);
    int64_t ct = 0, mt = 0, my_mem_base = 0;
    for (int r = 0; r < R; ++r) {
        bytes[2 * r] = all[3 * r] * (int64_t)sizeof(svim_cluster); bytes[2 * r + 1] = all[3 * r + 1] * 4;
        if (r < ctx->rank) my_mem_base += all[3 * r + 1];
        ct += all[3 * r]; mt += all[3 * r + 1];
    }
    // rebase this rank's member offsets to the global member array, then gather both arrays
    if (*n_clusters) { ctx->launches++; k_rebase_clusters<<<(*n_clusters + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_clusters.as<svim_cluster>(), *n_clusters, (uint32_t)my_mem_base); }
    SVIM_CUDA(ctx->d_xchg[0].ensure((size_t)(ct + 1) * sizeof(svim_cluster))); SVIM_CUDA(ctx->d_xchg[1].ensure((size_t)(mt + 1) * 4));
    const void* src[2] = {ctx->d_clusters.p, ctx->d_members.p};
    uint8_t* dst[2] = {ctx->d_xchg[0].as<uint8_t>(), ctx->d_xchg[1].as<uint8_t>()};
    rc = nccl_allgatherv_packed(ctx, 2, src, bytes.data(), dst); if (rc) return rc;
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
    if (const char* v = getenv("SVIM_PEER_INS")) ctx->peer_ins = atoi(v) != 0; else ctx->peer_ins = true;
    // Peer-mapped INS blobs need CUDA IPC between the ranks' processes.  Try it once on a probe buffer and let the ranks agree: if any
    // mapping fails (no P2P path, IPC not permitted in this container) every rank gathers the blobs instead.
    if (ctx->peer_ins && nranks > 1) {
        DevBuf probe;
        int64_t mine[9] = {0};
        cudaIpcMemHandle_t hd;
        if (probe.ensure(256) == cudaSuccess && cudaIpcGetMemHandle(&hd, probe.p) == cudaSuccess) { memcpy(mine + 1, &hd, 64); mine[0] = 1; }
        else cudaGetLastError();
        std::vector<int64_t> all((size_t)9 * nranks);
        int rc = nccl_allgather_i64(ctx, mine, 9, all.data()); if (rc) { probe.release(); return rc; }
        int64_t ok = 1;
        std::vector<void*> opened;
        for (int r = 0; r < nranks; ++r) {
            if (!all[(size_t)9 * r]) { ok = 0; continue; }
            if (r == rank) continue;
            cudaIpcMemHandle_t h; memcpy(&h, &all[(size_t)9 * r + 1], 64);
            void* p = nullptr; uint32_t word = 0;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; continue; }
            opened.push_back(p);
            if (cudaMemcpy(&word, p, 4, cudaMemcpyDefault) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
        std::vector<int64_t> oks((size_t)nranks);
        rc = nccl_allgather_i64(ctx, &ok, 1, oks.data());        // also: every rank is done with the probes before any is freed
        for (void* p : opened) cudaIpcCloseMemHandle(p);
        if (rc) { probe.release(); return rc; }
        for (int r = 0; r < nranks; ++r) if (!oks[r]) ctx->peer_ins = false;
        int64_t done = 1;
        rc = nccl_allgather_i64(ctx, &done, 1, oks.data());       // mappings closed everywhere
        probe.release();
        if (rc) return rc;
    }
    return 0;
}

int svimgpu_peer_ins_active(svimgpu_ctx* ctx) { return ctx && ctx->nccl_comm && ctx->nranks > 1 && ctx->peer_ins ? 1 : 0; }

int svimgpu_exchange_signatures(svimgpu_ctx* ctx, uint32_t aln_base, svim_collect_stats* stats) {
    if (!ctx) return SVIMGPU_ERR_ARG;
    if (!ctx->collected) { ctx->set_error(SVIMGPU_ERR_STATE, "collect has not run"); return SVIMGPU_ERR_STATE; }
    if (!ctx->nccl_comm) { ctx->set_error(SVIMGPU_ERR_STATE, "svimgpu_comm_init not called"); return SVIMGPU_ERR_STATE; }
    cudaSetDevice(ctx->device);
    // the local lists are replaced by the global ones: a host copy in flight (collect_host) is dropped and restarted at the end
    const bool want_host = ctx->host_copy[0] || ctx->host_copy[1] || ctx->host_copy_pending;
    if (ctx->host_copy[0] || ctx->host_copy[1]) { cudaStreamSynchronize(ctx->copy_stream); ctx->host_copy[0] = ctx->host_copy[1] = false; }
    ctx->host_copy_pending = false;
    timings_begin(ctx);
    const int R = ctx->nranks;
    {
        StageTimer t(ctx, T_EXCHANGE);
        // counts, and (peer mode) the CUDA-IPC handle of each list's INS blob: the bytes are not moved here.  Every rank later reads
        // the sequences its own partitions compare straight from the rank that holds them (cluster.cu, k_shard_ins_copy).
        constexpr int NW = 12 + 2 * 8;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
        const bool peer = ctx->peer_ins;
        int64_t mine[NW] = {ctx->sets[0].n, ctx->sets[0].ins_bytes, ctx->sets[1].n, ctx->sets[1].ins_bytes, (int64_t)aln_base,
                            ctx->cstats.n_sa_bad_fields, ctx->cstats.n_no_read_length, ctx->cstats.n_primaries, ctx->cstats.n_data_errors, 0, 0, 0};
        if (peer)
            for (int w = 0; w < 2; ++w)
                if (ctx->sets[w].ins_bytes > 0 && ctx->sets[w].ins.p) {
                    cudaIpcMemHandle_t hd;
                    if (cudaIpcGetMemHandle(&hd, ctx->sets[w].ins.p) == cudaSuccess) { memcpy(mine + 12 + 8 * w, &hd, 64); mine[9 + w] = 1; }
                    else cudaGetLastError();                 // the peers see "no handle" and fail in their cluster call
                }
        std::vector<int64_t> all((size_t)NW * R);
        int rc = nccl_allgather_i64(ctx, mine, NW, all.data()); if (rc) return rc;
        const int K = peer ? 2 : 4;        // arrays per rank in the payload gather: records of both lists (+ the INS blobs)
        std::vector<int64_t> bytes((size_t)K * R);
        int64_t nt[2] = {0, 0}, it[2] = {0, 0}, my_blob_base[2] = {0, 0};
        for (int r = 0; r < R; ++r)
            for (int w = 0; w < 2; ++w) {
                if (peer) bytes[2 * r + w] = all[(size_t)NW * r + 2 * w] * (int64_t)sizeof(svim_sig);
                else { bytes[4 * r + 2 * w] = all[(size_t)NW * r + 2 * w] * (int64_t)sizeof(svim_sig); bytes[4 * r + 2 * w + 1] = all[(size_t)NW * r + 2 * w + 1]; }
                if (r < ctx->rank) my_blob_base[w] += all[(size_t)NW * r + 2 * w + 1];
                nt[w] += all[(size_t)NW * r + 2 * w]; it[w] += all[(size_t)NW * r + 2 * w + 1];
            }
        if (peer) {
            ctx->peer_failed = false;
            for (int w = 0; w < 2; ++w) {
                SigSet& set = ctx->sets[w];
                set.seg_base.assign((size_t)R + 1, 0); set.seg_ptr.assign((size_t)R, nullptr);
                ctx->peer_map[w].resize((size_t)R);
                for (int r = 0; r < R; ++r) {
                    const int64_t len = all[(size_t)NW * r + 2 * w + 1];
                    set.seg_base[r + 1] = set.seg_base[r] + len;
                    if (len <= 0) continue;
                    if (r == ctx->rank) { set.seg_ptr[r] = set.ins.as<uint8_t>(); continue; }
                    PeerMap& m = ctx->peer_map[w][r];
                    const void* hd = &all[(size_t)NW * r + 12 + 8 * w];
                    if (!all[(size_t)NW * r + 9 + w]) { ctx->peer_failed = true; continue; }
                    if (!m.p || memcmp(&m.handle, hd, 64) != 0) {          // first time, or the peer's buffer was reallocated
                        if (m.p) { cudaIpcCloseMemHandle(m.p); m.p = nullptr; }
                        memcpy(&m.handle, hd, 64);
                        if (cudaIpcOpenMemHandle(&m.p, m.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); m.p = nullptr; ctx->peer_failed = true; continue; }
                    }
                    set.seg_ptr[r] = (const uint8_t*)m.p;
                }
            }
        }
        const void* src[4]; uint8_t* dst[4];
        for (int w = 0; w < 2; ++w) {
            SigSet& set = ctx->sets[w];
            // make local records global before sending: record index += aln_base, INS offset += blob base
            if (set.n) { ctx->launches++; k_rebase_sigs<<<(uint32_t)((set.n + 255) / 256), 256, 0, ctx->stream>>>(set.recs.as<svim_sig>(), (uint32_t)set.n, aln_base, (uint64_t)my_blob_base[w]); }
            SVIM_CUDA(ctx->d_xchg[2 * w].ensure((size_t)(nt[w] + 1) * sizeof(svim_sig)));
            if (peer) { src[w] = set.recs.p; dst[w] = ctx->d_xchg[2 * w].as<uint8_t>(); }
            else {
                SVIM_CUDA(ctx->d_xchg[2 * w + 1].ensure((size_t)it[w] + 16));
                src[2 * w] = set.recs.p; src[2 * w + 1] = set.ins.p;
                dst[2 * w] = ctx->d_xchg[2 * w].as<uint8_t>(); dst[2 * w + 1] = ctx->d_xchg[2 * w + 1].as<uint8_t>();
            }
        }
        rc = nccl_allgatherv_packed(ctx, K, src, bytes.data(), dst); if (rc) return rc;
        for (int w = 0; w < 2; ++w) {
            SigSet& set = ctx->sets[w];
            std::swap(set.recs, ctx->d_xchg[2 * w]);
            if (!peer) std::swap(set.ins, ctx->d_xchg[2 * w + 1]);
            set.n = nt[w]; set.ins_bytes = it[w]; set.segmented = peer;
        }
        svim_collect_stats& s = ctx->cstats;
        s.n_signatures = ctx->sets[0].n; s.ins_bytes = ctx->sets[0].ins_bytes; s.n_twin_signatures = ctx->sets[1].n; s.twin_ins_bytes = ctx->sets[1].ins_bytes;
        s.n_sa_bad_fields = s.n_no_read_length = s.n_primaries = s.n_data_errors = 0;
        for (int r = 0; r < R; ++r) { s.n_sa_bad_fields += all[(size_t)NW * r + 5]; s.n_no_read_length += all[(size_t)NW * r + 6]; s.n_primaries += all[(size_t)NW * r + 7]; s.n_data_errors += all[(size_t)NW * r + 8]; }
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
    SVIM_CUDA(ctx->d_xchg[6].ensure(16));
    SVIM_CUDA(cudaMemcpyAsync(ctx->d_xchg[6].p, value, 8, cudaMemcpyHostToDevice, ctx->stream));
    SVIM_NCCL(ncclAllReduce(ctx->d_xchg[6].p, ctx->d_xchg[6].p, 1, ncclDouble, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    SVIM_CUDA(cudaMemcpyAsync(value, ctx->d_xchg[6].p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SVIM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"

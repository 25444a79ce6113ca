// Context, device buffers, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>     // header-only; ranges cost nothing unless a tool (nsys / ncu --nvtx) is attached
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "collect.cuh"

#define SVIM_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ctx->set_error(SVIMGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SVIMGPU_ERR_CUDA;                                                        \
        }                                                                                   \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct PinBuf {     // page-locked host buffer that only grows (D2H / H2D at full PCIe rate, async on a stream)
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct DevSoa {
    int64_t n;
    const int32_t* tid; const int32_t* pos; const uint16_t* flag; const uint8_t* mapq;
    const uint32_t* n_cigar; const uint64_t* cigar_off; const int32_t* l_seq; const uint64_t* seq_off;
    const uint64_t* sa_off; const uint32_t* sa_len; const uint32_t* qname_id;
    const uint32_t* cigar; const uint8_t* seq; const uint8_t* sa;
};

struct SigQueue {
    svim_sig* recs;
    uint32_t* count;
    uint32_t cap;
};

// per-record segment summary (query-sorted mode)
struct SegSum { int64_t ref_end, q_start, q_end, read_len; };

// written by the CIGAR scan for every primary that has an SA tag and no hard clip
struct ChainWork {
    uint32_t aln_idx, ord_sig, ord_twin, pad;
    int64_t ref_end, q_start, q_end, read_len;
};

enum {  // device counters
    CNT_MAIN = 0, CNT_TWIN, CNT_WORK, CNT_NEXT_ALN, CNT_PRIMARIES, CNT_BAD_FIELDS, CNT_NO_READLEN, CNT_DATA_ERR,
    CNT_TOO_MANY, CNT_OVERFLOW, CNT_MYERS_NEXT, CNT_HOLES_MAIN, CNT_HOLES_TWIN, CNT_N
};

enum {  // timing slots
    T_H2D = 0, T_SCAN, T_CHAIN, T_SORTBACK, T_GATHER, T_COLLECT_D2H, T_CSIG, T_KEYSORT, T_PARTITION, T_SAMPLE, T_PAIRS,
    T_MYERS, T_LINKAGE, T_CONSOLIDATE, T_ORDER, T_CLUSTER_D2H, T_EXCHANGE, T_GENO_PREP, T_GENO, T_CUTPASTE,
    T_BAM_INFLATE, T_BAM_BOUNDS, T_BAM_ROWS, T_BAM_FILL, T_BAM_NAMES, T_PEER_INS, T_N
};

struct SigSet {   // one signature list on the device (main / all_bnds twins)
    DevBuf recs;        // svim_sig[n], emission order
    DevBuf ins;         // INS blob
    int64_t n = 0, ins_bytes = 0;
    // After svimgpu_exchange_signatures (peer mode): recs holds every rank's records, ins_bytes is the size of the GLOBAL blob the
    // records' seq_off point into, but the bytes stay where they were produced: rank r's piece [seg_base[r], seg_base[r+1]) is read
    // through seg_ptr[r] (this rank's own `ins`, or a CUDA-IPC mapping of the peer's buffer: NVLink loads).
    bool segmented = false;
    std::vector<int64_t> seg_base;
    std::vector<const uint8_t*> seg_ptr;
};

struct PeerMap { cudaIpcMemHandle_t handle; void* p = nullptr; };   // an opened mapping of a peer rank's INS blob

#define SVIM_AUX_STREAMS 20     // every edit-distance bucket on its own stream: the block scheduler packs their CTAs side by side

struct svimgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream[SVIM_AUX_STREAMS];      // concurrent Myers bins
    cudaEvent_t aux_ev[SVIM_AUX_STREAMS + 1];
    svim_params params;
    std::string err;
    int err_code = 0;

    // contigs / genome
    int32_t n_contigs = 0;
    std::vector<int32_t> h_rank, h_rank_to_tid;
    DevBuf d_names, d_name_off, d_rank, d_rank_to_tid;
    DevBuf d_genome, d_genome_off, d_genome_codes, d_ins_codes;     // *_codes: k_tpp_encode images for the thread-per-pair kernels
    int64_t genome_bytes = 0;
    int32_t genome_contigs = 0;

    // alignments
    DevBuf d_soa[14];
    DevBuf d_cig16, d_cig16_off, d_cig16_err;
    DevBuf d_bam_names, d_bam_name_off, d_bam_rec_of_id;   // svimgpu_decode_bam: read names (NUL-terminated), their offsets, first record of every name id
    int64_t bam_names_bytes = 0; uint32_t bam_n_names = 0;      // 16-bit packed CIGAR stream as uploaded (svim_aln_soa.cigar16), expanded into d_soa[11]
    DevSoa soa; bool have_soa = false;
    int64_t cigar_words = 0, seq_bytes = 0, sa_bytes = 0;
    // lazy SEQ (collect_host): SEQ stays on the host, only the bytes of emitted insertions are staged and uploaded
    bool lazy_seq = false; const uint8_t* h_seq = nullptr; const uint64_t* h_seq_off = nullptr; uint32_t lazy_aln_base = 0;
    uint8_t* h_stage = nullptr; size_t h_stage_cap = 0;
    std::vector<svim_sig> h_lazy_recs; std::vector<uint64_t> h_lazy_off;
    DevBuf d_stage, d_stage_off;

    // collect state
    DevBuf d_big_list, d_big_caps, d_big_scratch;    // large-read pass of the segment chain (reads above SVIM_MAX_SEGMENTS)
    DevBuf d_counters, d_queue[2], d_work, d_sort_tmp, d_keys[2], d_vals[2], d_scan;
    SigSet sets[2];
    bool collected = false;
    bool qs_mode = false;      // query-sorted COLLECT (SVIM_COLLECT.py:96-129)
    DevBuf d_qs_info, d_qs_grp, d_qs_segsum, d_qs_mem_off, d_qs_mem_idx;
    int myers_mode = 1;        // k_myers_fast formulation (env SVIM_MYERS_MODE): 0 ALU pipe, 1 FMA pipe, 2 FMA pipe + IMAD.HI
    int myers_trace = 0;       // env SVIM_MYERS_TRACE=1: per-launch start/end times of the edit-distance kernels on stderr
    DevBuf d_myers_trace;
    int myers_tpp = 1;         // thread-per-pair banded kernels for pairs whose window fits 28 blocks (env SVIM_MYERS_TPP=0: wavefront kernels only)
    int myers_band_num = 156, myers_band_add = 20;   // banded first pass with k = m*num/1024 + add (env SVIM_MYERS_BAND=num,add; 0 = off)
    int scan_chunks = 1;       // per-warp chunked queue-slot reservation (env SVIM_SCAN_CHUNKS=0: one atomic per signature)
    int scan_variant = 0;      // 0: 128-bit LDG streaming, 1: cp.async.bulk ring (env SVIM_SCAN_VARIANT)
    svim_collect_stats cstats;

    // host copy of the collected lists (collect_host): pinned, filled on a copy stream while CLUSTER runs
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_collect_done = nullptr, ev_host_copy[2] = {nullptr, nullptr};
    uint8_t* h_out[2] = {nullptr, nullptr}; size_t h_out_cap[2] = {0, 0};
    bool host_copy[2] = {false, false};      // sets[i] is (being) mirrored in h_out[i]: records, then the INS blob at ins_off

    // cluster state
    DevBuf d_csig, d_csig_sorted, d_cins;   // input signatures (emission order / key order), INS blob when uploaded
    const uint8_t* cluster_ins = nullptr; int64_t cluster_ins_bytes = 0;
    int64_t n_csig = 0; bool have_csig = false;
    DevBuf d_order, d_head, d_partid, d_part_off, d_samp_off, d_samp_idx, d_labels, d_part_ncl, d_part_nkept, d_part_stats;
    DevBuf d_plist, d_myers_scratch[48], d_myers_ctl;   // scratch: [0,10) unbanded bins, [10] 8-plane kernel, [12,22) banded shapes, [24,42) thread-per-pair buckets
    int64_t cluster_max_ins_len = 0;
    const int32_t* cluster_rank_to_tid = nullptr; int32_t cluster_n_ranks = 0;
    DevBuf d_user_rank_to_tid;
    uint32_t shard_lo = 0, shard_hi = 0;
    DevBuf d_cl_off, d_mem_off, d_clusters, d_clusters_sorted, d_members, d_pair_off, d_pair_ed, d_pairs, d_ckeys[2], d_cvals[2];
    // partition plan on the device (cluster.cu): per-partition scan, shard cuts, work lists; the host only sees a small header, the
    // sizes of the partitions above 100 (it owns the sampling stream) and the final records
    DevBuf d_pmeta, d_ppref, d_ptype, d_hdr, d_large_list, d_picks;
    PinBuf h_hdr, h_picks, h_clusters, h_members;
    uint32_t n_clusters_host = 0, n_members_host = 0;
    int64_t n_partitions = 0;
    svim_cluster_stats clstats;
    bool clustered = false;

    // genotype (SVIM_genotyping.py:34-93): works on the record rows + CIGAR left in HBM by the last upload
    bool rows_resident = false, geno_ready = false; int32_t geno_contigs = 0;
    DevBuf d_geno_end, d_geno_max, d_geno_rows, d_geno_cand, d_geno_out, d_geno_var, d_geno_clen;

    // multi-GPU
    void* nccl_comm = nullptr; int nranks = 1, rank = 0;
    bool expand8_staged = true;            // k_expand_cigar8_staged (shared-memory staged write-out); SVIM_EXPAND8_STAGED=0: k_expand_cigar8
    bool peer_ins = true;                  // SVIM_PEER_INS=0: all-gather the INS blobs too (every rank holds a full copy)
    bool mirror_gathered_ins = true;       // svimgpu_mirror_gathered_ins
    bool host_copy_has_ins[2] = {true, true};
    bool host_copy_pending = false;        // collect_host on a rank of a multi-GPU job: the mirror starts after the exchange
    bool peer_failed = false;              // a mapping could not be opened: the next sharded cluster reports it through its status agreement
    std::vector<PeerMap> peer_map[2];
    int cluster_seg = -1;                  // >= 0: the clustering input is the segmented set `cluster_seg` (cluster_ins is filled per call)
    DevBuf d_seg_tab, d_shard_ins, d_shard_len;
    DevBuf d_xchg[8];          // [0,4) gathered arrays, [4] packed send/receive slots, [5] count gather, [6] barrier word

    // timing
    cudaEvent_t ev[2 * T_N];
    double ms[T_N];
    bool ev_rec[T_N];
    int64_t launches = 0;
    cudaEvent_t user_ev[2];

    void set_error(int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
};

inline void svimgpu_ctx::set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    err = buf; err_code = code;
}

extern const char* const k_stage_names[];      // api.cu: one name per timing slot, also the NVTX range of the stage

struct StageTimer {      // CUDA-event timing of one stage on the context's stream + an NVTX range around its host side
    svimgpu_ctx* c; int slot;
    StageTimer(svimgpu_ctx* ctx, int s) : c(ctx), slot(s) { nvtxRangePushA(k_stage_names[s]); cudaEventRecord(c->ev[2 * s], c->stream); c->ev_rec[s] = true; }
    ~StageTimer() { cudaEventRecord(c->ev[2 * slot + 1], c->stream); nvtxRangePop(); }
};

inline void timings_begin(svimgpu_ctx* ctx) { for (int i = 0; i < T_N; ++i) { ctx->ev_rec[i] = false; ctx->ms[i] = 0.0; } }
inline void timings_end(svimgpu_ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < T_N; ++i)
        if (ctx->ev_rec[i]) { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]) == cudaSuccess) ctx->ms[i] = ms; }
}

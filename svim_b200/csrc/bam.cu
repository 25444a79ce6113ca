// BAM file -> resident record buffer, entirely on the GPU (the step in front of COLLECT,
// `bam.fetch(until_eof=True)` at SVIM_COLLECT.py:133).
//
// The host only indexes the BGZF block headers (csrc_host/bamio.cpp, ~0.1 s for 4.7 GB) and hands over the compressed file:
// it is also the smaller PCIe payload (4.7 GB against 6.3 GB of decoded CIGAR for BASELINE configs[1]).  On the device
//   k_inflate       one warp per BGZF block: lane 0 walks the Huffman stream, the warp copies the LZ77 matches (bam_kernels.cu)
//   k_starts/k_chain/k_verify   record boundaries without a sequential pass: every 64 KiB chunk of the inflated stream guesses its
//                   first record start and hops block_size fields into the next chunk; all guesses are right iff every hop chain
//                   lands on the next chunk's guess (induction from the header end), else the call declines (SVIMGPU_ERR_DATA)
//   k_rows, scans, k_fill       rows, blob offsets, CIGAR words (padded to 16 bytes) / packed SEQ / SA text / read names
//   k_name_hash, radix sort, k_name_groups ...   read-name ids: equal names <=> equal ids, numbered by first appearance like the
//                   host decoder (64-bit hash, groups verified byte for byte: a hash collision declines the call)
// and the buffers of svimgpu_upload_alignments are filled in place: svimgpu_collect runs on them with no PCIe round trip.
#pragma once
#include "ctx.cuh"
#include "bam_kernels.cu"

__global__ void k_name_hash(const uint8_t* __restrict__ names, const uint64_t* __restrict__ name_off, int64_t n, uint64_t* __restrict__ hash, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = names + name_off[i];
    uint64_t h = 0xcbf29ce484222325ull;                       // FNV-1a, then a finaliser (names of one run differ in a few trailing digits)
    for (int k = 0; k < 256 && p[k]; ++k) { h ^= p[k]; h *= 0x100000001b3ull; }
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    hash[i] = h; idx[i] = (uint32_t)i;
}

// sorted by (hash, record index): head[j] = first record of its name group; groups are verified byte for byte against their head
__global__ void k_name_heads(const uint8_t* __restrict__ names, const uint64_t* __restrict__ name_off, const uint64_t* __restrict__ hash, const uint32_t* __restrict__ idx,
                             int64_t n, uint32_t* __restrict__ head_flag, uint32_t* __restrict__ first_flag, uint32_t* __restrict__ bad) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const bool head = j == 0 || hash[j] != hash[j - 1];
    head_flag[j] = head ? 1u : 0u;
    if (head) first_flag[idx[j]] = 1u;                        // stable sort: the head is the group's first record in file order
    else {
        const uint8_t* a = names + name_off[idx[j]]; const uint8_t* b = names + name_off[idx[j - 1]];
        for (int k = 0; k < 256; ++k) { if (a[k] != b[k]) { atomicExch(bad, 1u); break; } if (!a[k]) break; }
    }
}

// group g (rank among heads in sorted order) -> its first record; record -> id = number of names that appear earlier in the file
__global__ void k_name_group_first(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head_flag, const uint32_t* __restrict__ head_rank, int64_t n,
                                   uint32_t* __restrict__ group_first) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && head_flag[j]) group_first[head_rank[j]] = idx[j];
}
__global__ void k_name_ids(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head_flag, const uint32_t* __restrict__ head_rank,
                           const uint32_t* __restrict__ group_first, const uint32_t* __restrict__ first_rank, int64_t n, uint32_t* __restrict__ qname_id,
                           uint32_t* __restrict__ rec_of_id) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t g = head_rank[j] - (head_flag[j] ? 0u : 1u);          // exclusive rank of heads: a non-head belongs to the previous head
    const uint32_t first = group_first[g];
    const uint32_t id = first_rank[first];
    qname_id[idx[j]] = id;
    if (head_flag[j]) rec_of_id[id] = first;
}

template <class T>
static int bam_scan(svimgpu_ctx* ctx, const T* in, T* out, int64_t n_items) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n_items, ctx->stream);
    SVIM_CUDA(ctx->d_sort_tmp.ensure(need));
    SVIM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, need, in, out, (int)n_items, ctx->stream));
    return 0;
}

static int bam_decode_run(svimgpu_ctx* ctx, const uint8_t* file, int64_t file_bytes, const BgBlock* blocks, int64_t n_blocks, int64_t first_record,
                          int32_t n_ref, svim_bam_info* info) {
    cudaStream_t st = ctx->stream;
    const uint64_t usize = n_blocks ? blocks[n_blocks - 1].uoff + blocks[n_blocks - 1].ulen : 0;
    DevBuf d_file, d_data, d_blocks, d_flags, d_st, d_en, d_cnt, d_base, d_recoff, d_tmp64[4], d_src[3], d_hash[2], d_idx[2], d_flag2[4];
    auto release_all = [&]() {
        DevBuf* all[] = {&d_file, &d_data, &d_blocks, &d_flags, &d_st, &d_en, &d_cnt, &d_base, &d_recoff, &d_tmp64[0], &d_tmp64[1], &d_tmp64[2], &d_tmp64[3],
                         &d_src[0], &d_src[1], &d_src[2], &d_hash[0], &d_hash[1], &d_idx[0], &d_idx[1], &d_flag2[0], &d_flag2[1], &d_flag2[2], &d_flag2[3]};
        for (DevBuf* b : all) b->release();
    };
    struct Guard { decltype(release_all)& f; ~Guard() { f(); } } guard{release_all};
    SVIM_CUDA(d_file.ensure((size_t)file_bytes + 64)); SVIM_CUDA(d_data.ensure((size_t)usize + 256)); SVIM_CUDA(d_blocks.ensure((size_t)(n_blocks + 1) * sizeof(BgBlock)));
    SVIM_CUDA(d_flags.ensure(128));
    SVIM_CUDA(cudaMemsetAsync(d_flags.p, 0, 128, st));
    uint32_t* flags_d = d_flags.as<uint32_t>();
    SVIM_CUDA(cudaMemcpyAsync(d_blocks.p, blocks, (size_t)n_blocks * sizeof(BgBlock), cudaMemcpyHostToDevice, st));
    // ---- compressed file H2D in slices, each slice's blocks inflated as soon as it has landed (copy stream / compute stream) ----------
    {
        StageTimer t(ctx, T_BAM_INFLATE);
        const int64_t SLICE = (int64_t)256 << 20;
        // expanding near matches on the decoding lane (SVIM_BAM_INLINE=1) loses on BAM streams: 82 % of their matches reach further back than the
        // 8-byte register tail (measured: 33.7 M of 41.3 M on BASELINE configs[1]), and each of those then costs a queue drain
        const int inline_mode = getenv("SVIM_BAM_INLINE") ? atoi(getenv("SVIM_BAM_INLINE")) : 0;
        int64_t b0 = 0;
        cudaEvent_t ev_done[2]; cudaEventCreateWithFlags(&ev_done[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_done[1], cudaEventDisableTiming);
        SVIM_CUDA(cudaEventRecord(ctx->aux_ev[SVIM_AUX_STREAMS], st));
        SVIM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->aux_ev[SVIM_AUX_STREAMS], 0));        // the block table and flags are on their way
        int slice = 0;
        while (b0 < n_blocks) {
            int64_t b1 = b0;
            const uint64_t lo = blocks[b0].coff;
            while (b1 < n_blocks && (int64_t)(blocks[b1].coff + blocks[b1].clen - lo) <= SLICE) ++b1;
            if (b1 == b0) ++b1;
            const uint64_t hi = blocks[b1 - 1].coff + blocks[b1 - 1].clen;
            SVIM_CUDA(cudaMemcpyAsync(d_file.as<uint8_t>() + lo, file + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, ctx->copy_stream));
            SVIM_CUDA(cudaEventRecord(ev_done[slice & 1], ctx->copy_stream));
            SVIM_CUDA(cudaStreamWaitEvent(st, ev_done[slice & 1], 0));
            ctx->launches++;
            k_inflate<<<(unsigned)((b1 - b0 + BG_WARPS - 1) / BG_WARPS), 32 * BG_WARPS, 0, st>>>(d_file.as<uint8_t>(), d_blocks.as<BgBlock>() + b0, b1 - b0, d_data.as<uint8_t>(), flags_d, inline_mode);
            b0 = b1; ++slice;
        }
        SVIM_CUDA(cudaGetLastError());
        cudaStreamSynchronize(ctx->copy_stream);
        cudaEventDestroy(ev_done[0]); cudaEventDestroy(ev_done[1]);
    }
    uint32_t flags[16];
    if (getenv("SVIM_BAM_DEBUG")) {
        unsigned long long dbg[6]; cudaStreamSynchronize(st); cudaMemcpy(dbg, flags_d + 8, 48, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[bam inflate] literals %llu  inline matches %llu  queued %llu (far %llu, long %llu)  pauses %llu\n", dbg[0], dbg[1], dbg[2], dbg[4], dbg[5], dbg[3]);
    }
    // ---- record boundaries --------------------------------------------------------------------------------------------------------
    const uint64_t first = (uint64_t)first_record;
    const int64_t n_chunks = usize > first ? (int64_t)((usize - first + BG_CHUNK - 1) / BG_CHUNK) : 0;
    int64_t n = 0;
    {
        StageTimer t(ctx, T_BAM_BOUNDS);
        SVIM_CUDA(d_st.ensure((size_t)(n_chunks + 2) * 8)); SVIM_CUDA(d_en.ensure((size_t)(n_chunks + 2) * 8)); SVIM_CUDA(d_cnt.ensure((size_t)(n_chunks + 2) * 4));
        SVIM_CUDA(d_base.ensure((size_t)(n_chunks + 2) * 8));
        if (n_chunks) {
            const unsigned gb = (unsigned)((n_chunks + 1 + 127) / 128);
            ctx->launches += 3;
            k_starts<<<gb, 128, 0, st>>>(d_data.as<uint8_t>(), usize, first, n_chunks, n_ref, d_st.as<uint64_t>());
            k_chain<<<gb, 128, 0, st>>>(d_data.as<uint8_t>(), usize, first, n_chunks, d_st.as<uint64_t>(), d_cnt.as<uint32_t>(), d_en.as<uint64_t>(), nullptr, nullptr);
            k_verify<<<gb, 128, 0, st>>>(d_st.as<uint64_t>(), d_en.as<uint64_t>(), n_chunks, flags_d + 1);
            SVIM_CUDA(cudaMemsetAsync(d_cnt.as<uint32_t>() + n_chunks, 0, 4, st));
            size_t need = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, need, d_cnt.as<uint32_t>(), d_base.as<uint64_t>(), (int)(n_chunks + 1), st);
            SVIM_CUDA(ctx->d_sort_tmp.ensure(need));
            SVIM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, need, d_cnt.as<uint32_t>(), d_base.as<uint64_t>(), (int)(n_chunks + 1), st));
            uint64_t total = 0;
            SVIM_CUDA(cudaMemcpyAsync(&total, d_base.as<uint64_t>() + n_chunks, 8, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaMemcpyAsync(flags, flags_d, 64, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            if (flags[0]) { ctx->set_error(SVIMGPU_ERR_DATA, "bam: malformed DEFLATE data in a BGZF block (code %u)", flags[0]); return SVIMGPU_ERR_DATA; }
            if (flags[1] & 2u) { ctx->set_error(SVIMGPU_ERR_DATA, "bam: truncated record stream"); return SVIMGPU_ERR_DATA; }
            if (flags[1] & 1u) { ctx->set_error(SVIMGPU_ERR_DATA, "bam: a speculative record start was wrong (decode this file on the host)"); return SVIMGPU_ERR_DATA; }
            n = (int64_t)total;
        } else {
            SVIM_CUDA(cudaMemcpyAsync(flags, flags_d, 64, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            if (flags[0]) { ctx->set_error(SVIMGPU_ERR_DATA, "bam: malformed DEFLATE data in a BGZF block (code %u)", flags[0]); return SVIMGPU_ERR_DATA; }
        }
        if (n >= ((int64_t)1 << 32) - 1) { ctx->set_error(SVIMGPU_ERR_LIMIT, "bam: more than 2^32 records"); return SVIMGPU_ERR_LIMIT; }
        SVIM_CUDA(d_recoff.ensure((size_t)(n + 1) * 8));
        if (n) { ctx->launches++; k_chain<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(d_data.as<uint8_t>(), usize, first, n_chunks, d_st.as<uint64_t>(), d_cnt.as<uint32_t>(),
                                                                                       d_en.as<uint64_t>(), d_recoff.as<uint64_t>(), d_base.as<uint64_t>()); }
    }
    // ---- rows + blob offsets: straight into the buffers of svimgpu_upload_alignments -------------------------------------------------
    const size_t row_bytes[11] = {4, 4, 2, 1, 4, 8, 4, 8, 8, 4, 4};
    for (int i = 0; i < 11; ++i) SVIM_CUDA(ctx->d_soa[i].ensure((size_t)(n + 1) * row_bytes[i] + 64));
    for (int k = 0; k < 4; ++k) SVIM_CUDA(d_tmp64[k].ensure((size_t)(n + 1) * 8));
    for (int k = 0; k < 3; ++k) SVIM_CUDA(d_src[k].ensure((size_t)(n + 1) * 4));
    BgRows r;
    r.tid = ctx->d_soa[0].as<int32_t>(); r.pos = ctx->d_soa[1].as<int32_t>(); r.flag = ctx->d_soa[2].as<uint16_t>(); r.mapq = ctx->d_soa[3].as<uint8_t>();
    r.n_cigar = ctx->d_soa[4].as<uint32_t>(); r.l_seq = ctx->d_soa[6].as<int32_t>(); r.sa_len = ctx->d_soa[9].as<uint32_t>();
    r.cig_words = ctx->d_soa[5].as<uint64_t>(); r.seq_bytes = ctx->d_soa[7].as<uint64_t>(); r.sa_bytes = ctx->d_soa[8].as<uint64_t>(); r.name_bytes = d_tmp64[0].as<uint64_t>();
    r.sa_src = d_src[0].as<uint32_t>(); r.cig_src = d_src[1].as<uint32_t>(); r.n_core = d_src[2].as<uint32_t>();
    uint64_t tot[4] = {0, 0, 0, 0};
    {
        StageTimer t(ctx, T_BAM_ROWS);
        if (n) {
            ctx->launches++;
            k_rows<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_data.as<uint8_t>(), d_recoff.as<uint64_t>(), n, r, flags_d + 2);
            uint64_t* sized[4] = {r.cig_words, r.seq_bytes, r.sa_bytes, r.name_bytes};
            for (int k = 0; k < 4; ++k) {
                SVIM_CUDA(cudaMemsetAsync(sized[k] + n, 0, 8, st));
                int rc = bam_scan(ctx, sized[k], sized[k], n + 1); if (rc) return rc;
                SVIM_CUDA(cudaMemcpyAsync(&tot[k], sized[k] + n, 8, cudaMemcpyDeviceToHost, st));
            }
            SVIM_CUDA(cudaMemcpyAsync(flags, flags_d, 64, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            if (flags[2]) { ctx->set_error(SVIMGPU_ERR_DATA, "bam: corrupt alignment record"); return SVIMGPU_ERR_DATA; }
        }
    }
    SVIM_CUDA(ctx->d_soa[11].ensure((size_t)tot[0] * 4 + 64)); SVIM_CUDA(ctx->d_soa[12].ensure((size_t)tot[1] + 64)); SVIM_CUDA(ctx->d_soa[13].ensure((size_t)tot[2] + 64));
    SVIM_CUDA(ctx->d_bam_names.ensure((size_t)tot[3] + 64)); SVIM_CUDA(ctx->d_bam_name_off.ensure((size_t)(n + 1) * 8)); SVIM_CUDA(ctx->d_bam_rec_of_id.ensure((size_t)(n + 1) * 4));
    {
        StageTimer t(ctx, T_BAM_FILL);
        if (n) {
            ctx->launches++;
            k_fill<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, st>>>(d_data.as<uint8_t>(), d_recoff.as<uint64_t>(), n, r, ctx->d_soa[11].as<uint32_t>(), ctx->d_soa[12].as<uint8_t>(),
                                                                              ctx->d_soa[13].as<uint8_t>(), ctx->d_bam_names.as<uint8_t>());
            SVIM_CUDA(cudaMemcpyAsync(ctx->d_bam_name_off.p, r.name_bytes, (size_t)(n + 1) * 8, cudaMemcpyDeviceToDevice, st));
        }
    }
    // ---- read-name ids ------------------------------------------------------------------------------------------------------------------
    uint32_t n_names = 0;
    {
        StageTimer t(ctx, T_BAM_NAMES);
        if (n) {
            const unsigned gb = (unsigned)((n + 255) / 256);
            for (int k = 0; k < 2; ++k) { SVIM_CUDA(d_hash[k].ensure((size_t)n * 8)); SVIM_CUDA(d_idx[k].ensure((size_t)n * 4)); }
            for (int k = 0; k < 4; ++k) SVIM_CUDA(d_flag2[k].ensure((size_t)(n + 1) * 4));
            ctx->launches++;
            k_name_hash<<<gb, 256, 0, st>>>(ctx->d_bam_names.as<uint8_t>(), ctx->d_bam_name_off.as<uint64_t>(), n, d_hash[0].as<uint64_t>(), d_idx[0].as<uint32_t>());
            cub::DoubleBuffer<uint64_t> dk(d_hash[0].as<uint64_t>(), d_hash[1].as<uint64_t>());
            cub::DoubleBuffer<uint32_t> dv(d_idx[0].as<uint32_t>(), d_idx[1].as<uint32_t>());
            size_t need = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, 64, st);
            SVIM_CUDA(ctx->d_sort_tmp.ensure(need));
            SVIM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->d_sort_tmp.p, need, dk, dv, (int)n, 0, 64, st));
            uint32_t* head_flag = d_flag2[0].as<uint32_t>(); uint32_t* head_rank = d_flag2[1].as<uint32_t>();
            uint32_t* first_flag = d_flag2[2].as<uint32_t>(); uint32_t* first_rank = d_flag2[3].as<uint32_t>();
            SVIM_CUDA(cudaMemsetAsync(first_flag, 0, (size_t)(n + 1) * 4, st)); SVIM_CUDA(cudaMemsetAsync(head_flag + n, 0, 4, st));
            ctx->launches += 3;
            k_name_heads<<<gb, 256, 0, st>>>(ctx->d_bam_names.as<uint8_t>(), ctx->d_bam_name_off.as<uint64_t>(), dk.Current(), dv.Current(), n, head_flag, first_flag, flags_d + 3);
            int rc = bam_scan(ctx, head_flag, head_rank, n + 1); if (rc) return rc;
            rc = bam_scan(ctx, first_flag, first_rank, n + 1); if (rc) return rc;
            uint32_t* group_first = d_idx[dv.selector ^ 1].as<uint32_t>();          // the sort's spare buffer
            k_name_group_first<<<gb, 256, 0, st>>>(dv.Current(), head_flag, head_rank, n, group_first);
            k_name_ids<<<gb, 256, 0, st>>>(dv.Current(), head_flag, head_rank, group_first, first_rank, n, ctx->d_soa[10].as<uint32_t>(), ctx->d_bam_rec_of_id.as<uint32_t>());
            SVIM_CUDA(cudaMemcpyAsync(&n_names, first_rank + n, 4, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaMemcpyAsync(flags, flags_d, 64, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            if (flags[3]) { ctx->set_error(SVIMGPU_ERR_DATA, "bam: two read names share a 64-bit hash (decode this file on the host)"); return SVIMGPU_ERR_DATA; }
        }
    }
    SVIM_CUDA(cudaGetLastError());
    DevSoa& d = ctx->soa;
    d.n = n;
    d.tid = ctx->d_soa[0].as<int32_t>(); d.pos = ctx->d_soa[1].as<int32_t>(); d.flag = ctx->d_soa[2].as<uint16_t>(); d.mapq = ctx->d_soa[3].as<uint8_t>();
    d.n_cigar = ctx->d_soa[4].as<uint32_t>(); d.cigar_off = ctx->d_soa[5].as<uint64_t>(); d.l_seq = ctx->d_soa[6].as<int32_t>();
    d.seq_off = ctx->d_soa[7].as<uint64_t>(); d.sa_off = ctx->d_soa[8].as<uint64_t>(); d.sa_len = ctx->d_soa[9].as<uint32_t>();
    d.qname_id = ctx->d_soa[10].as<uint32_t>(); d.cigar = ctx->d_soa[11].as<uint32_t>(); d.seq = ctx->d_soa[12].as<uint8_t>(); d.sa = ctx->d_soa[13].as<uint8_t>();
    ctx->cigar_words = (int64_t)tot[0]; ctx->seq_bytes = (int64_t)tot[1]; ctx->sa_bytes = (int64_t)tot[2];
    ctx->bam_names_bytes = (int64_t)tot[3]; ctx->bam_n_names = n_names;
    ctx->have_soa = true; ctx->collected = false; ctx->rows_resident = true; ctx->geno_ready = false;
    ctx->lazy_seq = false; ctx->h_seq = nullptr; ctx->h_seq_off = nullptr; ctx->lazy_aln_base = 0;
    if (info) { info->n_records = n; info->cigar_words = (int64_t)tot[0]; info->seq_bytes = (int64_t)tot[1]; info->sa_bytes = (int64_t)tot[2];
                info->names_bytes = (int64_t)tot[3]; info->n_names = n_names; info->inflated_bytes = (int64_t)usize; }
    return 0;
}

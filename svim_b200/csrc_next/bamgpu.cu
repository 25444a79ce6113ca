// On-GPU BAM decoder, first version — EXPERIMENTAL, opt-in, not on the default path:
// built into its own library (libsvimbamgpu.so) so that libsvimgpu.so is untouched; `svim_b200.io.read_bam_gpu` is the only
// caller and tests/test_gpu_next.py (marker gpu_next, not part of -m gpu) compares it with the host decoder.
//
//   host:  BGZF block table (csrc_host/bamio.cpp) + BAM header -> compressed file bytes H2D
//   k_inflate      one warp per BGZF block, lane 0 runs bgzf_inflate_block with the Huffman tables in shared memory
//   k_starts       one thread per 64 KiB chunk of the inflated stream: speculative first record start (bam_find_record_start)
//   k_chain        one thread per chunk: hop block_size fields to the next chunk's territory; k_verify: every chain must land
//                  on the next chunk's guess (then all guesses are right by induction from the header end)
//   scan + k_offsets   record start offsets in stream order
//   k_rows         one thread per record: fixed fields, SA tag lookup, blob sizes;  exclusive scans -> blob offsets
//   k_fill         one warp per record: CIGAR words (padded to 4), packed SEQ, SA text, read name -> blobs
// Everything heavy is a thin wrapper over the SVIM_HD functions of bgzf_core.cuh, which the CPU tests replay.
// CG:B,I tags of records with more than 65535 CIGAR operations are put back like htslib's bam_tag2cigar (k_rows / k_fill).
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdio.h>
#include <string>
#include <vector>
#include "bgzf_core.cuh"

#define BG_WARPS 8
#define BG_CHUNK (1ull << 16)

struct BgBlock { uint64_t coff; uint64_t uoff; uint32_t clen, ulen; };

__global__ void __launch_bounds__(32 * BG_WARPS) k_inflate(const uint8_t* __restrict__ file, const BgBlock* __restrict__ blocks, int64_t n_blocks,
                                                            uint8_t* __restrict__ out, uint32_t* __restrict__ status) {
    __shared__ uint32_t tab[BG_WARPS][BGZF_TABLE_WORDS];
    const int64_t b = (int64_t)blockIdx.x * BG_WARPS + (threadIdx.x >> 5);
    if (b >= n_blocks || (threadIdx.x & 31) != 0) return;
    const BgBlock bl = blocks[b];
    const int rc = bgzf_inflate_block(file + bl.coff, bl.clen, out + bl.uoff, bl.ulen, tab[threadIdx.x >> 5]);
    if (rc) atomicMax(status, (uint32_t)rc);
}

__global__ void k_starts(const uint8_t* __restrict__ data, uint64_t size, uint64_t first, int64_t n_chunks, int32_t n_ref, uint64_t* __restrict__ st) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    if (c == n_chunks) { st[c] = size; return; }
    st[c] = c == 0 ? first : bam_find_record_start(data, size, first + (uint64_t)c * BG_CHUNK, n_ref, 3);
}

__global__ void k_chain(const uint8_t* __restrict__ data, uint64_t size, uint64_t first, int64_t n_chunks, const uint64_t* __restrict__ st,
                        uint32_t* __restrict__ cnt, uint64_t* __restrict__ en, uint64_t* __restrict__ rec_off, const uint64_t* __restrict__ base) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint64_t limit = c + 1 < n_chunks ? first + (uint64_t)(c + 1) * BG_CHUNK : size;
    uint64_t o = st[c]; uint32_t n = 0; bool cut = false;
    while (o < limit && o < size) {
        if (o + 4 > size) { cut = true; break; }
        uint32_t bs; memcpy(&bs, data + o, 4);
        if (bs < 32 || o + 4ull + bs > size) { cut = true; break; }
        if (rec_off) rec_off[base[c] + n] = o;         // second pass: record starts in stream order
        o += 4ull + bs; ++n;
    }
    if (!rec_off) { cnt[c] = n; en[c] = cut ? ~0ull : o; }
}

__global__ void k_verify(const uint64_t* __restrict__ st, const uint64_t* __restrict__ en, int64_t n_chunks, uint32_t* __restrict__ bad) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    if (en[c] == ~0ull) atomicOr(bad, 2u);              // truncated stream
    else if (en[c] != st[c + 1]) atomicOr(bad, 1u);     // a speculative start was wrong
}

struct BgRows {
    int32_t* tid; int32_t* pos; uint16_t* flag; uint8_t* mapq; uint32_t* n_cigar; int32_t* l_seq; uint32_t* sa_len;
    uint64_t* cig_words; uint64_t* seq_bytes; uint64_t* sa_bytes; uint64_t* name_bytes;      // per-record sizes -> scanned in place into offsets
    uint32_t* sa_src;                                                                       // SA payload offset inside the record
    uint32_t* cig_src; uint32_t* n_core;                                                    // CIGAR words (core or CG:B,I payload) / ops in the core
};

// aux walk (same as csrc_host/bamio.cpp::scan_aux): SA:Z payload and CG:B,I payload, offsets relative to the record start
struct BgAux { uint32_t sa_off, sa_len, cg_off, cg_n; };
__device__ void bg_scan_aux(const uint8_t* rec, uint64_t aux_begin, uint64_t rec_len, BgAux& hit) {
    uint64_t o = aux_begin;
    while (o + 3 <= rec_len) {
        const uint8_t t0 = rec[o], t1 = rec[o + 1], ty = rec[o + 2];
        o += 3;
        uint64_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': {
                uint64_t e = o;
                while (e < rec_len && rec[e]) ++e;
                if (t0 == 'S' && t1 == 'A' && ty == 'Z' && hit.sa_off == 0) { hit.sa_off = (uint32_t)o; hit.sa_len = (uint32_t)(e - o); }
                o = e + 1;
                continue;
            }
            case 'B': {
                if (o + 5 > rec_len) return;
                const uint8_t sub = rec[o]; uint32_t cnt; memcpy(&cnt, rec + o + 1, 4);
                const uint64_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                if (t0 == 'C' && t1 == 'G' && sub == 'I' && hit.cg_off == 0 && o + 5 + (uint64_t)cnt * 4 <= rec_len) { hit.cg_off = (uint32_t)(o + 5); hit.cg_n = cnt; }
                o += 5 + (uint64_t)cnt * es;
                continue;
            }
            default: return;
        }
        o += sz;
    }
}

__global__ void k_rows(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, int64_t n, BgRows r, uint32_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = data + rec_off[i];
    uint32_t bs; memcpy(&bs, p, 4);
    const uint8_t* q = p + 4;
    int32_t tid, pos, l_seq; uint16_t n_cig, flag;
    memcpy(&tid, q, 4); memcpy(&pos, q + 4, 4); memcpy(&n_cig, q + 12, 2); memcpy(&flag, q + 14, 2); memcpy(&l_seq, q + 16, 4);
    const uint32_t l_rn = q[8];
    const uint64_t aux = 32ull + l_rn + 4ull * n_cig + ((uint64_t)(l_seq < 0 ? 0 : l_seq) + 1) / 2 + (uint64_t)(l_seq < 0 ? 0 : l_seq);
    if (l_seq < 0 || aux > bs) { atomicOr(bad, 4u); l_seq = 0; }
    BgAux hit = {0, 0, 0, 0};
    if (aux < bs) bg_scan_aux(q, aux, bs, hit);
    const uint32_t so = hit.sa_off, sl = hit.sa_len;
    uint32_t n_ops = n_cig, cig_src = 32u + l_rn;
    if (hit.cg_n && n_cig > 0 && tid >= 0 && pos >= 0) {          // htslib sam.c bam_tag2cigar
        uint32_t c0; memcpy(&c0, q + 32 + l_rn, 4);
        if ((c0 & 15u) == 4u && (int64_t)(c0 >> 4) == (int64_t)l_seq) { n_ops = hit.cg_n; cig_src = hit.cg_off; }
    }
    r.tid[i] = tid; r.pos[i] = pos; r.flag[i] = flag; r.mapq[i] = q[9]; r.n_cigar[i] = n_ops; r.l_seq[i] = l_seq; r.sa_len[i] = sl; r.sa_src[i] = so;
    r.cig_src[i] = cig_src; r.n_core[i] = n_cig;
    r.cig_words[i] = (n_ops + 3u) & ~3u; r.seq_bytes[i] = ((uint64_t)l_seq + 1) / 2; r.sa_bytes[i] = sl; r.name_bytes[i] = l_rn ? l_rn : 1u;   // names keep their NUL
}

__global__ void __launch_bounds__(256) k_fill(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, int64_t n, BgRows r,
                                              uint32_t* __restrict__ cigar, uint8_t* __restrict__ seq, uint8_t* __restrict__ sa, uint8_t* __restrict__ names) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint8_t* q = data + rec_off[i] + 4;
    const uint32_t l_rn = q[8], n_cig = r.n_cigar[i];
    const int64_t l_seq = r.l_seq[i];
    const uint8_t* cg = q + r.cig_src[i];                 // records are not 4-byte aligned in the stream: byte-wise assembly
    uint32_t* cd = cigar + r.cig_words[i];
    const uint32_t padded = (n_cig + 3u) & ~3u;
    for (uint32_t k = lane; k < padded; k += 32) {
        uint32_t w = 0;
        if (k < n_cig) { const uint8_t* b = cg + 4ull * k; w = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); }
        cd[k] = w;
    }
    const uint8_t* sq = q + 32 + l_rn + 4ull * r.n_core[i]; uint8_t* sd = seq + r.seq_bytes[i];
    for (int64_t k = lane; k < (l_seq + 1) / 2; k += 32) sd[k] = sq[k];
    const uint32_t sl = r.sa_len[i];
    if (sl) { const uint8_t* ss = q + r.sa_src[i]; uint8_t* dd = sa + r.sa_bytes[i]; for (uint32_t k = lane; k < sl; k += 32) dd[k] = ss[k]; }
    uint8_t* nd = names + r.name_bytes[i];
    if (l_rn == 0) { if (lane == 0) nd[0] = 0; }
    else for (uint32_t k = lane; k < l_rn; k += 32) nd[k] = q[32 + k];
}

struct BamGpu {
    int device = 0; cudaStream_t st = nullptr;
    std::vector<void*> bufs;
    int64_t n = 0, cigar_words = 0, seq_bytes = 0, sa_bytes = 0, names_bytes = 0;
    BgRows r{}; uint32_t* cigar = nullptr; uint8_t* seq = nullptr; uint8_t* sa = nullptr; uint8_t* names = nullptr;
    std::string err;
    double ms_h2d = 0, ms_inflate = 0, ms_parse = 0;
    template <class T> T* alloc(size_t n_items) {
        void* p = nullptr;
        if (cudaMalloc(&p, (n_items ? n_items : 1) * sizeof(T) + 64) != cudaSuccess) return nullptr;
        bufs.push_back(p);
        return (T*)p;
    }
    ~BamGpu() { for (void* p : bufs) cudaFree(p); if (st) cudaStreamDestroy(st); }
};

#define BG_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(_e); return -1; } } while (0)
#define BG_ALLOC(var, T, n_items) do { var = h->alloc<T>((size_t)(n_items)); if (!var) { h->err = "cudaMalloc failed (" #var ")"; return -1; } } while (0)

template <class T>
static int bg_scan(BamGpu* h, T* d, int64_t n, void*& tmp, size_t& tmp_bytes) {      // exclusive sum in place over n + 1 items (last = total)
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d, d, (int)(n + 1), h->st);
    if (need > tmp_bytes) { BG_ALLOC(tmp, uint8_t, need); tmp_bytes = need; }
    BG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, need, d, d, (int)(n + 1), h->st));
    return 0;
}

extern "C" {

struct bamgpu_info { int64_t n_records, cigar_words, seq_bytes, sa_bytes, names_bytes; double ms_h2d, ms_inflate, ms_parse; };

const char* bamgpu_error(void* hh) { return hh ? ((BamGpu*)hh)->err.c_str() : "null handle"; }
void bamgpu_free(void* hh) { delete (BamGpu*)hh; }

// file: the whole .bam in host memory; blocks[n_blocks] = {payload offset, inflated offset, payload bytes, inflated bytes};
// first_record = offset of the first alignment record in the inflated stream (after the BAM header); n_ref = contigs.
// Returns a handle (also on failure, for bamgpu_error), rc in *rc: 0 ok, -1 CUDA, -2 malformed DEFLATE, -3 boundary guess failed, -4 bad record.
void* bamgpu_decode(const uint8_t* file, int64_t file_bytes, const BgBlock* blocks, int64_t n_blocks, int64_t first_record, int32_t n_ref, int device,
                    bamgpu_info* info, int* rc_out) {
    BamGpu* h = new BamGpu(); h->device = device;
    auto run = [&]() -> int {
        BG_CUDA(cudaSetDevice(device));
        BG_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
        cudaEvent_t ev[4]; for (auto& e : ev) cudaEventCreate(&e);
        const uint64_t usize = n_blocks ? blocks[n_blocks - 1].uoff + blocks[n_blocks - 1].ulen : 0;
        uint8_t* d_file; uint8_t* d_data; BgBlock* d_blocks; uint32_t* d_flags;
        BG_ALLOC(d_file, uint8_t, file_bytes); BG_ALLOC(d_data, uint8_t, usize + 64); BG_ALLOC(d_blocks, BgBlock, n_blocks); BG_ALLOC(d_flags, uint32_t, 8);
        cudaEventRecord(ev[0], h->st);
        BG_CUDA(cudaMemcpyAsync(d_file, file, (size_t)file_bytes, cudaMemcpyHostToDevice, h->st));
        BG_CUDA(cudaMemcpyAsync(d_blocks, blocks, (size_t)n_blocks * sizeof(BgBlock), cudaMemcpyHostToDevice, h->st));
        BG_CUDA(cudaMemsetAsync(d_flags, 0, 32, h->st));
        cudaEventRecord(ev[1], h->st);
        if (n_blocks) k_inflate<<<(unsigned)((n_blocks + BG_WARPS - 1) / BG_WARPS), 32 * BG_WARPS, 0, h->st>>>(d_file, d_blocks, n_blocks, d_data, d_flags);
        cudaEventRecord(ev[2], h->st);
        uint32_t flags[8];
        BG_CUDA(cudaMemcpyAsync(flags, d_flags, 32, cudaMemcpyDeviceToHost, h->st));
        BG_CUDA(cudaStreamSynchronize(h->st));
        BG_CUDA(cudaGetLastError());
        if (flags[0]) { h->err = "malformed DEFLATE data in a BGZF block (code " + std::to_string(flags[0]) + ")"; return -2; }
        // ---- record boundaries ----
        const uint64_t first = (uint64_t)first_record;
        const int64_t n_chunks = usize > first ? (int64_t)((usize - first + BG_CHUNK - 1) / BG_CHUNK) : 0;
        uint64_t* st; uint64_t* en; uint32_t* cnt; uint64_t* base;
        BG_ALLOC(st, uint64_t, n_chunks + 1); BG_ALLOC(en, uint64_t, n_chunks + 1); BG_ALLOC(cnt, uint32_t, n_chunks + 1); BG_ALLOC(base, uint64_t, n_chunks + 1);
        void* tmp = nullptr; size_t tmp_bytes = 0;
        int64_t n = 0;
        if (n_chunks) {
            k_starts<<<(unsigned)((n_chunks + 1 + 127) / 128), 128, 0, h->st>>>(d_data, usize, first, n_chunks, n_ref, st);
            k_chain<<<(unsigned)((n_chunks + 127) / 128), 128, 0, h->st>>>(d_data, usize, first, n_chunks, st, cnt, en, nullptr, nullptr);
            k_verify<<<(unsigned)((n_chunks + 127) / 128), 128, 0, h->st>>>(st, en, n_chunks, d_flags + 1);
            // base[c] = records before chunk c (64-bit scan of the 32-bit counts)
            size_t need = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, base, (int)(n_chunks + 1), h->st);
            BG_ALLOC(tmp, uint8_t, need); tmp_bytes = need;
            BG_CUDA(cudaMemsetAsync(cnt + n_chunks, 0, 4, h->st));
            BG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, need, cnt, base, (int)(n_chunks + 1), h->st));
            uint64_t total = 0;
            BG_CUDA(cudaMemcpyAsync(&total, base + n_chunks, 8, cudaMemcpyDeviceToHost, h->st));
            BG_CUDA(cudaMemcpyAsync(flags, d_flags, 32, cudaMemcpyDeviceToHost, h->st));
            BG_CUDA(cudaStreamSynchronize(h->st));
            BG_CUDA(cudaGetLastError());
            if (flags[1] & 2u) { h->err = "truncated BAM stream"; return -4; }
            if (flags[1] & 1u) { h->err = "a speculative record start was wrong (use the host decoder)"; return -3; }
            n = (int64_t)total;
        }
        h->n = n;
        uint64_t* rec_off; BG_ALLOC(rec_off, uint64_t, n + 1);
        if (n) k_chain<<<(unsigned)((n_chunks + 127) / 128), 128, 0, h->st>>>(d_data, usize, first, n_chunks, st, cnt, en, rec_off, base);
        BgRows& r = h->r;
        BG_ALLOC(r.tid, int32_t, n); BG_ALLOC(r.pos, int32_t, n); BG_ALLOC(r.flag, uint16_t, n); BG_ALLOC(r.mapq, uint8_t, n); BG_ALLOC(r.n_cigar, uint32_t, n);
        BG_ALLOC(r.l_seq, int32_t, n); BG_ALLOC(r.sa_len, uint32_t, n); BG_ALLOC(r.sa_src, uint32_t, n);
        BG_ALLOC(r.cig_src, uint32_t, n); BG_ALLOC(r.n_core, uint32_t, n);
        BG_ALLOC(r.cig_words, uint64_t, n + 1); BG_ALLOC(r.seq_bytes, uint64_t, n + 1); BG_ALLOC(r.sa_bytes, uint64_t, n + 1); BG_ALLOC(r.name_bytes, uint64_t, n + 1);
        uint64_t tot[4] = {0, 0, 0, 0};
        if (n) {
            k_rows<<<(unsigned)((n + 127) / 128), 128, 0, h->st>>>(d_data, rec_off, n, r, d_flags + 2);
            uint64_t* sized[4] = {r.cig_words, r.seq_bytes, r.sa_bytes, r.name_bytes};
            for (int k = 0; k < 4; ++k) {
                BG_CUDA(cudaMemsetAsync(sized[k] + n, 0, 8, h->st));
                if (bg_scan(h, sized[k], n, tmp, tmp_bytes)) return -1;
                BG_CUDA(cudaMemcpyAsync(&tot[k], sized[k] + n, 8, cudaMemcpyDeviceToHost, h->st));
            }
            BG_CUDA(cudaMemcpyAsync(flags, d_flags, 32, cudaMemcpyDeviceToHost, h->st));
            BG_CUDA(cudaStreamSynchronize(h->st));
            BG_CUDA(cudaGetLastError());
            if (flags[2]) { h->err = "corrupt alignment record"; return -4; }
        }
        h->cigar_words = (int64_t)tot[0]; h->seq_bytes = (int64_t)tot[1]; h->sa_bytes = (int64_t)tot[2]; h->names_bytes = (int64_t)tot[3];
        BG_ALLOC(h->cigar, uint32_t, tot[0]); BG_ALLOC(h->seq, uint8_t, tot[1]); BG_ALLOC(h->sa, uint8_t, tot[2]); BG_ALLOC(h->names, uint8_t, tot[3]);
        if (n) k_fill<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, h->st>>>(d_data, rec_off, n, r, h->cigar, h->seq, h->sa, h->names);
        cudaEventRecord(ev[3], h->st);
        BG_CUDA(cudaStreamSynchronize(h->st));
        BG_CUDA(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[0], ev[1]); h->ms_h2d = ms; cudaEventElapsedTime(&ms, ev[1], ev[2]); h->ms_inflate = ms;
        cudaEventElapsedTime(&ms, ev[2], ev[3]); h->ms_parse = ms;
        for (auto& e : ev) cudaEventDestroy(e);
        return 0;
    };
    const int rc = run();
    if (rc_out) *rc_out = rc;
    if (info) { info->n_records = h->n; info->cigar_words = h->cigar_words; info->seq_bytes = h->seq_bytes; info->sa_bytes = h->sa_bytes;
                info->names_bytes = h->names_bytes; info->ms_h2d = h->ms_h2d; info->ms_inflate = h->ms_inflate; info->ms_parse = h->ms_parse; }
    return h;
}

// D2H of the decoded records into caller arrays sized from bamgpu_info (offsets: cigar in words, seq / sa / names in bytes; n entries each).
int bamgpu_fetch(void* hh, int32_t* tid, int32_t* pos, uint16_t* flag, uint8_t* mapq, uint32_t* n_cigar, uint64_t* cigar_off, int32_t* l_seq, uint64_t* seq_off,
                 uint64_t* sa_off, uint32_t* sa_len, uint64_t* name_off, uint32_t* cigar, uint8_t* seq, uint8_t* sa, uint8_t* names) {
    BamGpu* h = (BamGpu*)hh;
    if (!h) return -1;
    const size_t n = (size_t)h->n;
    const BgRows& r = h->r;
    BG_CUDA(cudaSetDevice(h->device));
#define BG_D2H(dst, src, bytes) if ((bytes) && (dst)) BG_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, h->st))
    BG_D2H(tid, r.tid, n * 4); BG_D2H(pos, r.pos, n * 4); BG_D2H(flag, r.flag, n * 2); BG_D2H(mapq, r.mapq, n); BG_D2H(n_cigar, r.n_cigar, n * 4);
    BG_D2H(cigar_off, r.cig_words, n * 8); BG_D2H(l_seq, r.l_seq, n * 4); BG_D2H(seq_off, r.seq_bytes, n * 8); BG_D2H(sa_off, r.sa_bytes, n * 8);
    BG_D2H(sa_len, r.sa_len, n * 4); BG_D2H(name_off, r.name_bytes, n * 8);
    BG_D2H(cigar, h->cigar, (size_t)h->cigar_words * 4); BG_D2H(seq, h->seq, (size_t)h->seq_bytes); BG_D2H(sa, h->sa, (size_t)h->sa_bytes);
    BG_D2H(names, h->names, (size_t)h->names_bytes);
#undef BG_D2H
    BG_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

}  // extern "C"

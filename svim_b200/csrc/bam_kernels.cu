// Kernels of the on-GPU BAM decoder (bam.cu drives them):
//   k_inflate      one warp per BGZF block: lane 0 walks the Huffman stream (tables in shared memory) and stores literals, the warp
//                  copies the LZ77 matches from a small queue, in order
//   k_starts       one thread per 64 KiB chunk of the inflated stream: speculative first record start (bam_find_record_start)
//   k_chain        one thread per chunk: hop block_size fields to the next chunk's territory; k_verify: every chain must land
//                  on the next chunk's guess (then all guesses are right by induction from the header end)
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

#define BG_WARPS 4       // warps (= BGZF blocks) per CTA: 23 KB of Huffman tables + match queues, 8 CTAs per SM
#define BG_CHUNK (1ull << 16)

struct BgBlock { uint64_t coff; uint64_t uoff; uint32_t clen, ulen; };

// Warp-cooperative raw DEFLATE of one BGZF payload.  Huffman decoding is inherently serial: lane 0 walks the bit stream.
// What makes it fast enough is keeping that lane off the memory system:
//   * the compressed bytes come through a ring of 32-bit words in shared memory that the whole warp tops up with coalesced
//     loads (lane 0 refills its 64-bit bit buffer one aligned word at a time, no global load on its path);
//   * literals are stored as they are decoded; the last 8 output bytes also live in a register (`tail`), so the short,
//     near matches that make up most of a BAM stream (CIGAR words repeat at distance 4) are expanded by lane 0 itself without
//     reading anything back; matches that are long or reach further go to a small queue that the whole warp drains in order,
//     32 bytes per step, each match reading only bytes that are already final.
#define BG_QCAP 32
#define BG_WINW 128            // input ring: 128 words = 512 bytes per warp
struct BgMatch { uint32_t pos; uint16_t len, dist; };

struct BgReader {              // lane 0's bit reader over the ring
    uint64_t buf; uint32_t n;  // n valid bits in buf
    uint32_t aw;               // next word of the aligned stream to load
};
__device__ __forceinline__ void bgr_refill(BgReader& r, const uint32_t* win) {
    if (r.n <= 32) { r.buf |= (uint64_t)win[r.aw & (BG_WINW - 1)] << r.n; r.n += 32; ++r.aw; }
}
__device__ __forceinline__ uint32_t bgr_peek(const BgReader& r, uint32_t k) { return (uint32_t)r.buf & ((1u << k) - 1u); }
__device__ __forceinline__ void bgr_drop(BgReader& r, uint32_t k) { r.buf >>= k; r.n -= k; }
__device__ __forceinline__ uint32_t bgr_take(BgReader& r, uint32_t k) { const uint32_t v = bgr_peek(r, k); bgr_drop(r, k); return v; }
__device__ __forceinline__ uint32_t bgr_decode(BgReader& r, const uint32_t* table, int root) {
    uint32_t e = table[bgr_peek(r, (uint32_t)root)];
    if (((e >> 12) & 15u) == BGZF_K_SUB) { bgr_drop(r, e & 0xffu); e = table[(e >> 16) + bgr_peek(r, (e >> 8) & 15u)]; }
    if (((e >> 12) & 15u) == BGZF_K_INVALID) return 0;
    bgr_drop(r, e & 0xffu);
    return e;
}

// all lanes: ring words [fill_pos, upto) <- aligned stream (zeros past its end); returns the new fill position
__device__ __forceinline__ uint32_t bg_fill(uint32_t* win, const uint32_t* __restrict__ g32, uint32_t total_words, uint32_t fill_pos, uint32_t upto, int lane) {
    for (uint32_t w = fill_pos + lane; w < upto; w += 32) win[w & (BG_WINW - 1)] = w < total_words ? g32[w] : 0u;
    return upto;
}

__device__ int bg_inflate_warp(const uint8_t* __restrict__ src, uint32_t clen, uint8_t* __restrict__ dst, uint32_t ulen, uint32_t* tab, BgMatch* q, uint32_t* win, int lane,
                               unsigned long long* dbg, int inline_mode) {
    unsigned long long d_lit = 0, d_inl = 0, d_q = 0, d_pause = 0, d_far = 0, d_long = 0;
    uint32_t* lit = tab; uint32_t* dis = tab + BGZF_LIT_ENOUGH;
    const uint32_t A = (uint32_t)((uintptr_t)src & 3u);                      // the stream is read as aligned words starting A bytes before src
    const uint32_t* g32 = (const uint32_t*)(src - A);
    const uint32_t total_words = (A + clen + 3u) >> 2;
    uint32_t fill_pos = bg_fill(win, g32, total_words, 0u, BG_WINW, lane);
    __syncwarp();
    BgReader r; r.buf = 0; r.n = 0; r.aw = 0;
    if (lane == 0) { bgr_refill(r, win); bgr_drop(r, 8u * A); }
    uint32_t out = 0;
    uint64_t tail = 0; uint32_t tv = 0;                                      // lane 0: the last 8 output bytes (newest in the low byte), how many are valid
    for (;;) {
        // ---- block header (lane 0): 0 = stored, 1 = Huffman tables ready, >= 16: error code + 16 -----------------------------
        int kind = 0; uint32_t final_block = 0, st_p = 0, st_len = 0;
        if (lane == 0) {
            bgr_refill(r, win);
            final_block = bgr_take(r, 1);
            const uint32_t type = bgr_take(r, 2);
            if (type == 0) {
                bgr_drop(r, r.n & 7u);                                         // to the next byte boundary
                bgr_refill(r, win);
                const uint32_t len = bgr_take(r, 16);
                bgr_refill(r, win);
                const uint32_t nlen = bgr_take(r, 16);
                st_p = r.aw * 4u - (r.n >> 3) - A;                             // payload offset of the stored bytes
                st_len = len;
                if ((len ^ nlen) != 0xffffu) kind = 16 + BGZF_E_STORED;
                else if (st_p > clen || clen - st_p < len) kind = 16 + BGZF_E_INPUT;
                else if (ulen - out < len) kind = 16 + BGZF_E_OUTPUT;
            } else if (type == 1 || type == 2) {
                kind = 1;
                uint8_t lens[288 + 32];
                int hlit = 288, hdist = 32;
                if (type == 1) {
                    for (int i = 0; i < 144; ++i) lens[i] = 8;
                    for (int i = 144; i < 256; ++i) lens[i] = 9;
                    for (int i = 256; i < 280; ++i) lens[i] = 7;
                    for (int i = 280; i < 288; ++i) lens[i] = 8;
                    for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
                } else {
                    bgr_refill(r, win);
                    hlit = (int)bgr_take(r, 5) + 257; hdist = (int)bgr_take(r, 5) + 1;
                    const int hclen = (int)bgr_take(r, 4) + 4;
                    if (hlit > 286 || hdist > 30) kind = 16 + BGZF_E_HEADER;
                    else {
                        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                        uint8_t cl[19];
                        for (int i = 0; i < 19; ++i) cl[i] = 0;
                        for (int i = 0; i < hclen; ++i) { bgr_refill(r, win); cl[order[i]] = (uint8_t)bgr_take(r, 3); }
                        if (!bgzf_build(lit, BGZF_LIT_ENOUGH, 7, cl, 19, 2)) kind = 16 + BGZF_E_CODE;
                        int i = 0;
                        while (kind == 1 && i < hlit + hdist) {                 // <= 316 symbols of <= 14 bits: within the ring topped up before the header
                            bgr_refill(r, win);
                            const uint32_t e = bgr_decode(r, lit, 7);
                            if (!e) { kind = 16 + BGZF_E_SYMBOL; break; }
                            const uint32_t sym = e >> 16;
                            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                            uint32_t rep; uint8_t val = 0;
                            if (sym == 16) { if (i == 0) { kind = 16 + BGZF_E_HEADER; break; } val = lens[i - 1]; rep = 3 + bgr_take(r, 2); }
                            else if (sym == 17) rep = 3 + bgr_take(r, 3);
                            else rep = 11 + bgr_take(r, 7);
                            if (i + (int)rep > hlit + hdist) { kind = 16 + BGZF_E_HEADER; break; }
                            while (rep--) lens[i++] = val;
                        }
                        if (kind == 1 && lens[256] == 0) kind = 16 + BGZF_E_HEADER;
                        if (kind == 1) {
                            uint8_t dl[32];
                            for (int k = 0; k < hdist; ++k) dl[k] = lens[hlit + k];
                            for (int k = hlit; k < 288; ++k) lens[k] = 0;
                            for (int k = 0; k < 32; ++k) lens[288 + k] = k < hdist ? dl[k] : 0;
                        }
                    }
                }
                if (kind == 1 && !bgzf_build(lit, BGZF_LIT_ENOUGH, BGZF_LIT_ROOT, lens, type == 1 ? 288 : hlit, 0)) kind = 16 + BGZF_E_CODE;
                if (kind == 1 && !bgzf_build(dis, BGZF_DIST_ENOUGH, BGZF_DIST_ROOT, lens + 288, type == 1 ? 32 : hdist, 1)) kind = 16 + BGZF_E_CODE;
            } else kind = 16 + BGZF_E_HEADER;
        }
        kind = __shfl_sync(0xffffffffu, kind, 0);
        final_block = __shfl_sync(0xffffffffu, final_block, 0);
        if (kind >= 16) return kind - 16;
        if (kind == 0) {                                   // stored block: plain copy by the whole warp, then the reader restarts behind it
            st_p = __shfl_sync(0xffffffffu, st_p, 0); st_len = __shfl_sync(0xffffffffu, st_len, 0);
            for (uint32_t k = lane; k < st_len; k += 32) dst[out + k] = src[st_p + k];
            out += st_len;
            const uint32_t p = st_p + st_len + A;                              // aligned-stream byte position of the next block header
            __syncwarp();
            fill_pos = p >> 2;
            fill_pos = bg_fill(win, g32, total_words, fill_pos, (p >> 2) + BG_WINW, lane);
            __syncwarp();
            if (lane == 0) { r.buf = 0; r.n = 0; r.aw = p >> 2; bgr_refill(r, win); bgr_drop(r, 8u * (p & 3u)); tv = 0; }
        } else {
            // ---- symbols: lane 0 decodes; the warp tops up the input ring and drains the match queue between its runs -----------------
            for (;;) {
                int nq = 0, state = 0;                     // state: 0 pause (queue full / input low / tail needed), 1 end of block, >= 16 error
                if (lane == 0) {
                    for (;;) {
                        if (r.aw + 3u > fill_pos) break;                      // less than two symbols' worth of input left in the ring
                        bgr_refill(r, win);
                        const uint32_t e = bgr_decode(r, lit, BGZF_LIT_ROOT);
                        if (!e) { state = 16 + BGZF_E_SYMBOL; break; }
                        const uint32_t k = (e >> 12) & 15u;
                        if (k == BGZF_K_LITERAL) {
                            if (out >= ulen) { state = 16 + BGZF_E_OUTPUT; break; }
                            const uint32_t c = e >> 16;
                            dst[out++] = (uint8_t)c;
                            tail = (tail << 8) | c; tv += tv < 8u; ++d_lit;
                            continue;
                        }
                        if (k == BGZF_K_EOB) { state = 1; break; }
                        const uint32_t length = (e >> 16) + bgr_take(r, (e >> 8) & 15u);
                        bgr_refill(r, win);
                        const uint32_t o = bgr_decode(r, dis, BGZF_DIST_ROOT);
                        if (!o) { state = 16 + BGZF_E_SYMBOL; break; }
                        const uint32_t xb = (o >> 8) & 15u;
                        if (xb > 8u) bgr_refill(r, win);                       // up to 13 extra bits
                        const uint32_t dist = (o >> 16) + bgr_take(r, xb);
                        if (dist > out) { state = 16 + BGZF_E_DISTANCE; break; }
                        if (length > ulen - out) { state = 16 + BGZF_E_OUTPUT; break; }
                        if (inline_mode && dist <= 8u && length <= 24u && dist <= tv) {
                            // near and short: expand from the register copy of the last bytes (valid: nothing stale is queued before it)
                            const uint32_t sh = 8u * (dist - 1u);
                            for (uint32_t j = 0; j < length; ++j) { const uint32_t c = (uint32_t)(tail >> sh) & 0xffu; dst[out + j] = (uint8_t)c; tail = (tail << 8) | c; }
                            out += length; tv = tv + length < 8u ? tv + length : 8u; ++d_inl;
                            continue;
                        }
                        q[nq].pos = out; q[nq].len = (uint16_t)length; q[nq].dist = (uint16_t)dist;
                        out += length; ++d_q; d_far += dist > 8u; d_long += length > 24u;
                        if (dist == 1u && length >= 8u && tv >= 1u) { const uint64_t c = tail & 0xffu; tail = c * 0x0101010101010101ull; tv = 8u; }   // a run: the tail is known
                        else tv = 0;                                           // the tail is stale until the queue has been drained
                        if (++nq == BG_QCAP || (inline_mode && tv == 0)) break;
                    }
                }
                ++d_pause;
                nq = __shfl_sync(0xffffffffu, nq, 0); state = __shfl_sync(0xffffffffu, state, 0);
                __syncwarp();                              // lane 0's literals and queue entries are visible to the warp
                for (int j = 0; j < nq; ++j) {
                    const BgMatch m = q[j];
                    const uint32_t from = m.pos - m.dist;
                    if (m.dist >= m.len) { for (uint32_t k = lane; k < m.len; k += 32) dst[m.pos + k] = dst[from + k]; }
                    else if (m.dist == 1) { const uint8_t v = dst[from]; for (uint32_t k = lane; k < m.len; k += 32) dst[m.pos + k] = v; }
                    else { for (uint32_t k = lane; k < m.len; k += 32) dst[m.pos + k] = dst[from + k % m.dist]; }
                    __syncwarp();                          // the next match may read what this one wrote
                }
                if (state >= 16) return state - 16;
                // top up the input ring behind the reader; reload the tail if the queue made it stale
                const uint32_t aw = __shfl_sync(0xffffffffu, r.aw, 0);
                fill_pos = bg_fill(win, g32, total_words, fill_pos, aw + BG_WINW - 2u, lane);
                __syncwarp();
                if (lane == 0 && inline_mode && tv == 0 && nq > 0) {
                    const uint32_t have = out < 8u ? out : 8u;
                    tail = 0;
                    for (uint32_t j = 0; j < have; ++j) tail |= (uint64_t)dst[out - 1u - j] << (8u * j);
                    tv = have;
                }
                if (state == 1) break;
            }
            out = __shfl_sync(0xffffffffu, out, 0);
            // a full ring in front of the next block header (its code-length section can be ~300 bytes)
            const uint32_t aw = __shfl_sync(0xffffffffu, r.aw, 0);
            fill_pos = bg_fill(win, g32, total_words, fill_pos, aw + BG_WINW - 2u, lane);
            __syncwarp();
        }
        if (final_block) break;
    }
    if (dbg && lane == 0) { atomicAdd(dbg, d_lit); atomicAdd(dbg + 1, d_inl); atomicAdd(dbg + 2, d_q); atomicAdd(dbg + 3, d_pause); atomicAdd(dbg + 4, d_far); atomicAdd(dbg + 5, d_long); }
    int rc = BGZF_OK;
    if (lane == 0) {
        const uint64_t consumed = (uint64_t)r.aw * 32u - r.n - 8u * A;         // bits taken from the payload
        rc = consumed > 8ull * clen ? BGZF_E_INPUT : (out == ulen ? BGZF_OK : BGZF_E_OUTPUT);
    }
    return __shfl_sync(0xffffffffu, rc, 0);
}

__global__ void __launch_bounds__(32 * BG_WARPS) k_inflate(const uint8_t* __restrict__ file, const BgBlock* __restrict__ blocks, int64_t n_blocks,
                                                            uint8_t* __restrict__ out, uint32_t* __restrict__ status, int inline_mode) {
    __shared__ uint32_t tab[BG_WARPS][BGZF_TABLE_WORDS];
    __shared__ BgMatch queue[BG_WARPS][BG_QCAP];
    __shared__ uint32_t ring[BG_WARPS][BG_WINW];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * BG_WARPS + wib;
    if (b >= n_blocks) return;
    const BgBlock bl = blocks[b];
    const int rc = bg_inflate_warp(file + bl.coff, bl.clen, out + bl.uoff, bl.ulen, tab[wib], queue[wib], ring[wib], lane, (unsigned long long*)(status + 8), inline_mode);
    if (rc && lane == 0) atomicMax(status, (uint32_t)rc);
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

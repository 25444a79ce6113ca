// Host+device core of the on-GPU BAM decoder (bam.cu / bam_kernels.cu).
//
// Host+device (SVIM_HD) building blocks, written so that one GPU thread (lane 0 of a warp that owns a BGZF block, tables in
// shared memory) and the host test harness run the same code:
//   * bgzf_inflate_block   raw DEFLATE (RFC 1951) of one BGZF payload with the exact output size known from ISIZE; canonical
//                          Huffman tables with zlib's geometry (9-bit literal/length root, 6-bit distance root, 852 + 592 entries
//                          = 5.8 KB, i.e. 32 warps of tables per SM); byte-exact, never reads or writes out of bounds, returns
//                          an error code on malformed input;
//   * bam_record_plausible structural test of "a BAM record starts here" used to find record starts inside an inflated stream
//                          without a sequential pass (a chunk's speculative start is accepted only if the chain of the previous
//                          chunk lands on it, so the test only has to be selective, not infallible).
// tests/hostcheck builds this with g++ and checks it against zlib and against the true record starts of synthetic BAM files.
#pragma once
#include <stdint.h>
#include <string.h>

#ifndef SVIM_HD
#ifdef __CUDACC__
#define SVIM_HD __host__ __device__ __forceinline__
#else
#define SVIM_HD inline
#endif
#endif

#define BGZF_LIT_ROOT 9
#define BGZF_DIST_ROOT 6
#define BGZF_LIT_ENOUGH 852      // zlib ENOUGH_LENS: 286 symbols, root 9, max length 15
#define BGZF_DIST_ENOUGH 592     // zlib ENOUGH_DISTS: 30 symbols, root 6, max length 15
#define BGZF_TABLE_WORDS (BGZF_LIT_ENOUGH + BGZF_DIST_ENOUGH)

// table entry (32 bit): bits 0..7 = bits consumed at this level, bits 8..11 = extra bits (length / distance) or sub-table
// index bits (pointer), bits 12..15 = kind, bits 16..31 = literal / base value / sub-table start
enum { BGZF_K_INVALID = 0, BGZF_K_LITERAL = 1, BGZF_K_BASE = 2, BGZF_K_EOB = 3, BGZF_K_SUB = 4 };
enum { BGZF_OK = 0, BGZF_E_HEADER = 1, BGZF_E_CODE = 2, BGZF_E_SYMBOL = 3, BGZF_E_DISTANCE = 4, BGZF_E_OUTPUT = 5, BGZF_E_INPUT = 6, BGZF_E_STORED = 7 };

struct BgzfBits {
    const uint8_t* src; uint32_t len, pos;      // next byte to load
    uint64_t buf; uint32_t n;                   // n valid bits in buf
    uint32_t pad;                               // zero bytes fed past the end
};

SVIM_HD void bgzf_refill(BgzfBits& b) {
    while (b.n <= 56) {
        if (b.pos < b.len) b.buf |= (uint64_t)b.src[b.pos++] << b.n; else ++b.pad;
        b.n += 8;
    }
}
SVIM_HD uint32_t bgzf_peek(const BgzfBits& b, uint32_t k) { return (uint32_t)(b.buf & ((1ull << k) - 1ull)); }
SVIM_HD void bgzf_drop(BgzfBits& b, uint32_t k) { b.buf >>= k; b.n -= k; }
SVIM_HD uint32_t bgzf_take(BgzfBits& b, uint32_t k) { const uint32_t v = bgzf_peek(b, k); bgzf_drop(b, k); return v; }

SVIM_HD uint32_t bgzf_rev(uint32_t v, int k) {
    uint32_t r = 0;
    for (int i = 0; i < k; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

SVIM_HD uint32_t bgzf_len_base(uint32_t i) {      // length symbols 257..285 -> i = 0..28
    const uint16_t t[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    return t[i];
}
SVIM_HD uint32_t bgzf_len_extra(uint32_t i) { return (i < 8 || i == 28) ? 0u : (i - 4) >> 2; }
SVIM_HD uint32_t bgzf_dist_extra(uint32_t i) { return i < 4 ? 0u : (i - 2) >> 1; }
SVIM_HD uint32_t bgzf_dist_base(uint32_t i) { return i < 4 ? i + 1 : ((2u + (i & 1u)) << bgzf_dist_extra(i)) + 1u; }

SVIM_HD uint32_t bgzf_entry(int kind_alphabet, uint32_t sym) {      // 0 literal/length, 1 distance, 2 code lengths
    if (kind_alphabet == 0) {
        if (sym < 256) return (sym << 16) | (BGZF_K_LITERAL << 12);
        if (sym == 256) return BGZF_K_EOB << 12;
        if (sym > 285) return 0;
        return (bgzf_len_base(sym - 257) << 16) | (BGZF_K_BASE << 12) | (bgzf_len_extra(sym - 257) << 8);
    }
    if (kind_alphabet == 1) {
        if (sym > 29) return 0;
        return (bgzf_dist_base(sym) << 16) | (BGZF_K_BASE << 12) | (bgzf_dist_extra(sym) << 8);
    }
    return (sym << 16) | (BGZF_K_LITERAL << 12);
}

// Canonical Huffman decode table.  false: over-subscribed, or incomplete other than a single 1-bit code (zlib's rule).
SVIM_HD bool bgzf_build(uint32_t* table, uint32_t cap, int root, const uint8_t* lens, int n_syms, int alphabet) {
    int count[16];
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int s = 0; s < n_syms; ++s) count[lens[s]]++;
    count[0] = 0;
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) --max_len;
    const uint32_t primary = 1u << root;
    for (uint32_t i = 0; i < primary; ++i) table[i] = 0;
    if (max_len == 0) return alphabet == 1;          // no distance codes at all: legal for an all-literal block
    int left = 1, used = 0;
    for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; used += count[l]; }
    if (left > 0 && (alphabet == 2 || !(used == 1 && count[1] == 1))) return false;
    uint32_t next[16]; uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    uint32_t next_free = primary;
    // symbols in canonical order: increasing length, then increasing symbol; long codes with the same root prefix are
    // contiguous in that order and the longest comes last, so a sub-table is sized when its prefix is first seen by looking
    // ahead over the remaining codes of that prefix (counts per length are enough: zlib's inflate_table does the same)
    for (int l = 1; l <= max_len; ++l) {
        for (int s = 0; s < n_syms; ++s) {
            if (lens[s] != l) continue;
            const uint32_t c = next[l]++;
            const uint32_t e = bgzf_entry(alphabet, (uint32_t)s);
            const uint32_t rev = bgzf_rev(c, l);
            if (l <= root) {
                for (uint32_t i = rev; i < primary; i += 1u << l) table[i] = e ? (e | (uint32_t)l) : 0u;
                continue;
            }
            const uint32_t pre = rev & (primary - 1u);
            if (((table[pre] >> 12) & 15u) != BGZF_K_SUB) {
                // size: smallest `bits` such that the codes of lengths l, l+1, ... still to come under this prefix fill 2^bits
                int bits = l - root; int room = 1 << bits;
                // codes of length l not yet assigned (including this one), then longer ones
                int remaining_l = 0;
                for (int t = s; t < n_syms; ++t) remaining_l += lens[t] == l;
                room -= remaining_l;
                int ll = l;
                while (room > 0 && ll < max_len) { ++ll; ++bits; room = (room << 1) - count[ll]; }
                if (next_free + (1u << bits) > cap) return false;
                table[pre] = (next_free << 16) | (BGZF_K_SUB << 12) | ((uint32_t)bits << 8) | (uint32_t)root;
                for (uint32_t i = 0; i < (1u << bits); ++i) table[next_free + i] = 0;
                next_free += 1u << bits;
            }
            const uint32_t start = table[pre] >> 16, bits = (table[pre] >> 8) & 15u;
            const int ll = l - root;
            if ((uint32_t)ll > bits) return false;
            for (uint32_t i = rev >> root; i < (1u << bits); i += 1u << ll) table[start + i] = e ? (e | (uint32_t)ll) : 0u;
        }
    }
    return true;
}

// one symbol through a root table with optional second level; returns the entry (kind INVALID on a hole), consumes its bits
SVIM_HD uint32_t bgzf_decode(BgzfBits& b, const uint32_t* table, int root) {
    uint32_t e = table[bgzf_peek(b, (uint32_t)root)];
    if (((e >> 12) & 15u) == BGZF_K_SUB) {
        bgzf_drop(b, e & 0xffu);
        e = table[(e >> 16) + bgzf_peek(b, (e >> 8) & 15u)];
    }
    if (((e >> 12) & 15u) == BGZF_K_INVALID) return 0;
    bgzf_drop(b, e & 0xffu);
    return e;
}

// Inflate one raw DEFLATE stream whose output size is known.  `tab`: BGZF_TABLE_WORDS words of scratch (shared memory on the GPU).
SVIM_HD int bgzf_inflate_block(const uint8_t* src, uint32_t clen, uint8_t* dst, uint32_t ulen, uint32_t* tab) {
    BgzfBits b; b.src = src; b.len = clen; b.pos = 0; b.buf = 0; b.n = 0; b.pad = 0;
    uint32_t* lit = tab; uint32_t* dis = tab + BGZF_LIT_ENOUGH;
    uint32_t out = 0;
    for (;;) {
        bgzf_refill(b);
        const uint32_t final_block = bgzf_take(b, 1), type = bgzf_take(b, 2);
        if (type == 0) {
            bgzf_drop(b, b.n & 7u);
            bgzf_refill(b);
            const uint32_t len = bgzf_take(b, 16), nlen = bgzf_take(b, 16);
            const uint32_t held = b.n >> 3;                       // whole bytes still buffered; the last `pad` of them are padding
            if ((len ^ nlen) != 0xffffu || b.pad > held) return BGZF_E_STORED;
            const uint32_t p = b.pos - (held - b.pad);
            if (clen - p < len) return BGZF_E_INPUT;
            if (ulen - out < len) return BGZF_E_OUTPUT;
            for (uint32_t k = 0; k < len; ++k) dst[out + k] = src[p + k];
            out += len;
            b.pos = p + len; b.buf = 0; b.n = 0; b.pad = 0;
        } else if (type == 1 || type == 2) {
            uint8_t lens[288 + 32];
            int hlit = 288, hdist = 32;
            if (type == 1) {
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
            } else {
                bgzf_refill(b);
                hlit = (int)bgzf_take(b, 5) + 257; hdist = (int)bgzf_take(b, 5) + 1;
                const int hclen = (int)bgzf_take(b, 4) + 4;
                if (hlit > 286 || hdist > 30) return BGZF_E_HEADER;
                const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19];
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < hclen; ++i) { bgzf_refill(b); cl[order[i]] = (uint8_t)bgzf_take(b, 3); }
                if (!bgzf_build(lit, BGZF_LIT_ENOUGH, 7, cl, 19, 2)) return BGZF_E_CODE;       // 128-entry table in the literal area
                int i = 0;
                while (i < hlit + hdist) {
                    bgzf_refill(b);
                    const uint32_t e = bgzf_decode(b, lit, 7);
                    if (!e) return BGZF_E_SYMBOL;
                    const uint32_t sym = e >> 16;
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    uint32_t rep; uint8_t val = 0;
                    if (sym == 16) { if (i == 0) return BGZF_E_HEADER; val = lens[i - 1]; rep = 3 + bgzf_take(b, 2); }
                    else if (sym == 17) rep = 3 + bgzf_take(b, 3);
                    else rep = 11 + bgzf_take(b, 7);
                    if (i + (int)rep > hlit + hdist) return BGZF_E_HEADER;
                    while (rep--) lens[i++] = val;
                }
                if (lens[256] == 0) return BGZF_E_HEADER;
                uint8_t dl[32];
                for (int k = 0; k < hdist; ++k) dl[k] = lens[hlit + k];
                for (int k = hlit; k < 288; ++k) lens[k] = 0;
                for (int k = 0; k < 32; ++k) lens[288 + k] = k < hdist ? dl[k] : 0;
            }
            if (!bgzf_build(lit, BGZF_LIT_ENOUGH, BGZF_LIT_ROOT, lens, type == 1 ? 288 : hlit, 0)) return BGZF_E_CODE;
            if (!bgzf_build(dis, BGZF_DIST_ENOUGH, BGZF_DIST_ROOT, lens + 288, type == 1 ? 32 : hdist, 1)) return BGZF_E_CODE;
            for (;;) {
                bgzf_refill(b);
                const uint32_t e = bgzf_decode(b, lit, BGZF_LIT_ROOT);
                if (!e) return BGZF_E_SYMBOL;
                const uint32_t kind = (e >> 12) & 15u;
                if (kind == BGZF_K_LITERAL) {
                    if (out >= ulen) return BGZF_E_OUTPUT;
                    dst[out++] = (uint8_t)(e >> 16);
                    continue;
                }
                if (kind == BGZF_K_EOB) break;
                const uint32_t length = (e >> 16) + bgzf_take(b, (e >> 8) & 15u);
                bgzf_refill(b);
                const uint32_t o = bgzf_decode(b, dis, BGZF_DIST_ROOT);
                if (!o) return BGZF_E_SYMBOL;
                const uint32_t dist = (o >> 16) + bgzf_take(b, (o >> 8) & 15u);
                if (dist > out) return BGZF_E_DISTANCE;
                if (length > ulen - out) return BGZF_E_OUTPUT;
                for (uint32_t k = 0; k < length; ++k) dst[out + k] = dst[out + k - dist];
                out += length;
                if (b.pad > 8) return BGZF_E_INPUT;
            }
        } else return BGZF_E_HEADER;
        if (final_block) break;
    }
    if (b.pad * 8u > b.n) return BGZF_E_INPUT;      // consumed bits that were never in the input
    return out == ulen ? BGZF_OK : BGZF_E_OUTPUT;
}

// ---- record starts inside an inflated BAM stream ---------------------------------------------------------------------
// Does a BAM alignment record plausibly start at p (p points at its block_size field)?  Structural checks only (SAMv1 §4.2):
// sizes consistent, reference ids in range, name NUL-terminated and printable, CIGAR op codes valid.
SVIM_HD bool bam_record_plausible(const uint8_t* p, uint64_t avail, int32_t n_ref) {
    if (avail < 36) return false;
    uint32_t bs; memcpy(&bs, p, 4);
    if (bs < 32 || bs > (1u << 28)) return false;
    int32_t ref_id, pos, l_seq, next_ref, next_pos; uint16_t n_cig; 
    memcpy(&ref_id, p + 4, 4); memcpy(&pos, p + 8, 4);
    const uint32_t l_rn = p[12];
    memcpy(&n_cig, p + 16, 2); memcpy(&l_seq, p + 20, 4); memcpy(&next_ref, p + 24, 4); memcpy(&next_pos, p + 28, 4);
    if (ref_id < -1 || ref_id >= n_ref || next_ref < -1 || next_ref >= n_ref) return false;
    if (pos < -1 || next_pos < -1 || l_seq < 0 || l_rn < 1) return false;
    const uint64_t need = 32ull + l_rn + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (need > bs) return false;
    const uint64_t visible = avail - 4 < bs ? avail - 4 : bs;      // bytes of the record that can be inspected
    const uint8_t* r = p + 4;
    if (32 + (uint64_t)l_rn <= visible) {
        if (r[32 + l_rn - 1] != 0) return false;
        for (uint32_t k = 0; k + 1 < l_rn; ++k) { const uint8_t c = r[32 + k]; if (c < 33 || c > 126) return false; }
        const uint64_t cig_vis = (visible - 32 - l_rn) / 4 < n_cig ? (visible - 32 - l_rn) / 4 : n_cig;
        for (uint64_t k = 0; k < cig_vis && k < 16; ++k) { uint32_t w; memcpy(&w, r + 32 + l_rn + 4 * k, 4); if ((w & 15u) > 8u) return false; }
    }
    return true;
}

// First offset >= from at which `depth` consecutive plausible records start (the last ones may be cut by `size`); size if none.
SVIM_HD uint64_t bam_find_record_start(const uint8_t* data, uint64_t size, uint64_t from, int32_t n_ref, int depth) {
    for (uint64_t o = from; o + 36 <= size; ++o) {
        uint64_t q = o; int ok = 0;
        while (ok < depth && q + 36 <= size && bam_record_plausible(data + q, size - q, n_ref)) {
            uint32_t bs; memcpy(&bs, data + q, 4);
            q += 4ull + bs; ++ok;
        }
        if (ok == depth || (ok > 0 && q + 36 > size)) return o;
    }
    return size;
}

// Hop over block_size fields from a record start until the first record that starts at or after `limit` (the next chunk's
// territory); counts the records that start before `limit`.  Returns that offset, `size` at the end of the stream, or
// UINT64_MAX when a record is cut by the end of the stream (truncated file).
SVIM_HD uint64_t bam_chain(const uint8_t* data, uint64_t size, uint64_t start, uint64_t limit, uint32_t* n_records) {
    uint64_t o = start; uint32_t n = 0;
    while (o < limit && o < size) {
        if (o + 4 > size) return ~0ull;
        uint32_t bs; memcpy(&bs, data + o, 4);
        if (bs < 32 || o + 4ull + bs > size) return ~0ull;
        o += 4ull + bs; ++n;
    }
    *n_records = n;
    return o;
}

// Multi-threaded BAM -> flattened record buffer (svim_aln_soa) decoder.
//
// once the kernels run at TB/s the end-to-end rate is set by host BAM
// decompression.  The reference iterates pysam/htslib records one by one (SVIM_COLLECT.py:133); this reader
// inflates all BGZF blocks in parallel (zlib raw inflate, one task per block), indexes the records with one
// sequential hop over the block_size fields, and fills the structure-of-arrays + CIGAR / SEQ / SA blobs in
// parallel.  Host code only — nothing here runs on the GPU path's timed kernels.
//
// File format: SAMv1 §4.1 (BGZF) and §4.2 (BAM).
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <memory>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

extern "C" {

struct bamio_info {
    int64_t n_records, cigar_words, seq_bytes, sa_bytes, n_qnames, names_bytes;
    int32_t n_contigs, sorted_coordinate;
};

struct bamio_out {
    int32_t* tid; int32_t* pos; uint16_t* flag; uint8_t* mapq; uint32_t* n_cigar; uint64_t* cigar_off;
    int32_t* l_seq; uint64_t* seq_off; uint64_t* sa_off; uint32_t* sa_len; uint32_t* qname_id;
    uint32_t* cigar; uint8_t* seq; uint8_t* sa;
};

}  // extern "C"

namespace {

struct Block { size_t coff, clen, uoff, ulen; };

struct Rec { size_t off; uint32_t sa_off_in_rec, sa_len; };

// uninitialised byte buffer: pages are first touched by the worker threads, not zero-filled by one thread
// Large first-touch buffers are page-fault bound with 4 KiB pages; ask for transparent huge pages.
static void advise_huge(void* p, size_t bytes) {
#ifdef MADV_HUGEPAGE
    const uintptr_t a = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
    const uintptr_t e = ((uintptr_t)p + bytes) & ~(uintptr_t)((2u << 20) - 1);
    if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
#endif
}

struct RawBuf {
    uint8_t* ptr = nullptr; size_t n = 0;
    RawBuf() = default;
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { release(); }
    void release() { if (ptr) free(ptr); ptr = nullptr; n = 0; }
    void alloc(size_t bytes) {
        release();
        void* q = nullptr;
        if (posix_memalign(&q, (size_t)2 << 20, bytes ? bytes : 1) != 0) q = nullptr;
        ptr = (uint8_t*)q; n = bytes;
        if (ptr) advise_huge(ptr, bytes);
    }
    uint8_t* data() { return ptr; }
    const uint8_t* data() const { return ptr; }
    size_t size() const { return n; }
};

struct Handle {
    RawBuf data;                    // inflated stream
    std::vector<Rec> recs;
    std::vector<std::string> contigs; std::vector<int64_t> contig_len;
    std::vector<uint32_t> qid;
    std::vector<std::string_view> qnames;
    std::string sort_order;
    bamio_info info;
    int threads = 1;
};

// SVIM_BAMIO_TRACE=1: per-phase wall times on stderr
struct Trace {
    bool on; std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("SVIM_BAMIO_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bamio] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

template <class F>
void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n < 64) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([=] { f(n * t / threads, n * (t + 1) / threads); });
    for (auto& x : th) x.join();
}

// SA:Z payload inside the aux area -> (offset relative to record start, length), walking typed fields
bool find_sa(const uint8_t* rec, size_t aux_begin, size_t rec_len, uint32_t& off, uint32_t& len) {
    size_t o = aux_begin;
    while (o + 3 <= rec_len) {
        const uint8_t t0 = rec[o], t1 = rec[o + 1], ty = rec[o + 2];
        o += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': {
                size_t e = o;
                while (e < rec_len && rec[e]) ++e;
                if (t0 == 'S' && t1 == 'A' && ty == 'Z') { off = (uint32_t)o; len = (uint32_t)(e - o); return true; }
                o = e + 1;
                continue;
            }
            case 'B': {
                if (o + 5 > rec_len) return false;
                const uint8_t sub = rec[o]; const uint32_t cnt = rd32(rec + o + 1);
                size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                o += 5 + (size_t)cnt * es;
                continue;
            }
            default: return false;
        }
        o += sz;
    }
    return false;
}

}  // namespace

extern "C" {

void* bamio_open(const char* path, int n_threads, bamio_info* info, char* err, int errcap) {
    auto fail = [&](const char* m) -> void* { if (err && errcap > 0) snprintf(err, errcap, "%s", m); return nullptr; };
    Trace tr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail("cannot open file");
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return fail("cannot stat file"); }
    const size_t fsz = (size_t)sb.st_size;
    RawBuf file; file.alloc(fsz);
    if (fsz && !file.data()) { close(fd); return fail("out of memory"); }
    {   // parallel pread: the file usually sits in the page cache, the copy is what costs
        std::atomic<int> short_read{0};
        const size_t CH = (size_t)64 << 20;
        parallel_for((fsz + CH - 1) / CH, std::max(1, n_threads), [&](size_t lo, size_t hi) {
            for (size_t c = lo; c < hi; ++c) {
                size_t off = c * CH, len = std::min(CH, fsz - off);
                while (len) { const ssize_t r = pread(fd, file.data() + off, len, (off_t)off); if (r <= 0) { short_read = 1; return; } off += (size_t)r; len -= (size_t)r; }
            }
        });
        close(fd);
        if (short_read) return fail("short read");
    }
    tr.mark("read file");
    // ---- BGZF block index ----------------------------------------------------------------------------
    std::vector<Block> blocks;
    size_t o = 0, uoff = 0;
    while (o + 18 <= file.size()) {
        const uint8_t* p = file.data() + o;
        if (!(p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 4))) return fail("not a BGZF file");
        const uint16_t xlen = rd16(p + 10);
        size_t x = 12, xe = 12 + xlen; int bsize = -1;
        while (x + 4 <= xe) {
            const uint16_t slen = rd16(p + x + 2);
            if (p[x] == 66 && p[x + 1] == 67 && slen == 2) bsize = rd16(p + x + 4);
            x += 4 + slen;
        }
        if (bsize < 0 || o + (size_t)bsize + 1 > file.size()) return fail("bad BGZF block");
        const size_t total = (size_t)bsize + 1;
        const uint32_t isize = rd32(p + total - 4);
        blocks.push_back({o + 12 + xlen, total - xlen - 20, uoff, isize});
        uoff += isize; o += total;
    }
    tr.mark("block index");
    Handle* h = new Handle();
    h->threads = std::max(1, n_threads);
    h->data.alloc(uoff);
    if (uoff && !h->data.data()) { delete h; return fail("out of memory"); }
    std::atomic<int> bad{0};
    parallel_for(blocks.size(), h->threads, [&](size_t lo, size_t hi) {
        z_stream zs;
        for (size_t b = lo; b < hi; ++b) {
            if (blocks[b].ulen == 0) continue;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            zs.next_in = file.data() + blocks[b].coff; zs.avail_in = (uInt)blocks[b].clen;
            zs.next_out = h->data.data() + blocks[b].uoff; zs.avail_out = (uInt)blocks[b].ulen;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) bad = 1;
        }
    });
    if (bad) { delete h; return fail("inflate failed"); }
    tr.mark("inflate");
    file.release();
    tr.mark("free file buffer");
    // ---- header --------------------------------------------------------------------------------------------
    const uint8_t* d = h->data.data(); const size_t n = h->data.size();
    if (n < 12 || memcmp(d, "BAM\1", 4) != 0) { delete h; return fail("not a BAM stream"); }
    const uint32_t l_text = rd32(d + 4);
    std::string text((const char*)d + 8, strnlen((const char*)d + 8, l_text));
    h->sort_order = "unknown";
    {
        size_t p = text.find("@HD");
        if (p != std::string::npos) {
            size_t e = text.find('\n', p), s = text.find("SO:", p);
            if (s != std::string::npos && (e == std::string::npos || s < e)) {
                size_t t = s + 3; size_t q = t;
                while (q < text.size() && text[q] != '\t' && text[q] != '\n') ++q;
                h->sort_order = text.substr(t, q - t);
            }
        }
    }
    size_t p = 8 + l_text;
    const uint32_t n_ref = rd32(d + p); p += 4;
    for (uint32_t r = 0; r < n_ref; ++r) {
        const uint32_t ln = rd32(d + p); p += 4;
        h->contigs.emplace_back((const char*)d + p, ln ? ln - 1 : 0); p += ln;
        h->contig_len.push_back((int32_t)rd32(d + p)); p += 4;
    }
    // ---- record index (sequential hop) ---------------------------------------------------------------------------
    while (p + 4 <= n) {
        const uint32_t bs = rd32(d + p);
        if (p + 4 + bs > n) { delete h; return fail("truncated record"); }
        h->recs.push_back({p + 4, 0, 0});
        p += 4 + (size_t)bs;
    }
    const size_t nr = h->recs.size();
    tr.mark("record index (serial hop)");
    // ---- per-record sizes (parallel) ---------------------------------------------------------------------------------
    parallel_for(nr, h->threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const uint8_t* r = d + h->recs[i].off;
            const uint32_t bs = rd32(r - 4);
            const uint8_t l_rn = r[8]; const uint16_t n_cig = rd16(r + 12); const int32_t l_seq = (int32_t)rd32(r + 16);
            const size_t aux = 32 + (size_t)l_rn + 4 * (size_t)n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
            uint32_t so = 0, sl = 0;
            if (aux < bs) find_sa(r, aux, bs, so, sl);
            h->recs[i].sa_off_in_rec = so; h->recs[i].sa_len = sl;
        }
    });
    tr.mark("SA tag search");
    // ---- read-name ids ---------------------------------------------------------------------------------------------------
    h->qid.resize(nr);
    {
        std::unordered_map<std::string_view, uint32_t> ids;
        ids.reserve(nr * 2);
        for (size_t i = 0; i < nr; ++i) {
            const uint8_t* r = d + h->recs[i].off;
            std::string_view nm((const char*)r + 32, r[8] ? r[8] - 1 : 0);
            auto it = ids.find(nm);
            if (it == ids.end()) { it = ids.emplace(nm, (uint32_t)h->qnames.size()).first; h->qnames.push_back(nm); }
            h->qid[i] = it->second;
        }
    }
    tr.mark("read-name ids (serial hash)");
    bamio_info& inf = h->info;
    memset(&inf, 0, sizeof(inf));
    inf.n_records = (int64_t)nr; inf.n_contigs = (int32_t)n_ref; inf.n_qnames = (int64_t)h->qnames.size();
    inf.sorted_coordinate = h->sort_order == "coordinate";
    for (size_t i = 0; i < nr; ++i) {
        const uint8_t* r = d + h->recs[i].off;
        const uint16_t n_cig = rd16(r + 12); const int32_t l_seq = (int32_t)rd32(r + 16);
        inf.cigar_words += (n_cig + 3) & ~3; inf.seq_bytes += (l_seq + 1) / 2; inf.sa_bytes += h->recs[i].sa_len;
    }
    for (auto& q : h->qnames) inf.names_bytes += (int64_t)q.size() + 1;
    *info = inf;
    tr.mark("sizes (serial)");
    return h;
}

// contig names NUL-separated into names_out (cap bytes), lengths[n_contigs]; sort order into so (16 bytes)
int bamio_header(void* hh, char* names_out, int64_t cap, int64_t* lengths, char* so) {
    Handle* h = (Handle*)hh;
    int64_t o = 0;
    for (size_t i = 0; i < h->contigs.size(); ++i) {
        const std::string& s = h->contigs[i];
        if (o + (int64_t)s.size() + 1 > cap) return -1;
        memcpy(names_out + o, s.c_str(), s.size() + 1); o += (int64_t)s.size() + 1;
        lengths[i] = h->contig_len[i];
    }
    snprintf(so, 16, "%s", h->sort_order.c_str());
    return 0;
}

int bamio_qnames(void* hh, char* out, int64_t cap) {
    Handle* h = (Handle*)hh;
    int64_t o = 0;
    for (auto& q : h->qnames) {
        if (o + (int64_t)q.size() + 1 > cap) return -1;
        memcpy(out + o, q.data(), q.size()); out[o + q.size()] = 0; o += (int64_t)q.size() + 1;
    }
    return 0;
}

int bamio_fill(void* hh, bamio_out* out) {
    Handle* h = (Handle*)hh;
    Trace tr;
    advise_huge(out->cigar, (size_t)h->info.cigar_words * 4); advise_huge(out->seq, (size_t)h->info.seq_bytes);
    const uint8_t* d = h->data.data();
    const size_t nr = h->recs.size();
    uint64_t co = 0, so = 0, sao = 0;
    for (size_t i = 0; i < nr; ++i) {
        const uint8_t* r = d + h->recs[i].off;
        const uint16_t n_cig = rd16(r + 12); const int32_t l_seq = (int32_t)rd32(r + 16);
        out->cigar_off[i] = co; co += (n_cig + 3) & ~3;
        out->seq_off[i] = so; so += (uint64_t)(l_seq + 1) / 2;
        out->sa_off[i] = sao; sao += h->recs[i].sa_len;
    }
    tr.mark("offsets (serial)");
    parallel_for(nr, h->threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const uint8_t* r = d + h->recs[i].off;
            const uint8_t l_rn = r[8]; const uint16_t n_cig = rd16(r + 12); const int32_t l_seq = (int32_t)rd32(r + 16);
            out->tid[i] = (int32_t)rd32(r); out->pos[i] = (int32_t)rd32(r + 4); out->mapq[i] = r[9]; out->flag[i] = rd16(r + 14);
            out->n_cigar[i] = n_cig; out->l_seq[i] = l_seq; out->sa_len[i] = h->recs[i].sa_len; out->qname_id[i] = h->qid[i];
            const uint8_t* cg = r + 32 + l_rn;
            uint32_t* dst = out->cigar + out->cigar_off[i];
            memcpy(dst, cg, 4 * (size_t)n_cig);
            for (uint32_t k = n_cig; k < ((n_cig + 3u) & ~3u); ++k) dst[k] = 0;
            memcpy(out->seq + out->seq_off[i], cg + 4 * (size_t)n_cig, (size_t)(l_seq + 1) / 2);
            if (h->recs[i].sa_len) memcpy(out->sa + out->sa_off[i], r + h->recs[i].sa_off_in_rec, h->recs[i].sa_len);
        }
    });
    tr.mark("SoA fill");
    return 0;
}

void bamio_close(void* hh) { delete (Handle*)hh; }

// ---- writer: SoA -> BAM (BGZF blocks compressed in parallel).  Used to materialise the synthetic configs as files. ----
struct bamio_in {
    int64_t n; const int32_t* tid; const int32_t* pos; const uint16_t* flag; const uint8_t* mapq; const uint32_t* n_cigar; const uint64_t* cigar_off;
    const int32_t* l_seq; const uint64_t* seq_off; const uint64_t* sa_off; const uint32_t* sa_len; const uint32_t* qname_id;
    const uint32_t* cigar; const uint8_t* seq; const uint8_t* sa;
    const char* qnames; const int64_t* qname_off;          // may be null: names are "read<id>"
    int32_t n_contigs; const char* contig_names;           // NUL separated
    const int64_t* contig_len; const char* sort_order;
};

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

int bamio_write(const char* path, const bamio_in* in, int level, int n_threads) {
    std::vector<uint8_t> raw;
    auto put32 = [&](uint32_t v) { uint8_t b[4]; memcpy(b, &v, 4); raw.insert(raw.end(), b, b + 4); };
    auto put16 = [&](uint16_t v) { uint8_t b[2]; memcpy(b, &v, 2); raw.insert(raw.end(), b, b + 2); };
    std::string text = std::string("@HD\tVN:1.6\tSO:") + (in->sort_order ? in->sort_order : "unknown") + "\n";
    std::vector<std::string> names;
    const char* p = in->contig_names;
    for (int i = 0; i < in->n_contigs; ++i) { names.emplace_back(p); p += names.back().size() + 1; }
    for (int i = 0; i < in->n_contigs; ++i) text += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(in->contig_len[i]) + "\n";
    raw.insert(raw.end(), {'B', 'A', 'M', 1});
    put32((uint32_t)text.size()); raw.insert(raw.end(), text.begin(), text.end());
    put32((uint32_t)in->n_contigs);
    for (int i = 0; i < in->n_contigs; ++i) {
        put32((uint32_t)names[i].size() + 1); raw.insert(raw.end(), names[i].begin(), names[i].end()); raw.push_back(0);
        put32((uint32_t)in->contig_len[i]);
    }
    // record sizes -> offsets, then parallel fill
    const int64_t n = in->n;
    std::vector<size_t> roff(n + 1);
    std::vector<std::string> qn(in->qnames ? 0 : 0);
    size_t o = raw.size();
    auto name_len = [&](int64_t i) -> size_t {
        if (in->qnames) return (size_t)(in->qname_off[in->qname_id[i] + 1] - in->qname_off[in->qname_id[i]]);
        char buf[32]; return (size_t)snprintf(buf, sizeof(buf), "read%u", in->qname_id[i]);
    };
    for (int64_t i = 0; i < n; ++i) {
        roff[i] = o;
        o += 4 + 32 + name_len(i) + 1 + 4 * (size_t)in->n_cigar[i] + (size_t)(in->l_seq[i] + 1) / 2 + (size_t)in->l_seq[i] + (in->sa_len[i] ? 4 + in->sa_len[i] : 0);
    }
    roff[n] = o;
    raw.resize(o);
    const int threads = std::max(1, n_threads);
    parallel_for((size_t)n, threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            uint8_t* r = raw.data() + roff[i];
            const uint32_t bs = (uint32_t)(roff[i + 1] - roff[i] - 4);
            memcpy(r, &bs, 4); r += 4;
            const uint32_t nc = in->n_cigar[i]; const int32_t ls = in->l_seq[i];
            const uint32_t* cg = in->cigar + in->cigar_off[i];
            int64_t rlen = 0;
            for (uint32_t k = 0; k < nc; ++k) { const uint32_t op = cg[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += cg[k] >> 4; }
            char nb[32]; const char* nm; size_t nl;
            if (in->qnames) { nm = in->qnames + in->qname_off[in->qname_id[i]]; nl = name_len(i); }
            else { nl = (size_t)snprintf(nb, sizeof(nb), "read%u", in->qname_id[i]); nm = nb; }
            const int32_t pos = in->pos[i];
            const int64_t b0 = pos < 0 ? 0 : pos;
            const uint16_t bin = (uint16_t)reg2bin(b0, b0 + (rlen ? rlen : 1));
            int32_t f32[2] = {in->tid[i], pos}; memcpy(r, f32, 8);
            r[8] = (uint8_t)(nl + 1); r[9] = in->mapq[i]; memcpy(r + 10, &bin, 2);
            const uint16_t nc16 = (uint16_t)nc; memcpy(r + 12, &nc16, 2); memcpy(r + 14, &in->flag[i], 2);
            memcpy(r + 16, &ls, 4);
            const int32_t m1 = -1, z = 0; memcpy(r + 20, &m1, 4); memcpy(r + 24, &m1, 4); memcpy(r + 28, &z, 4);
            uint8_t* q = r + 32;
            memcpy(q, nm, nl); q[nl] = 0; q += nl + 1;
            memcpy(q, cg, 4 * (size_t)nc); q += 4 * (size_t)nc;
            memcpy(q, in->seq + in->seq_off[i], (size_t)(ls + 1) / 2); q += (size_t)(ls + 1) / 2;
            memset(q, 0xff, (size_t)ls); q += ls;
            if (in->sa_len[i]) { q[0] = 'S'; q[1] = 'A'; q[2] = 'Z'; memcpy(q + 3, in->sa + in->sa_off[i], in->sa_len[i]); q[3 + in->sa_len[i]] = 0; }
        }
    });
    // BGZF: fixed 0xff00-byte payloads, compressed in parallel
    const size_t CH = 0xff00;
    const size_t nblk = (raw.size() + CH - 1) / CH;
    std::vector<std::vector<uint8_t>> comp(nblk);
    std::atomic<int> bad{0};
    parallel_for(nblk, threads, [&](size_t lo, size_t hi) {
        for (size_t b = lo; b < hi; ++b) {
            const size_t off = b * CH, len = std::min(CH, raw.size() - off);
            std::vector<uint8_t>& out = comp[b];
            out.resize(18 + compressBound((uLong)len) + 8 + 64);
            z_stream zs; memset(&zs, 0, sizeof(zs));
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
            zs.next_in = raw.data() + off; zs.avail_in = (uInt)len; zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 26);
            const int rc = deflate(&zs, Z_FINISH);
            const size_t clen = zs.total_out;
            deflateEnd(&zs);
            if (rc != Z_STREAM_END || clen + 26 > 65536) { bad = 1; return; }
            const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
            memcpy(out.data(), hdr, 16);
            const uint16_t bsz = (uint16_t)(clen + 25); memcpy(out.data() + 16, &bsz, 2);
            const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), raw.data() + off, (uInt)len), isz = (uint32_t)len;
            memcpy(out.data() + 18 + clen, &crc, 4); memcpy(out.data() + 22 + clen, &isz, 4);
            out.resize(26 + clen);
        }
    });
    if (bad) return -2;
    FILE* fh = fopen(path, "wb");
    if (!fh) return -1;
    for (auto& c : comp) fwrite(c.data(), 1, c.size(), fh);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, fh);
    fclose(fh);
    return 0;
}

}  // extern "C"

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
#include <algorithm>
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

struct Handle {
    std::vector<uint8_t> data;      // inflated stream
    std::vector<Rec> recs;
    std::vector<std::string> contigs; std::vector<int64_t> contig_len;
    std::vector<uint32_t> qid;
    std::vector<std::string_view> qnames;
    std::string sort_order;
    bamio_info info;
    int threads = 1;
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
    FILE* fh = fopen(path, "rb");
    if (!fh) return fail("cannot open file");
    fseek(fh, 0, SEEK_END); const long fsz = ftell(fh); fseek(fh, 0, SEEK_SET);
    std::vector<uint8_t> file((size_t)fsz);
    if (fsz && fread(file.data(), 1, (size_t)fsz, fh) != (size_t)fsz) { fclose(fh); return fail("short read"); }
    fclose(fh);
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
    Handle* h = new Handle();
    h->threads = std::max(1, n_threads);
    h->data.resize(uoff);
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
    std::vector<uint8_t>().swap(file);
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
    return 0;
}

void bamio_close(void* hh) { delete (Handle*)hh; }

}  // extern "C"

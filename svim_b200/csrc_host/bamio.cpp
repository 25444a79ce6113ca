// Multi-threaded BAM -> flattened record buffer (svim_aln_soa) decoder.
//
// once the kernels run at TB/s the end-to-end rate is set by host BAM
// decompression.  The reference iterates pysam/htslib records one by one (SVIM_COLLECT.py:133).
//
// Streaming design (no inflated copy of the file is ever held):
//   * the file is mmap'ed; one sequential pass over the BGZF headers indexes the blocks;
//   * the blocks are cut into units of ~1 MiB of inflated data; worker threads claim units in order, inflate a unit into a
//     thread-local buffer that stays cache-warm (zlib raw inflate per block), and then take their turn in a CHAIN that is
//     ordered by unit: the chain turn prepends the bytes the previous unit left over (a record cut by the unit boundary),
//     hops over the record block_size fields, assigns every record its row number and its offsets into the CIGAR / SEQ /
//     SA blobs (running totals travel along the chain), and hands the tail to the next unit.  The chain turn touches a few
//     bytes per record; everything heavy — inflate before it, the copy of CIGAR / SEQ / SA bytes into the caller's blobs
//     after it — runs in parallel on the unit's own warm buffer.
//   * QUAL, names (kept once, for the id table) and other aux fields never leave the unit buffer.
// Host code only — nothing here runs on the GPU path's timed kernels.
//
// File format: SAMv1 §4.1 (BGZF) and §4.2 (BAM).
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <sched.h>
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
    int64_t n_records, cigar_words, seq_bytes, sa_bytes, n_qnames, names_bytes;     // exact, valid after bamio_decode
    int64_t cigar_bound_words, seq_bound_bytes;                                     // upper bounds, valid after bamio_open
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

// Large first-touch buffers are page-fault bound with 4 KiB pages; ask for transparent huge pages.
static void advise_huge(void* p, size_t bytes) {
#ifdef MADV_HUGEPAGE
    const uintptr_t a = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
    const uintptr_t e = ((uintptr_t)p + bytes) & ~(uintptr_t)((2u << 20) - 1);
    if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
#endif
}

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

// one record's row, produced in the chain turn
struct Row {
    int32_t tid, pos; uint16_t flag; uint8_t mapq, l_rn; uint32_t n_cigar; int32_t l_seq; uint32_t sa_len;
    uint64_t cigar_off, seq_off, sa_off;
    uint32_t src_off, sa_src;          // record start / SA payload inside the unit's contiguous view
    uint32_t cig_src;                  // CIGAR words relative to the record start (core CIGAR, or the CG:B,I payload)
    uint32_t n_cigar_core;             // operations stored in the record core (2 for a CG-tagged record): SEQ follows them
    uint32_t name_off;                 // into the unit's name arena
    uint32_t c8_len = 0;               // bytes of the record's 8-bit packed CIGAR stream (multiple of 16), when asked for
};

struct Unit {
    size_t b0 = 0, b1 = 0, ulen = 0;   // blocks [b0, b1), inflated bytes
    std::vector<Row> rows;
    std::string names;                 // NUL-terminated read names of the unit's records, in order
    std::string sa;                    // SA tag payloads of the unit's records, concatenated (small: copied out at the end)
    uint64_t row_base = 0, sa_base = 0;
    std::vector<uint8_t> c8;           // svim_aln_soa.cigar8 streams of the unit's records, back to back (bamio_set_pack(h, 8))
};

struct Handle {
    int fd = -1; const uint8_t* file = nullptr; size_t fsz = 0;
    std::vector<Block> blocks;
    std::vector<Unit> units;
    std::vector<std::string> contigs; std::vector<int64_t> contig_len;
    std::vector<uint32_t> qid;
    std::vector<std::string_view> qnames;
    std::string sort_order = "unknown";
    bamio_info info;
    int threads = 1;
    bool header_done = false;
    size_t header_bytes = 0;          // BAM header length in the inflated stream = offset of the first alignment record
    int pack = 0;                     // 8: also emit the 8-bit packed CIGAR stream while the unit is warm (bamio_set_pack)
    ~Handle() { if (file && fsz) munmap((void*)file, fsz); if (fd >= 0) close(fd); }
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

// Walk the typed aux fields of a record: SA:Z payload -> (offset relative to record start, length); CG:B,I payload (the real
// CIGAR of a record with more than 65535 operations, SAMv1 §4.2.2) -> (offset of its first word, number of words).
struct AuxHits { uint32_t sa_off = 0, sa_len = 0, cg_off = 0, cg_n = 0; };
void scan_aux(const uint8_t* rec, size_t aux_begin, size_t rec_len, AuxHits& hit) {
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
                if (t0 == 'S' && t1 == 'A' && ty == 'Z' && hit.sa_len == 0 && hit.sa_off == 0) { hit.sa_off = (uint32_t)o; hit.sa_len = (uint32_t)(e - o); }
                o = e + 1;
                continue;
            }
            case 'B': {
                if (o + 5 > rec_len) return;
                const uint8_t sub = rec[o]; const uint32_t cnt = rd32(rec + o + 1);
                size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                if (t0 == 'C' && t1 == 'G' && sub == 'I' && hit.cg_off == 0 && o + 5 + (size_t)cnt * 4 <= rec_len) { hit.cg_off = (uint32_t)(o + 5); hit.cg_n = cnt; }
                o += 5 + (size_t)cnt * es;
                continue;
            }
            default: return;
        }
        o += sz;
    }
}

// One BGZF payload (raw DEFLATE, exact output size known from ISIZE).  The z_stream is per thread and reset per block:
// inflateInit2 / inflateEnd would malloc and free the decoder state for each of the ~10^5 blocks.
struct ZStream {
    z_stream zs; bool ready = false;
    ~ZStream() { if (ready) inflateEnd(&zs); }
};

bool inflate_block(const uint8_t* src, size_t clen, uint8_t* dst, size_t ulen) {
    static thread_local ZStream z;
    if (!z.ready) {
        memset(&z.zs, 0, sizeof(z.zs));
        if (inflateInit2(&z.zs, -15) != Z_OK) return false;
        z.ready = true;
    } else if (inflateReset(&z.zs) != Z_OK) return false;
    z.zs.next_in = const_cast<uint8_t*>(src); z.zs.avail_in = (uInt)clen; z.zs.next_out = dst; z.zs.avail_out = (uInt)ulen;
    const int rc = inflate(&z.zs, Z_FINISH);
    return rc == Z_STREAM_END && z.zs.avail_out == 0;
}

// BAM header (magic, text, references) at the front of `d`: returns bytes consumed, 0 = not all there yet, -1 = not BAM
int64_t parse_header(Handle* h, const uint8_t* d, size_t n) {
    if (n < 12) return 0;
    if (memcmp(d, "BAM\1", 4) != 0) return -1;
    const uint32_t l_text = rd32(d + 4);
    size_t p = 8 + (size_t)l_text;
    if (p + 4 > n) return 0;
    const uint32_t n_ref = rd32(d + p); p += 4;
    std::vector<std::string> names; std::vector<int64_t> lens;
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (p + 4 > n) return 0;
        const uint32_t ln = rd32(d + p); p += 4;
        if (p + (size_t)ln + 4 > n) return 0;
        names.emplace_back((const char*)d + p, ln ? ln - 1 : 0); p += ln;
        lens.push_back((int32_t)rd32(d + p)); p += 4;
    }
    std::string text((const char*)d + 8, strnlen((const char*)d + 8, l_text));
    size_t q = text.find("@HD");
    if (q != std::string::npos) {
        size_t e = text.find('\n', q), s = text.find("SO:", q);
        if (s != std::string::npos && (e == std::string::npos || s < e)) {
            size_t t = s + 3, u = t;
            while (u < text.size() && text[u] != '\t' && text[u] != '\n') ++u;
            h->sort_order = text.substr(t, u - t);
        }
    }
    h->contigs.swap(names); h->contig_len.swap(lens);
    h->header_done = true; h->header_bytes = p;
    return (int64_t)p;
}

}  // namespace

extern "C" {

// Maps the file, indexes the BGZF blocks and reads the BAM header.  info: bounds for the caller-allocated blobs.
void* bamio_open(const char* path, int n_threads, bamio_info* info, char* err, int errcap) {
    auto fail = [&](const char* m) -> void* { if (err && errcap > 0) snprintf(err, errcap, "%s", m); return nullptr; };
    Trace tr;
    std::unique_ptr<Handle> h(new Handle());
    h->threads = std::max(1, n_threads);
    h->fd = open(path, O_RDONLY);
    if (h->fd < 0) return fail("cannot open file");
    struct stat sb;
    if (fstat(h->fd, &sb) != 0) return fail("cannot stat file");
    h->fsz = (size_t)sb.st_size;
    if (h->fsz) {
        void* m = mmap(nullptr, h->fsz, PROT_READ, MAP_PRIVATE, h->fd, 0);
        if (m == MAP_FAILED) { h->fsz = 0; return fail("cannot map file"); }
        h->file = (const uint8_t*)m;
        madvise(m, h->fsz, MADV_WILLNEED);
    }
    // ---- BGZF block index ----------------------------------------------------------------------------
    size_t o = 0, uoff = 0;
    while (o + 18 <= h->fsz) {
        const uint8_t* p = h->file + o;
        if (!(p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 4))) return fail("not a BGZF file");
        const uint16_t xlen = rd16(p + 10);
        size_t x = 12, xe = 12 + xlen; int bsize = -1;
        while (x + 4 <= xe) {
            const uint16_t slen = rd16(p + x + 2);
            if (p[x] == 66 && p[x + 1] == 67 && slen == 2) bsize = rd16(p + x + 4);
            x += 4 + slen;
        }
        if (bsize < 0 || o + (size_t)bsize + 1 > h->fsz || (size_t)bsize + 1 < (size_t)xlen + 20) return fail("bad BGZF block");
        const size_t total = (size_t)bsize + 1;
        const uint32_t isize = rd32(p + total - 4);
        if (isize) h->blocks.push_back({o + 12 + xlen, total - xlen - 20, uoff, isize});
        uoff += isize; o += total;
    }
    tr.mark("map + block index");
    // units of ~1 MiB inflated
    const size_t UNIT = (size_t)1 << 20;
    for (size_t b = 0; b < h->blocks.size();) {
        Unit u; u.b0 = b;
        while (b < h->blocks.size() && (u.ulen == 0 || u.ulen + h->blocks[b].ulen <= UNIT)) { u.ulen += h->blocks[b].ulen; ++b; }
        u.b1 = b;
        h->units.push_back(std::move(u));
    }
    // header: inflate leading blocks until it parses (contig names are needed before the caller can allocate anything)
    {
        std::vector<uint8_t> head;
        for (size_t b = 0; b < h->blocks.size() && !h->header_done; ++b) {
            const size_t at = head.size();
            head.resize(at + h->blocks[b].ulen);
            if (!inflate_block(h->file + h->blocks[b].coff, h->blocks[b].clen, head.data() + at, h->blocks[b].ulen)) return fail("inflate failed");
            const int64_t used = parse_header(h.get(), head.data(), head.size());
            if (used < 0) return fail("not a BAM stream");
        }
        if (!h->header_done) return fail("truncated BAM header");
        h->header_done = false;      // the chain parses it again to find where the records start
    }
    bamio_info& inf = h->info;
    memset(&inf, 0, sizeof(inf));
    inf.n_contigs = (int32_t)h->contigs.size();
    inf.sorted_coordinate = h->sort_order == "coordinate";
    const int64_t rec_bound = (int64_t)(uoff / 36) + 1;                 // a record is at least 4 + 32 bytes
    inf.cigar_bound_words = (int64_t)(uoff / 4) + 3 * rec_bound + 4;    // + padding to multiples of 4 words
    inf.seq_bound_bytes = (int64_t)(uoff / 3) + rec_bound + 16;         // packed SEQ is at most a third of SEQ + QUAL
    *info = inf;
    tr.mark("header + units");
    return h.release();
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

// Streams the whole file: fills the caller's CIGAR / SEQ blobs (allocated to the bounds of bamio_open) and keeps the
// per-record rows and the (small) SA payloads inside the handle; info gets the exact counts.  Returns 0, or a negative code with text in err.
int bamio_decode(void* hh, uint32_t* cigar, uint8_t* seq, bamio_info* info, char* err, int errcap) {
    Handle* h = (Handle*)hh;
    Trace tr;
    if (tr.on) {
        cpu_set_t cs; CPU_ZERO(&cs);
        const int aff = sched_getaffinity(0, sizeof(cs), &cs) == 0 ? CPU_COUNT(&cs) : -1;
        char quota[64] = "?"; if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) { if (!fgets(quota, sizeof(quota), f)) quota[0] = 0; fclose(f); quota[strcspn(quota, "\n")] = 0; }
        char thp[128] = "?"; if (FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r")) { if (!fgets(thp, sizeof(thp), f)) thp[0] = 0; fclose(f); thp[strcspn(thp, "\n")] = 0; }
        fprintf(stderr, "[bamio] threads %d, hardware %u, affinity %d, cgroup cpu.max '%s', THP '%s', %zu blocks in %zu units\n", h->threads,
                std::thread::hardware_concurrency(), aff, quota, thp, h->blocks.size(), h->units.size());
    }
    advise_huge(cigar, (size_t)h->info.cigar_bound_words * 4); advise_huge(seq, (size_t)h->info.seq_bound_bytes);
    const size_t n_units = h->units.size();
    std::atomic<size_t> next_unit{0}, chain_turn{0};
    std::atomic<int> failed{0};
    const char* fail_msg = "decode failed";
    // chain state (only touched by the thread whose turn it is)
    std::vector<uint8_t> carry;
    uint64_t n_rows = 0, cig_words = 0, seq_bytes = 0, sa_bytes = 0;
    const size_t HEADROOM = (size_t)256 << 10;

    std::atomic<int64_t> ns_inflate{0}, ns_wait{0}, ns_chain{0}, ns_fill{0};
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point a) { return (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - a).count(); };
    auto worker = [&]() {
        int64_t t_inf = 0, t_wait = 0, t_chain = 0, t_fill = 0;
        std::vector<uint8_t> buf;            // [headroom | unit data]
        std::vector<uint8_t> joined;         // only when the carry does not fit the headroom
        for (;;) {
            const size_t u = next_unit.fetch_add(1);
            if (u >= n_units) { ns_inflate += t_inf; ns_wait += t_wait; ns_chain += t_chain; ns_fill += t_fill; return; }
            Unit& un = h->units[u];
            auto t0 = now();
            if (buf.size() < HEADROOM + un.ulen) buf.resize(HEADROOM + un.ulen);
            bool ok = !failed.load(std::memory_order_relaxed);
            if (ok) {
                size_t at = HEADROOM;
                for (size_t b = un.b0; b < un.b1 && ok; ++b) {
                    ok = inflate_block(h->file + h->blocks[b].coff, h->blocks[b].clen, buf.data() + at, h->blocks[b].ulen);
                    at += h->blocks[b].ulen;
                }
                if (!ok) { fail_msg = "inflate failed"; failed = 1; }
            }
            t_inf += since(t0); t0 = now();
            // ---- chain turn ----
            for (unsigned spins = 0; chain_turn.load(std::memory_order_acquire) != u; ++spins) if (spins > 64) std::this_thread::yield();
            t_wait += since(t0); t0 = now();
            const uint8_t* d = nullptr; size_t n = 0;
            if (!failed.load()) {
                if (carry.size() <= HEADROOM) {
                    d = buf.data() + HEADROOM - carry.size();
                    if (!carry.empty()) memcpy(buf.data() + HEADROOM - carry.size(), carry.data(), carry.size());
                    n = carry.size() + un.ulen;
                } else {
                    joined.resize(carry.size() + un.ulen);
                    memcpy(joined.data(), carry.data(), carry.size()); memcpy(joined.data() + carry.size(), buf.data() + HEADROOM, un.ulen);
                    d = joined.data(); n = joined.size();
                }
                size_t p = 0;
                if (!h->header_done) {
                    const int64_t used = parse_header(h, d, n);
                    if (used < 0) { fail_msg = "not a BAM stream"; failed = 1; }
                    else p = (size_t)used;               // 0: header continues in the next unit, everything is carried
                }
                un.row_base = n_rows; un.sa_base = sa_bytes;
                if (h->header_done && !failed.load()) {
                    while (p + 4 <= n) {
                        const uint32_t bs = rd32(d + p);
                        if (bs < 32) { fail_msg = "corrupt record"; failed = 1; break; }
                        if (p + 4 + (size_t)bs > n) break;
                        const uint8_t* r = d + p + 4;
                        Row w;
                        w.tid = (int32_t)rd32(r); w.pos = (int32_t)rd32(r + 4); w.l_rn = r[8]; w.mapq = r[9];
                        w.n_cigar = rd16(r + 12); w.flag = rd16(r + 14); w.l_seq = (int32_t)rd32(r + 16);
                        const size_t aux = 32 + (size_t)w.l_rn + 4 * (size_t)w.n_cigar + (size_t)(w.l_seq + 1) / 2 + (size_t)w.l_seq;
                        if (w.l_seq < 0 || aux > bs) { fail_msg = "corrupt record"; failed = 1; break; }
                        AuxHits hit;
                        if (aux < bs) scan_aux(r, aux, bs, hit);
                        const uint32_t so = hit.sa_off, sl = hit.sa_len;
                        w.sa_len = sl; w.sa_src = 0; w.src_off = (uint32_t)(p + 4);
                        if (sl) un.sa.append((const char*)r + so, sl);
                        w.n_cigar_core = w.n_cigar; w.cig_src = 32u + w.l_rn;
                        if (hit.cg_n && w.n_cigar > 0 && w.tid >= 0 && w.pos >= 0) {          // htslib sam.c bam_tag2cigar
                            const uint32_t c0 = rd32(r + 32 + w.l_rn);
                            if ((c0 & 15u) == 4u && (int64_t)(c0 >> 4) == (int64_t)w.l_seq) { w.n_cigar = hit.cg_n; w.cig_src = hit.cg_off; }
                        }
                        w.cigar_off = cig_words; w.seq_off = seq_bytes; w.sa_off = sa_bytes;
                        cig_words += (w.n_cigar + 3u) & ~3u; seq_bytes += (uint64_t)(w.l_seq + 1) / 2; sa_bytes += sl;
                        w.name_off = (uint32_t)un.names.size();
                        un.names.append((const char*)r + 32, w.l_rn ? w.l_rn - 1 : 0); un.names.push_back('\0');
                        un.rows.push_back(w);
                        p += 4 + (size_t)bs;
                    }
                    n_rows += un.rows.size();
                }
                carry.assign(d + p, d + n);              // the tail of a record cut by the unit boundary (or of the header)
            }
            chain_turn.store(u + 1, std::memory_order_release);
            t_chain += since(t0); t0 = now();
            if (failed.load()) continue;
            // ---- fill: this unit's records from the warm buffer into the caller's blobs ----
            for (const Row& w : un.rows) {
                const uint8_t* r = d + w.src_off;
                uint32_t* dst = cigar + w.cigar_off;
                memcpy(dst, r + w.cig_src, 4 * (size_t)w.n_cigar);
                for (uint32_t k = w.n_cigar; k < ((w.n_cigar + 3u) & ~3u); ++k) dst[k] = 0;
                memcpy(seq + w.seq_off, r + 32 + w.l_rn + 4 * (size_t)w.n_cigar_core, (size_t)(w.l_seq + 1) / 2);
            }
            if (h->pack == 8) {
                // the 8-bit packed stream of include/svimgpu.h (bamio_pack_cigar8 below is the stand-alone version), from the same warm bytes
                size_t len8 = 0;
                un.c8.resize(std::max<size_t>(un.c8.size(), un.ulen / 3 + 64));
                for (Row& w : un.rows) {
                    const size_t worst = len8 + (size_t)w.n_cigar * 7 + 16;              // 28-bit length: six extension bytes + the operation byte
                    if (un.c8.size() < worst) un.c8.resize(std::max(worst, un.c8.size() * 2));
                    uint8_t* o = un.c8.data() + len8;
                    const uint8_t* cg = d + w.src_off + w.cig_src;
                    size_t q = 0;
                    for (uint32_t k = 0; k < w.n_cigar; ++k) {
                        const uint32_t c = rd32(cg + 4 * (size_t)k), len = c >> 4;
                        if (len >= 16u) {
                            int e = 0; for (uint32_t l = len; l >= 16u; l >>= 4) ++e;
                            for (; e > 0; --e) o[q++] = (uint8_t)((((len >> (4 * e)) & 15u) << 4) | 15u);
                        }
                        o[q++] = (uint8_t)(((len & 15u) << 4) | (c & 15u));
                    }
                    while (q & 15u) o[q++] = 0x0F;
                    w.c8_len = (uint32_t)q; len8 += q;
                }
                un.c8.resize(len8);
            }
            t_fill += since(t0);
        }
    };
    {
        std::vector<std::thread> th;
        const int nt = (int)std::min<size_t>((size_t)h->threads, std::max<size_t>(1, n_units));
        for (int t = 1; t < nt; ++t) th.emplace_back(worker);
        worker();
        for (auto& x : th) x.join();
    }
    if (!failed && !carry.empty()) { fail_msg = h->header_done ? "truncated record" : "truncated BAM header"; failed = 1; }
    if (failed) { if (err && errcap > 0) snprintf(err, errcap, "%s", fail_msg); return -1; }
    tr.mark("inflate + chain + fill");
    if (tr.on) fprintf(stderr, "[bamio]   thread-time sums: inflate %.0f ms, wait for chain turn %.0f ms, chain turn %.0f ms, fill %.0f ms\n",
                       ns_inflate.load() / 1e6, ns_wait.load() / 1e6, ns_chain.load() / 1e6, ns_fill.load() / 1e6);
    // ---- read-name ids -------------------------------------------------------------------------------------------------
    const size_t nr = (size_t)n_rows;
    h->qid.resize(nr);
    {
        std::unordered_map<std::string_view, uint32_t> ids;
        ids.reserve(nr * 2);
        size_t i = 0;
        for (const Unit& un : h->units)
            for (const Row& w : un.rows) {
                std::string_view nm(un.names.data() + w.name_off, w.l_rn ? w.l_rn - 1 : 0);
                auto it = ids.find(nm);
                if (it == ids.end()) { it = ids.emplace(nm, (uint32_t)h->qnames.size()).first; h->qnames.push_back(nm); }
                h->qid[i++] = it->second;
            }
    }
    tr.mark("read-name ids (serial hash)");
    bamio_info& inf = h->info;
    inf.n_records = (int64_t)nr; inf.cigar_words = (int64_t)cig_words; inf.seq_bytes = (int64_t)seq_bytes; inf.sa_bytes = (int64_t)sa_bytes;
    inf.n_qnames = (int64_t)h->qnames.size(); inf.names_bytes = 0;
    for (auto& q : h->qnames) inf.names_bytes += (int64_t)q.size() + 1;
    *info = inf;
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

// per-record rows into caller arrays of n_records entries and the SA blob (sa_bytes); out->cigar / seq are ignored
// (bamio_decode filled them)
int bamio_fill(void* hh, bamio_out* out) {
    Handle* h = (Handle*)hh;
    Trace tr;
    parallel_for(h->units.size(), h->threads, [&](size_t lo, size_t hi) {
        for (size_t u = lo; u < hi; ++u) {
            const Unit& un = h->units[u];
            size_t i = (size_t)un.row_base;
            for (const Row& w : un.rows) {
                out->tid[i] = w.tid; out->pos[i] = w.pos; out->mapq[i] = w.mapq; out->flag[i] = w.flag;
                out->n_cigar[i] = w.n_cigar; out->l_seq[i] = w.l_seq; out->sa_len[i] = w.sa_len; out->qname_id[i] = h->qid[i];
                out->cigar_off[i] = w.cigar_off; out->seq_off[i] = w.seq_off; out->sa_off[i] = w.sa_off;
                ++i;
            }
            if (!un.sa.empty()) memcpy(out->sa + un.sa_base, un.sa.data(), un.sa.size());
        }
    });
    tr.mark("row arrays");
    return 0;
}

// 8-bit packed CIGAR stream emitted by the decode itself.  bamio_set_pack(h, 8) before bamio_decode; afterwards
// bamio_cigar8_bytes = total stream bytes, bamio_cigar8_fill writes off8[n + 1] (bytes) and the stream.
int bamio_set_pack(void* hh, int bits) { Handle* h = (Handle*)hh; if (bits != 0 && bits != 8) return -1; h->pack = bits; return 0; }

int64_t bamio_cigar8_bytes(void* hh) {
    Handle* h = (Handle*)hh;
    int64_t t = 0;
    for (const Unit& un : h->units) t += (int64_t)un.c8.size();
    return t;
}

int bamio_cigar8_fill(void* hh, uint64_t* off8, uint8_t* out8) {
    Handle* h = (Handle*)hh;
    if (h->pack != 8 || !off8) return -1;
    std::vector<uint64_t> ubase(h->units.size() + 1, 0);
    for (size_t u = 0; u < h->units.size(); ++u) ubase[u + 1] = ubase[u] + h->units[u].c8.size();
    parallel_for(h->units.size(), h->threads, [&](size_t lo, size_t hi) {
        for (size_t u = lo; u < hi; ++u) {
            const Unit& un = h->units[u];
            size_t i = (size_t)un.row_base; uint64_t at = ubase[u];
            for (const Row& w : un.rows) { off8[i++] = at; at += w.c8_len; }
            if (out8 && !un.c8.empty()) memcpy(out8 + ubase[u], un.c8.data(), un.c8.size());
        }
    });
    uint64_t n = 0;
    for (const Unit& un : h->units) n += un.rows.size();
    off8[n] = ubase.back();
    return 0;
}

// Layout for the experimental on-GPU decoder (csrc_next/bamgpu.cu): the mapped file, the BGZF block table
// {payload offset, inflated offset, payload bytes, inflated bytes} and where the alignment records start in the inflated stream.
struct bamio_block { uint64_t coff, uoff; uint32_t clen, ulen; };
int bamio_layout(void* hh, const uint8_t** file, int64_t* file_bytes, int64_t* n_blocks, int64_t* first_record) {
    Handle* h = (Handle*)hh;
    *file = h->file; *file_bytes = (int64_t)h->fsz; *n_blocks = (int64_t)h->blocks.size(); *first_record = (int64_t)h->header_bytes;
    return 0;
}
int bamio_blocks(void* hh, bamio_block* out) {
    Handle* h = (Handle*)hh;
    for (size_t i = 0; i < h->blocks.size(); ++i) out[i] = {h->blocks[i].coff, h->blocks[i].uoff, (uint32_t)h->blocks[i].clen, (uint32_t)h->blocks[i].ulen};
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
        const size_t nc_all = in->n_cigar[i];
        const size_t cig_bytes = nc_all <= 65535 ? 4 * nc_all : 8 + 8 + 4 * nc_all;      // placeholder kSmN in the core + CG:B,I tag
        o += 4 + 32 + name_len(i) + 1 + cig_bytes + (size_t)(in->l_seq[i] + 1) / 2 + (size_t)in->l_seq[i] + (in->sa_len[i] ? 4 + in->sa_len[i] : 0);
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
            const bool long_cigar = nc > 65535;            // SAMv1 §4.2.2: the real CIGAR goes to CG:B,I, the core keeps <l_seq>S<rlen>N
            const uint16_t nc16 = long_cigar ? (uint16_t)2 : (uint16_t)nc; memcpy(r + 12, &nc16, 2); memcpy(r + 14, &in->flag[i], 2);
            memcpy(r + 16, &ls, 4);
            const int32_t m1 = -1, z = 0; memcpy(r + 20, &m1, 4); memcpy(r + 24, &m1, 4); memcpy(r + 28, &z, 4);
            uint8_t* q = r + 32;
            memcpy(q, nm, nl); q[nl] = 0; q += nl + 1;
            if (long_cigar) {
                const uint32_t core[2] = {((uint32_t)ls << 4) | 4u, ((uint32_t)rlen << 4) | 3u};
                memcpy(q, core, 8); q += 8;
            } else { memcpy(q, cg, 4 * (size_t)nc); q += 4 * (size_t)nc; }
            memcpy(q, in->seq + in->seq_off[i], (size_t)(ls + 1) / 2); q += (size_t)(ls + 1) / 2;
            memset(q, 0xff, (size_t)ls); q += ls;
            if (in->sa_len[i]) { q[0] = 'S'; q[1] = 'A'; q[2] = 'Z'; memcpy(q + 3, in->sa + in->sa_off[i], in->sa_len[i]); q[3 + in->sa_len[i]] = 0; q += 4 + in->sa_len[i]; }
            if (long_cigar) { q[0] = 'C'; q[1] = 'G'; q[2] = 'B'; q[3] = 'I'; memcpy(q + 4, &nc, 4); memcpy(q + 8, cg, 4 * (size_t)nc); }
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


// ---- 16-bit packed CIGAR (svim_aln_soa.cigar16, include/svimgpu.h) ------------------------------------------------------------
// The record buffer crosses PCIe as CIGAR words; BAM's uint32 (len << 4 | op) is mostly zeros for long reads (mean run ~5 bases).
// Packed form: uint16 (len << 4 | op) for len < 4096; a longer length is split: words with op nibble 0xF carry 12 more
// significant bits each and precede the op word (decoder: acc = acc << 12 | hi12 on 0xF words; len = acc << 12 | hi12 on the
// op word).  Every record's stream is padded to a multiple of 8 words with 0x000F (no-ops), so records stay 16-byte aligned.
// Pass 1 (out16 == nullptr): off16[i] for every record and off16[n] = total words.  Pass 2: fills out16.
int bamio_pack_cigar16(int64_t n, const uint32_t* n_cigar, const uint64_t* cigar_off, const uint32_t* cigar, uint64_t* off16, uint16_t* out16, int n_threads) {
    if (n < 0 || (n > 0 && (!n_cigar || !cigar_off || !cigar || !off16))) return -1;
    if (n_threads < 1) n_threads = 1;
    auto words_of = [&](int64_t i) {
        uint64_t w = 0;
        const uint32_t* c = cigar + cigar_off[i];
        for (uint32_t k = 0; k < n_cigar[i]; ++k) { const uint32_t len = c[k] >> 4; w += 1 + (len >= (1u << 12)) + (len >= (1u << 24)); }
        return (w + 7) & ~7ull;
    };
    if (!out16) {
        std::vector<uint64_t> cnt((size_t)n);
        parallel_for((size_t)n, n_threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) cnt[i] = words_of((int64_t)i); });
        uint64_t at = 0;
        for (int64_t i = 0; i < n; ++i) { off16[i] = at; at += cnt[(size_t)i]; }
        off16[n] = at;
        return 0;
    }
    parallel_for((size_t)n, n_threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) {
        const uint32_t* c = cigar + cigar_off[i];
        uint16_t* o = out16 + off16[i];
        uint64_t w = 0;
        for (uint32_t k = 0; k < n_cigar[i]; ++k) {
            const uint32_t len = c[k] >> 4, op = c[k] & 15u;
            if (len >= (1u << 24)) o[w++] = (uint16_t)(((len >> 24) << 4) | 15u);
            if (len >= (1u << 12)) o[w++] = (uint16_t)((((len >> 12) & 0xFFFu) << 4) | 15u);
            o[w++] = (uint16_t)(((len & 0xFFFu) << 4) | op);
        }
        const uint64_t end = off16[i + 1] - off16[i];
        while (w < end) o[w++] = 0x000F;
    } });
    return 0;
}

// ---- 8-bit packed CIGAR (svim_aln_soa.cigar8, include/svimgpu.h) -------------------------------------------------------------
// One byte per operation of fewer than 16 bases (len << 4 | op); longer ones are preceded by extension bytes (op nibble 0xF,
// 4 more significant length bits each, most significant first, no leading zero extension).  Every record's stream is padded to a
// multiple of 16 bytes with 0x0F.  Pass 1 (out8 == nullptr): off8[i] and off8[n] = total bytes.  Pass 2: fills out8.
int bamio_pack_cigar8(int64_t n, const uint32_t* n_cigar, const uint64_t* cigar_off, const uint32_t* cigar, uint64_t* off8, uint8_t* out8, int n_threads) {
    if (n < 0 || (n > 0 && (!n_cigar || !cigar_off || !cigar || !off8))) return -1;
    if (n_threads < 1) n_threads = 1;
    auto ext_of = [](uint32_t len) { int e = 0; while (len >= 16u) { len >>= 4; ++e; } return e; };     // extension bytes of a length
    auto bytes_of = [&](int64_t i) {
        uint64_t w = 0;
        const uint32_t* c = cigar + cigar_off[i];
        for (uint32_t k = 0; k < n_cigar[i]; ++k) w += 1 + (uint64_t)ext_of(c[k] >> 4);
        return (w + 15) & ~15ull;
    };
    if (!out8) {
        std::vector<uint64_t> cnt((size_t)n);
        parallel_for((size_t)n, n_threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) cnt[i] = bytes_of((int64_t)i); });
        uint64_t at = 0;
        for (int64_t i = 0; i < n; ++i) { off8[i] = at; at += cnt[(size_t)i]; }
        off8[n] = at;
        return 0;
    }
    parallel_for((size_t)n, n_threads, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) {
        const uint32_t* c = cigar + cigar_off[i];
        uint8_t* o = out8 + off8[i];
        uint64_t w = 0;
        for (uint32_t k = 0; k < n_cigar[i]; ++k) {
            const uint32_t len = c[k] >> 4, op = c[k] & 15u;
            for (int e = ext_of(len); e > 0; --e) o[w++] = (uint8_t)((((len >> (4 * e)) & 15u) << 4) | 15u);
            o[w++] = (uint8_t)(((len & 15u) << 4) | op);
        }
        const uint64_t end = off8[i + 1] - off8[i];
        while (w < end) o[w++] = 0x0F;
    } });
    return 0;
}

}  // extern "C"

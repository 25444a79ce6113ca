// Synthetic long-read alignment generator -> flattened record buffer (svim_aln_soa).
//
// Bench / test infrastructure for the BASELINE.json configs: the
// reference ships no data and this image has no aligner, so coordinate-sorted
// alignments are synthesised directly in the SoA layout the COLLECT kernels read
// (svim_b200/records.py).  Deterministic in (seed, read index) – independent of the
// thread count.
//
// Model: reads are sampled from a sample genome = reference + planted SVs.  A read
// that carries a DEL/INS shows it either inside its CIGAR or as a split alignment;
// INV / tandem DUP / inter-contig BND / interspersed DUP are always split
// alignments (primary soft-clipped, supplementaries hard-clipped, reciprocal SA
// tags), at most one split event per read.  Small errors follow a CLR-like profile
// (geometric M runs broken by short I / D).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

struct SynthSv {
    int32_t tid, pos, len;
    int32_t type;      // 0 DEL, 1 INS, 2 INV, 3 DUP_TAN, 4 BND, 5 DUP_INT
    int32_t copies;    // DUP_TAN: extra copies
    int32_t tid2, pos2;  // BND: partner locus; DUP_INT: source locus (region [pos2,pos2+len) on tid2)
    int32_t strand2;   // BND: 0 '+', 1 '-' partner orientation
    float vaf;
    int32_t pad;
    int64_t allele_off;  // INS: offset of the inserted sequence (ASCII) in allele blob
};

struct SynthConfig {
    uint64_t seed;
    int32_t n_contigs;
    int32_t pad0;
    const int64_t* contig_len;
    const char* contig_names;     // concatenated, NUL separated
    int64_t n_reads;
    double len_mean, len_sd;
    int32_t len_min, len_max;
    double p_ins, p_del;          // per reference base
    double geo_ins, geo_del;      // geometric success prob of the error length
    double p_lowmapq, p_secondary, p_unmapped;
    double p_split;               // carried DEL/INS shown as split read instead of in CIGAR
    int64_t n_sv;
    const SynthSv* svs;           // sorted by (tid, pos)
    const uint8_t* alleles;
};

struct SynthSizes { int64_t n_records, cigar_words, seq_bytes, sa_bytes; };

struct SynthOut {
    int32_t* tid; int32_t* pos; uint16_t* flag; uint8_t* mapq; uint32_t* n_cigar; uint64_t* cigar_off;
    int32_t* l_seq; uint64_t* seq_off; uint64_t* sa_off; uint32_t* sa_len; uint32_t* qname_id;
    uint32_t* cigar; uint8_t* seq; uint8_t* sa;
};

}  // extern "C"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    inline double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline int64_t below(int64_t n) { return (int64_t)(uni() * (double)n); }
    inline int geo(double p) {  // support 1,2,...
        double u = uni();
        if (u <= 0) u = 1e-300;
        return 1 + (int)std::floor(std::log(u) / std::log(1.0 - p));
    }
    inline double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 <= 0) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

inline uint64_t mix(uint64_t a, uint64_t b) {
    Rng r(a * 0x9e3779b97f4a7c15ull ^ (b + 0x7f4a7c15ull));
    r.next();
    return r.next();
}

// one aligned piece of a read, in read order
struct Piece {
    int32_t tid; int64_t lo, hi; int rev;   // reference interval, strand
    int64_t gap_after;                      // unaligned read bases after this piece
    int64_t gap_allele;                     // allele offset filling that gap (-1 random)
    int32_t gap_allele_len;
    uint64_t gap_seed;
};

struct CigEvent { int64_t pos; int type; int32_t len; int64_t allele_off; uint64_t seed; };

struct Placement { int64_t q_off; int64_t allele_off; int32_t allele_len; uint64_t seed; int64_t out_len; };

// mutated copy of an allele; returns length; writes ASCII bases if out != nullptr
int64_t mutate_allele(const uint8_t* al, int32_t len, uint64_t seed, uint8_t* out) {
    static const char B[4] = {'A', 'C', 'G', 'T'};
    Rng r(seed);
    int64_t n = 0;
    for (int32_t i = 0; i < len; ++i) {
        double u = r.uni();
        if (u < 0.02) continue;
        if (u < 0.05) { uint8_t x = B[r.next() & 3]; if (out) out[n] = x; ++n; }
        uint8_t c = al[i];
        if (u >= 0.05 && u < 0.065) c = B[r.next() & 3];
        if (out) out[n] = c;
        ++n;
    }
    return n;
}

struct Body {
    std::vector<uint32_t> ops;
    std::vector<Placement> ins;   // q_off relative to body start (reference orientation)
    int64_t qlen = 0, n_ops = 0;
    bool store = false;
    inline void emit(uint32_t op, int64_t len) {
        if (len <= 0) return;
        // BAM op length is 28 bits
        while (len > 0) {
            int64_t l = std::min<int64_t>(len, (1 << 28) - 1);
            if (store) ops.push_back((uint32_t)(l << 4) | op);
            ++n_ops; len -= l;
        }
    }
};

// CLR-like walk over [lo,hi) with in-CIGAR events (sorted by pos, inside (lo,hi))
void gen_body(const SynthConfig& c, uint64_t seed, int64_t lo, int64_t hi, const std::vector<CigEvent>& ev, Body& b) {
    Rng r(seed);
    const double p_err = c.p_ins + c.p_del;
    const double p_is_ins = p_err > 0 ? c.p_ins / p_err : 0;
    int64_t cur = lo;
    size_t e = 0;
    int64_t m_run = 0;  // pending M length
    while (cur < hi) {
        int64_t bound = hi;
        bool is_ev = false;
        while (e < ev.size() && ev[e].pos <= cur) ++e;   // skipped (overlapped) events
        if (e < ev.size() && ev[e].pos < hi - 1) { bound = ev[e].pos; is_ev = true; }
        int64_t run = p_err > 0 ? r.geo(p_err) : (hi - lo + 1);
        if (cur + run >= bound) {
            m_run += bound - cur; cur = bound;
            if (!is_ev) break;
            const CigEvent& v = ev[e++];
            if (v.type == 0) {  // DEL
                int64_t dl = std::min<int64_t>(v.len, hi - 1 - cur);
                if (dl > 0) { b.emit(0, m_run); b.qlen += m_run; m_run = 0; b.emit(2, dl); cur += dl; }
            } else {            // INS
                int64_t il = mutate_allele(c.alleles + v.allele_off, v.len, v.seed, nullptr);
                if (il > 0) {
                    b.emit(0, m_run); b.qlen += m_run; m_run = 0;
                    b.ins.push_back({b.qlen, v.allele_off, v.len, v.seed, il});
                    b.emit(1, il); b.qlen += il;
                }
            }
            continue;
        }
        m_run += run; cur += run;
        if (r.uni() < p_is_ins) {
            int il = r.geo(c.geo_ins);
            b.emit(0, m_run); b.qlen += m_run; m_run = 0;
            b.emit(1, il); b.qlen += il;
        } else {
            int64_t dl = r.geo(c.geo_del);
            if (cur + dl >= bound) continue;   // keep a trailing M before any boundary
            b.emit(0, m_run); b.qlen += m_run; m_run = 0;
            b.emit(2, dl); cur += dl;
        }
    }
    b.emit(0, m_run); b.qlen += m_run;
}

struct RecMeta {
    int32_t tid, pos; uint16_t flag; uint8_t mapq; uint32_t n_cigar; int32_t l_seq; uint32_t sa_len;
    uint32_t read; uint8_t piece;
};

struct ReadPlan {
    int n_rec = 0;
    RecMeta rec[4];
};

struct Gen {
    const SynthConfig& c;
    std::vector<int64_t> contig_cum;
    std::vector<std::string> names;
    std::vector<int64_t> sv_begin;   // first sv index per contig (+1 sentinel)
    explicit Gen(const SynthConfig& cfg) : c(cfg) {
        contig_cum.assign(c.n_contigs + 1, 0);
        const char* p = c.contig_names;
        for (int i = 0; i < c.n_contigs; ++i) {
            contig_cum[i + 1] = contig_cum[i] + c.contig_len[i];
            names.emplace_back(p); p += names.back().size() + 1;
        }
        sv_begin.assign(c.n_contigs + 1, c.n_sv);
        for (int64_t k = c.n_sv - 1; k >= 0; --k) sv_begin[c.svs[k].tid] = k;
        for (int i = c.n_contigs - 1; i >= 0; --i) if (sv_begin[i] > sv_begin[i + 1]) sv_begin[i] = sv_begin[i + 1];
        // contigs without SVs inherit the next contig's begin; fix so ranges are empty
        for (int i = 0; i < c.n_contigs; ++i) {
            int64_t b = sv_begin[i];
            if (b < c.n_sv && c.svs[b].tid != i) sv_begin[i] = sv_begin[i + 1];
        }
    }

    // Build the read: pieces in read order + in-CIGAR events per piece.
    // Returns false for an unmapped read.
    void make_read(int64_t r, ReadPlan* plan, SynthOut* out, const std::vector<int64_t>* slots) const {
        Rng rng(mix(c.seed, (uint64_t)r));
        plan->n_rec = 0;
        // --- unmapped / secondary --------------------------------------------------------
        double ucls = rng.uni();
        bool unmapped = ucls < c.p_unmapped;
        bool secondary = !unmapped && ucls < c.p_unmapped + c.p_secondary;
        int64_t L = (int64_t)std::llround(c.len_mean + c.len_sd * rng.normal());
        L = std::max<int64_t>(c.len_min, std::min<int64_t>(c.len_max, L));
        int64_t g = rng.below(contig_cum[c.n_contigs]);
        int tid = (int)(std::upper_bound(contig_cum.begin(), contig_cum.end(), g) - contig_cum.begin()) - 1;
        int64_t clen = c.contig_len[tid];
        if (L > clen - 2) L = clen - 2;
        int64_t lo = rng.below(clen - L);
        int64_t hi = lo + L;
        int read_rev = (int)(rng.next() & 1);

        if (unmapped) {
            RecMeta& m = plan->rec[plan->n_rec++];
            m = RecMeta{-1, -1, 4, 0, 0, (int32_t)std::min<int64_t>(L, 2000), 0, (uint32_t)r, 0};
            if (out && (*slots)[0] >= 0) {
                int64_t k = (*slots)[0];
                fill_fixed(out, k, m);
                fill_random_seq(out->seq + out->seq_off[k], m.l_seq, mix(c.seed ^ 0x5eed, (uint64_t)r));
            }
            return;
        }
        // --- carried SVs ----------------------------------------------------------------------
        std::vector<Piece> pieces;
        std::vector<std::vector<CigEvent>> pev;
        std::vector<CigEvent> cig;     // in-CIGAR events on the main contig interval
        const SynthSv* split = nullptr;
        int64_t split_pos = 0, split_len = 0;
        if (!secondary) {
            const int64_t margin = 300;
            int64_t b0 = sv_begin[tid], b1 = sv_begin[tid + 1];
            const SynthSv* first = c.svs + b0; const SynthSv* last = c.svs + b1;
            const SynthSv* it = std::lower_bound(first, last, lo + margin, [](const SynthSv& s, int64_t v) { return s.pos < v; });
            int64_t prev_end = lo + margin;
            for (; it < last && it->pos < hi - margin; ++it) {
                uint64_t sseed = mix(mix(c.seed, (uint64_t)r), (uint64_t)(it - c.svs) + 17);
                Rng sr(sseed);
                if (sr.uni() >= it->vaf) continue;
                int64_t jpos = it->pos + (int64_t)std::llround(10.0 * sr.normal());
                int64_t jlen = it->len;
                if (it->type == 0 || it->type == 2 || it->type == 3)
                    jlen = std::max<int64_t>(1, (int64_t)std::llround(it->len * (1.0 + 0.05 * (2 * sr.uni() - 1))));
                int64_t foot = (it->type == 1 || it->type == 4 || it->type == 5) ? 1 : jlen;   // reference footprint
                if (it->type == 3) foot = 1;
                if (jpos < prev_end + 50) continue;
                bool wants_split = (it->type >= 2) || (sr.uni() < c.p_split);
                if (wants_split) {
                    if (split) { if (it->type >= 2) continue; wants_split = false; }
                }
                if (it->type <= 1 && !wants_split) {
                    if (jpos + foot >= hi - margin) continue;
                    cig.push_back({jpos, it->type, (int32_t)jlen, it->allele_off, sseed ^ 0xabcdef});
                    prev_end = jpos + foot;
                } else {
                    // geometry checks for split events
                    if (it->type == 0 && jpos + jlen >= hi - margin) continue;
                    if (it->type == 3 && jpos - jlen < 0) continue;
                    split = it; split_pos = jpos; split_len = jlen;
                    prev_end = jpos + ((it->type == 0 || it->type == 2) ? jlen : 1);
                    if (it->type == 2 || it->type == 3) prev_end = hi;   // keep the rest simple after INV/DUP
                    if (it->type == 4 || it->type == 5) prev_end = hi;
                }
            }
        }
        auto add_piece = [&](int32_t t, int64_t a, int64_t b, int rev) {
            pieces.push_back(Piece{t, a, b, rev, 0, -1, 0, 0});
        };
        if (!split) {
            add_piece(tid, lo, hi, 0);
        } else {
            const SynthSv& s = *split;
            uint64_t gseed = mix(mix(c.seed, (uint64_t)r), 0x6a9);
            switch (s.type) {
            case 0:  // split DEL
                add_piece(tid, lo, split_pos, 0); add_piece(tid, split_pos + split_len, hi, 0); break;
            case 1: {  // split INS: unaligned gap holding the inserted sequence
                add_piece(tid, lo, split_pos, 0);
                pieces.back().gap_after = mutate_allele(c.alleles + s.allele_off, s.len, gseed, nullptr);
                pieces.back().gap_allele = s.allele_off; pieces.back().gap_allele_len = s.len; pieces.back().gap_seed = gseed;
                add_piece(tid, split_pos, hi, 0); break;
            }
            case 2: {  // INV of [p, p+len): sample = ref[..p) + rc(ref[p..p+len)) + ref[p+len..)
                int64_t p = split_pos, q = split_pos + split_len;
                if (lo < p) {
                    add_piece(tid, lo, p, 0);
                    if (hi <= q) add_piece(tid, q - (hi - p), q, 1);
                    else { add_piece(tid, p, q, 1); add_piece(tid, q, hi, 0); }
                } else add_piece(tid, lo, hi, 0);
                break;
            }
            case 3: {  // tandem DUP of [p-len, p) inserted at p, `copies` extra copies
                int64_t p = split_pos, a = split_pos - split_len;
                add_piece(tid, lo, p, 0);
                int64_t remaining = hi - p;
                for (int k = 0; k < s.copies && remaining > 200; ++k) {
                    int64_t take = std::min<int64_t>(split_len, remaining);
                    if (k == s.copies - 1 || take < split_len) { add_piece(tid, a, std::min<int64_t>(a + remaining, c.contig_len[tid] - 1), 0); remaining = 0; }
                    else { add_piece(tid, a, p, 0); remaining -= split_len; }
                }
                break;
            }
            case 4: {  // reciprocal-style junction to (tid2,pos2)
                int64_t rest = hi - split_pos;
                add_piece(tid, lo, split_pos, 0);
                int64_t l2 = c.contig_len[s.tid2];
                if (!s.strand2) add_piece(s.tid2, s.pos2, std::min<int64_t>(s.pos2 + rest, l2 - 1), 0);
                else add_piece(s.tid2, std::max<int64_t>(1, s.pos2 - rest), s.pos2, 1);
                break;
            }
            default: {  // 5: interspersed DUP: source [pos2,pos2+len) on tid2 inserted at (tid,pos)
                add_piece(tid, lo, split_pos, 0);
                int64_t rest = hi - split_pos;
                if (rest > s.len + 300) { add_piece(s.tid2, s.pos2, s.pos2 + s.len, 0); add_piece(tid, split_pos, hi - s.len, 0); }
                else add_piece(s.tid2, s.pos2, s.pos2 + std::max<int64_t>(200, std::min<int64_t>(rest, s.len)), 0);
                break;
            }
            }
            // drop degenerate pieces
            std::vector<Piece> ok;
            for (auto& pc : pieces) if (pc.hi - pc.lo >= 100) ok.push_back(pc); else if (!ok.empty()) ok.back().gap_after += 0;
            pieces.swap(ok);
            if (pieces.empty()) add_piece(tid, lo, hi, 0);
        }
        if (pieces.size() > 4) pieces.resize(4);
        // distribute in-CIGAR events onto the pieces lying on the main contig, forward strand
        pev.resize(pieces.size());
        for (auto& ev : cig)
            for (size_t k = 0; k < pieces.size(); ++k)
                if (pieces[k].tid == tid && !pieces[k].rev && ev.pos > pieces[k].lo + 100 && ev.pos + (ev.type == 0 ? ev.len : 1) < pieces[k].hi - 100) {
                    pev[k].push_back(ev); break;
                }
        // whole read on the reverse strand: reverse piece order, flip strands
        if (read_rev) {
            std::vector<Piece> rp(pieces.rbegin(), pieces.rend());
            std::vector<std::vector<CigEvent>> re(pev.rbegin(), pev.rend());
            for (size_t k = 0; k < rp.size(); ++k) {
                rp[k].rev ^= 1;
                // gap_after moves to the previous piece in the new order
            }
            for (size_t k = 0; k < rp.size(); ++k) { rp[k].gap_after = 0; rp[k].gap_allele = -1; }
            for (size_t k = 0; k + 1 < pieces.size(); ++k) {
                size_t nk = pieces.size() - 2 - k;   // gap after old k sits after new index (n-2-k)
                rp[nk].gap_after = pieces[k].gap_after; rp[nk].gap_allele = pieces[k].gap_allele;
                rp[nk].gap_allele_len = pieces[k].gap_allele_len; rp[nk].gap_seed = pieces[k].gap_seed;
            }
            pieces.swap(rp); pev.swap(re);
        }
        const int K = (int)pieces.size();
        // --- bodies --------------------------------------------------------------------------------
        std::vector<Body> body(K);
        std::vector<int64_t> qs(K), qe(K);
        int64_t q = 0;
        for (int k = 0; k < K; ++k) {
            body[k].store = out != nullptr;
            gen_body(c, mix(mix(c.seed, (uint64_t)r), 100 + k), pieces[k].lo, pieces[k].hi, pev[k], body[k]);
            qs[k] = q; q += body[k].qlen; qe[k] = q; q += pieces[k].gap_after;
        }
        const int64_t Lq = q;
        int prim = 0;
        for (int k = 1; k < K; ++k) if (body[k].qlen > body[prim].qlen) prim = k;
        // --- per-piece record metadata -----------------------------------------------------------
        std::vector<uint8_t> mapq(K);
        for (int k = 0; k < K; ++k) {
            Rng mr(mix(mix(c.seed, (uint64_t)r), 500 + k));
            mapq[k] = mr.uni() < c.p_lowmapq ? 5 : 60;
        }
        auto sa_entry = [&](int k, char* buf) -> int {
            int64_t lead = pieces[k].rev ? (Lq - qe[k]) : qs[k];
            int64_t trail = pieces[k].rev ? qs[k] : (Lq - qe[k]);
            int64_t ql = body[k].qlen, rl = pieces[k].hi - pieces[k].lo;
            int n = std::sprintf(buf, "%s,%lld,%c,", names[pieces[k].tid].c_str(), (long long)pieces[k].lo + 1, pieces[k].rev ? '-' : '+');
            if (lead > 0) n += std::sprintf(buf + n, "%lldS", (long long)lead);
            if (ql >= rl) { n += std::sprintf(buf + n, "%lldM", (long long)rl); if (ql > rl) n += std::sprintf(buf + n, "%lldI", (long long)(ql - rl)); }
            else { n += std::sprintf(buf + n, "%lldM%lldD", (long long)ql, (long long)(rl - ql)); }
            if (trail > 0) n += std::sprintf(buf + n, "%lldS", (long long)trail);
            n += std::sprintf(buf + n, ",%d,%lld;", (int)mapq[k], (long long)(ql / 8));
            return n;
        };
        char buf[4][192];
        int blen[4];
        for (int k = 0; k < K; ++k) blen[k] = K > 1 ? sa_entry(k, buf[k]) : 0;
        for (int k = 0; k < K; ++k) {
            bool is_prim = (k == prim);
            int64_t lead = pieces[k].rev ? (Lq - qe[k]) : qs[k];
            int64_t trail = pieces[k].rev ? qs[k] : (Lq - qe[k]);
            RecMeta& m = plan->rec[plan->n_rec++];
            m.tid = pieces[k].tid; m.pos = (int32_t)pieces[k].lo;
            m.flag = (uint16_t)((pieces[k].rev ? 16 : 0) | (is_prim ? 0 : 2048) | (secondary ? 256 : 0));
            m.mapq = mapq[k];
            m.n_cigar = (uint32_t)(body[k].n_ops + (lead > 0) + (trail > 0));
            m.l_seq = (int32_t)(is_prim ? Lq : body[k].qlen);
            m.sa_len = 0;
            for (int j = 0; j < K; ++j) if (j != k) m.sa_len += blen[j];
            m.read = (uint32_t)r; m.piece = (uint8_t)k;
            if (!out) continue;
            // ---- fill ---------------------------------------------------------------------------
            int64_t slot = (*slots)[k];
            if (slot < 0) continue;                       // record outside the range being filled (synth_fill_range)
            fill_fixed(out, slot, m);
            uint32_t* cg = out->cigar + out->cigar_off[slot];
            size_t w = 0;
            uint32_t clip = is_prim ? 4u : 5u;
            if (lead > 0) cg[w++] = (uint32_t)(lead << 4) | clip;
            std::memcpy(cg + w, body[k].ops.data(), body[k].ops.size() * 4); w += body[k].ops.size();
            if (trail > 0) cg[w++] = (uint32_t)(trail << 4) | clip;
            while (w & 3) cg[w++] = 0;
            uint8_t* sa = out->sa + out->sa_off[slot];
            for (int j = 0; j < K; ++j) if (j != k) { std::memcpy(sa, buf[j], blen[j]); sa += blen[j]; }
            // SEQ: random bases, then overwrite the inserted sequences (reference orientation)
            uint8_t* sq = out->seq + out->seq_off[slot];
            fill_random_seq(sq, m.l_seq, mix(mix(c.seed, (uint64_t)r), 900 + k));
            int64_t base = is_prim ? lead : 0;   // BAM-SEQ offset of this record's aligned part
            std::vector<uint8_t> tmp;
            for (auto& pl : body[k].ins) {
                tmp.resize(pl.out_len);
                mutate_allele(c.alleles + pl.allele_off, pl.allele_len, pl.seed, tmp.data());
                put_bases(sq, base + pl.q_off, tmp.data(), pl.out_len);
            }
            if (is_prim) {
                for (int j = 0; j + 1 < K; ++j) {
                    if (pieces[j].gap_after <= 0 || pieces[j].gap_allele < 0) continue;
                    // gap occupies read coords [qe[j], qs[j+1]); BAM coords depend on the primary's strand
                    int64_t off = pieces[prim].rev ? (Lq - qs[j + 1]) : qe[j];
                    tmp.resize(pieces[j].gap_after);
                    mutate_allele(c.alleles + pieces[j].gap_allele, pieces[j].gap_allele_len, pieces[j].gap_seed, tmp.data());
                    put_bases(sq, off, tmp.data(), pieces[j].gap_after);
                }
            }
        }
    }

    static void fill_fixed(SynthOut* o, int64_t k, const RecMeta& m) {
        o->tid[k] = m.tid; o->pos[k] = m.pos; o->flag[k] = m.flag; o->mapq[k] = m.mapq; o->n_cigar[k] = m.n_cigar;
        o->l_seq[k] = m.l_seq; o->sa_len[k] = m.sa_len; o->qname_id[k] = m.read;
    }
    static void fill_random_seq(uint8_t* sq, int64_t l_seq, uint64_t seed) {
        static const uint8_t NIB[4] = {1, 2, 4, 8};
        Rng r(seed);
        int64_t nb = (l_seq + 1) / 2;
        int64_t i = 0;
        while (i < nb) {
            uint64_t x = r.next();
            for (int t = 0; t < 16 && i < nb; ++t, ++i, x >>= 4) sq[i] = (uint8_t)((NIB[x & 3] << 4) | NIB[(x >> 2) & 3]);
        }
        if (l_seq & 1) sq[nb - 1] &= 0xf0;
    }
    static void put_bases(uint8_t* sq, int64_t off, const uint8_t* ascii, int64_t n) {
        for (int64_t i = 0; i < n; ++i) {
            uint8_t code = ascii[i] == 'A' ? 1 : ascii[i] == 'C' ? 2 : ascii[i] == 'G' ? 4 : ascii[i] == 'T' ? 8 : 15;
            int64_t p = off + i;
            if (p & 1) sq[p >> 1] = (uint8_t)((sq[p >> 1] & 0xf0) | code);
            else sq[p >> 1] = (uint8_t)((sq[p >> 1] & 0x0f) | (code << 4));
        }
    }
};

struct PlanState {
    std::vector<ReadPlan> plans;
    std::vector<int64_t> slot_of;   // [read*4 + piece] -> record slot
    SynthSizes sizes;
    std::vector<RecMeta> sorted;
};

}  // namespace

extern "C" {

// Phase 1: decide every record's metadata and the blob sizes.
void* synth_plan(const SynthConfig* cfg, SynthSizes* sizes) {
    Gen gen(*cfg);
    PlanState* st = new PlanState();
    st->plans.resize(cfg->n_reads);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < cfg->n_reads; ++r) gen.make_read(r, &st->plans[r], nullptr, nullptr);
    std::vector<RecMeta>& recs = st->sorted;
    for (auto& p : st->plans) for (int k = 0; k < p.n_rec; ++k) recs.push_back(p.rec[k]);
    std::stable_sort(recs.begin(), recs.end(), [](const RecMeta& a, const RecMeta& b) {
        uint32_t ta = (uint32_t)a.tid, tb = (uint32_t)b.tid;   // -1 (unmapped) sorts last
        if (ta != tb) return ta < tb;
        return a.pos < b.pos;
    });
    st->slot_of.assign((size_t)cfg->n_reads * 4, -1);
    SynthSizes s{(int64_t)recs.size(), 0, 0, 0};
    for (size_t i = 0; i < recs.size(); ++i) {
        st->slot_of[(size_t)recs[i].read * 4 + recs[i].piece] = (int64_t)i;
        s.cigar_words += (recs[i].n_cigar + 3) & ~3u;
        s.seq_bytes += (recs[i].l_seq + 1) / 2;
        s.sa_bytes += recs[i].sa_len;
    }
    st->sizes = s;
    *sizes = s;
    return st;
}

// Phase 2: write everything into caller-allocated arrays.
void synth_fill(const SynthConfig* cfg, void* handle, SynthOut* out) {
    PlanState* st = (PlanState*)handle;
    Gen gen(*cfg);
    uint64_t co = 0, so = 0, sao = 0;
    for (size_t i = 0; i < st->sorted.size(); ++i) {
        const RecMeta& m = st->sorted[i];
        out->cigar_off[i] = co; co += (m.n_cigar + 3) & ~3u;
        out->seq_off[i] = so; so += (m.l_seq + 1) / 2;
        out->sa_off[i] = sao; sao += m.sa_len;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < cfg->n_reads; ++r) {
        ReadPlan tmp;
        std::vector<int64_t> slots(st->slot_of.begin() + r * 4, st->slot_of.begin() + r * 4 + 4);
        gen.make_read(r, &tmp, out, &slots);
    }
}

// n_cigar of every planned record, in output order (callers shard the record range by CIGAR volume before filling)
void synth_plan_ncigar(void* handle, uint32_t* out) {
    PlanState* st = (PlanState*)handle;
    for (size_t i = 0; i < st->sorted.size(); ++i) out[i] = st->sorted[i].n_cigar;
}

// Blob sizes of the record range [lo, hi) of the plan.
void synth_range_sizes(void* handle, int64_t lo, int64_t hi, SynthSizes* sizes) {
    PlanState* st = (PlanState*)handle;
    SynthSizes s{hi - lo, 0, 0, 0};
    for (int64_t i = lo; i < hi; ++i) { const RecMeta& m = st->sorted[(size_t)i]; s.cigar_words += (m.n_cigar + 3) & ~3u; s.seq_bytes += (m.l_seq + 1) / 2; s.sa_bytes += m.sa_len; }
    *sizes = s;
}

// Phase 2 for one contiguous record range [lo, hi) of the coordinate-sorted output: rows and blobs of those records only, offsets
// relative to the range (one rank's shard of a large input; the other records are never materialised).
void synth_fill_range(const SynthConfig* cfg, void* handle, int64_t lo, int64_t hi, SynthOut* out) {
    PlanState* st = (PlanState*)handle;
    Gen gen(*cfg);
    uint64_t co = 0, so = 0, sao = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const RecMeta& m = st->sorted[(size_t)i];
        out->cigar_off[i - lo] = co; co += (m.n_cigar + 3) & ~3u;
        out->seq_off[i - lo] = so; so += (m.l_seq + 1) / 2;
        out->sa_off[i - lo] = sao; sao += m.sa_len;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < cfg->n_reads; ++r) {
        std::vector<int64_t> slots(4, -1);
        bool any = false;
        for (int k = 0; k < 4; ++k) { const int64_t sl = st->slot_of[(size_t)r * 4 + k]; if (sl >= lo && sl < hi) { slots[k] = sl - lo; any = true; } }
        if (!any) continue;
        ReadPlan tmp;
        gen.make_read(r, &tmp, out, &slots);
    }
}

void synth_free(void* handle) { delete (PlanState*)handle; }

}  // extern "C"

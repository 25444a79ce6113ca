// Per-read logic of COLLECT that is scalar and branchy: SA-tag parsing
// (SVIM_COLLECT.py:44-93) and the split-read decision tree (SVIM_inter.py:24-302).
// One GPU thread runs this per primary alignment that carries an SA tag (kernel
// k_segment_chain in collect.cu); SVIM_HD so tests/hostcheck can drive it on the CPU.
#pragma once
#include "common.cuh"

struct ContigTable {
    int32_t n;
    const char* names;         // concatenated
    const int32_t* name_off;   // n+1
    const int32_t* rank;       // string-order rank per tid
};

struct ChainParams {
    int64_t min_sv, max_sv, tol_g, tol_o;
    int32_t min_mapq, all_bnds;
};

struct PrimaryInfo {
    uint32_t aln_idx, qname_id;
    int64_t l_seq;        // 0 == query_sequence None
    int64_t read_len;     // primary.infer_read_length(), -1 == None
};

// error bits accumulated per read
enum { CH_BAD_FIELDS = 1, CH_NO_READLEN = 2, CH_DATA_ERROR = 4, CH_TOO_MANY = 8 };

SVIM_HD int32_t contig_lookup(const ContigTable& ct, const uint8_t* s, int len) {
    for (int32_t t = 0; t < ct.n; ++t) {
        int32_t o = ct.name_off[t], l = ct.name_off[t + 1] - o;
        if (l != len) continue;
        int k = 0;
        while (k < len && (uint8_t)ct.names[o + k] == s[k]) ++k;
        if (k == len) return t;
    }
    return -1;
}

// Python int(): optional sign, digits (surrounding whitespace/underscores not handled -> data error)
SVIM_HD bool parse_int(const uint8_t* s, int len, int64_t& v) {
    int i = 0; bool neg = false;
    if (len > 0 && (s[0] == '-' || s[0] == '+')) { neg = s[0] == '-'; i = 1; }
    if (i >= len) return false;
    int64_t x = 0;
    for (; i < len; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        if (x < (int64_t)1 << 52) x = x * 10 + (s[i] - '0');
    }
    v = neg ? -x : x;
    return true;
}

SVIM_HD int cigar_char_op(uint8_t c) {
    switch (c) {
        case 'M': return 0; case 'I': return 1; case 'D': return 2; case 'N': return 3; case 'S': return 4;
        case 'H': return 5; case 'P': return 6; case '=': return 7; case 'X': return 8; case 'B': return 9; default: return -1;
    }
}

// retrieve_other_alignments: append the SA-derived segments that pass the MAPQ
// filter of SVIM_COLLECT.py:154.  `chain[0]` is the primary.  Returns new count.
SVIM_HD int parse_sa_segments(const uint8_t* sa, int sa_len, const ContigTable& ct, const ChainParams& p,
                              int64_t prim_l_seq, Seg* chain, int n, uint32_t& err, int cap = SVIM_MAX_SEGMENTS) {
    int i = 0;
    while (i < sa_len) {
        int e = i;
        while (e < sa_len && sa[e] != ';') ++e;
        // element = sa[i:e]
        if (e > i) {
            int fs[7]; int nf = 0; fs[0] = i;
            for (int k = i; k < e; ++k) if (sa[k] == ',') { ++nf; if (nf < 7) fs[nf] = k + 1; }
            ++nf;   // number of fields
            if (nf != 6) { err |= CH_BAD_FIELDS; }
            else {
                fs[6] = e + 1;
                const uint8_t* f[6]; int fl[6];
                for (int k = 0; k < 6; ++k) { f[k] = sa + fs[k]; fl[k] = fs[k + 1] - 1 - fs[k]; }
                int64_t pos = 0, mapq = 0, nm = 0;
                bool ok = parse_int(f[1], fl[1], pos) && parse_int(f[4], fl[4], mapq) && parse_int(f[5], fl[5], nm);
                int32_t tid = contig_lookup(ct, f[0], fl[0]);
                CigarSummary cs; cigsum_init(cs);
                // pysam's cigarstring setter keeps the matches of (\d+)([MIDNSHP=XB]) and ignores anything else
                int64_t num = 0; bool have = false;
                for (int k = 0; k < fl[3]; ++k) {
                    uint8_t c = f[3][k];
                    if (c >= '0' && c <= '9') { if (num < ((int64_t)1 << 40)) num = num * 10 + (c - '0'); have = true; }
                    else { int op = cigar_char_op(c); if (op >= 0 && have) cigsum_add(cs, (uint32_t)op, num); num = 0; have = false; }
                }
                if (cs.n_ops == 0) ok = false;   // no CIGAR: reference_end would be None and the reference raises
                if (!ok || tid < 0) err |= CH_DATA_ERROR;   // the reference raises here
                else {
                    if (mapq < 0 || mapq > 255) mapq = 0;   // OverflowError branch, SVIM_COLLECT.py:81-84
                    if (mapq >= p.min_mapq) {
                        bool rev = !(fl[2] == 1 && f[2][0] == '+');
                        Seg s; int64_t rl; s.tid = tid;
                        cigsum_finish(cs, prim_l_seq, pos - 1, rev ? 1 : 0, s, rl);
                        if (rev && rl < 0) err |= CH_NO_READLEN;      // SVIM_inter.py:31-34: skipped
                        else if (n >= cap) err |= CH_TOO_MANY;       // the caller's arrays are full: the read goes to the large-read pass
                        else chain[n++] = s;
                    }
                }
            }
        }
        i = e + 1;
    }
    return n;
}

// is_similar (SVIM_inter.py:11-21), FP64 in the reference's operation order.
SVIM_HD bool seg_similar(int32_t chr1, double start1, double end1, int32_t chr2, double start2, double end2, double thr, uint32_t& err) {
    double span1 = end1 - start1, span2 = end2 - start2;
    double c1 = floor((start1 + end1) / 2.0), c2 = floor((start2 + end2) / 2.0);
    double mx = span1 > span2 ? span1 : span2;
    if (mx == 0.0) { err |= CH_DATA_ERROR; return false; }   // ZeroDivisionError in the reference
    double pd = fabs(c1 - c2) / 900.0;
    double sd = fabs(span1 - span2) / mx;
    return chr1 == chr2 && (pd + sd) < thr;
}

struct Junction { int32_t d1, d2, c1, c2; int64_t p1, p2; };
struct Tandem { int32_t chr; int64_t start, end; int32_t full, fwd; };

// Emitter concept: void sig(const svim_sig&), void twin(const svim_sig&)
template <class E>
SVIM_HD void emit_bnd(E& out, bool twin, const ContigTable& ct, const PrimaryInfo& pi, uint32_t& ord,
                      int32_t c1, int64_t p1, int32_t d1, int32_t c2, int64_t p2, int32_t d2, uint8_t src_flag) {
    // SignatureTranslocation.__init__ (SVSignature.py:193-214)
    svim_sig s; memset(&s, 0, sizeof(s));
    int32_t r1 = ct.rank[c1], r2 = ct.rank[c2];
    bool keep = r1 < r2 || (r1 == r2 && p1 < p2);
    if (!keep) { int32_t tc = c1; c1 = c2; c2 = tc; int64_t tp = p1; p1 = p2; p2 = tp; int32_t nd1 = d2 ^ 1, nd2 = d1 ^ 1; d1 = nd1; d2 = nd2; }
    s.type = SVIM_BND; s.contig1 = c1; s.contig2 = c2; s.start = (int32_t)p1; s.end = (int32_t)(p1 + 1); s.pos = (int32_t)p2;
    s.flags = (uint8_t)(src_flag | (d1 ? SVIM_F_DIR1_REV : 0) | (d2 ? SVIM_F_DIR2_REV : 0));
    s.aln_idx = pi.aln_idx; s.qname_id = pi.qname_id;
    if (twin) { s.ordinal = ord; out.twin(s); } else { s.ordinal = ord++; out.sig(s); }
}

// analyze_read_segments (SVIM_inter.py:49-300) on the q-sorted chain.
// ord_sig / ord_twin: running emission ordinals (bit 31 set by the caller).
template <class E>
SVIM_HD void analyze_chain(const Seg* chain, int n, const PrimaryInfo& pi, const ChainParams& p, const ContigTable& ct,
                           E& out, uint32_t ord_sig, uint32_t ord_twin, uint32_t& err, Junction* junc, Tandem* tand) {
    int nj = 0, nt = 0;            // junc / tand: room for n entries each (caller's storage)
    const int64_t lo = p.min_sv, hi = p.max_sv, tol_o = p.tol_o, tol_g = p.tol_g;

    for (int k = 0; k + 1 < n; ++k) {
        const Seg& cur = chain[k]; const Seg& nxt = chain[k + 1];
        int64_t dr = nxt.q_start - cur.q_end;
        // breakend where the read leaves `cur` / enters `nxt`
        int64_t p1 = cur.rev ? cur.ref_start : cur.ref_end - 1; int32_t d1 = cur.rev;
        int64_t p2 = nxt.rev ? nxt.ref_end - 1 : nxt.ref_start; int32_t d2 = nxt.rev;
        bool junction = false, twin = false;

        if (cur.tid != nxt.tid) {                                       // SVIM_inter.py:205-240
            if (-tol_o <= dr && dr <= tol_g) junction = true;
        } else if (cur.rev == nxt.rev) {                                // :66-150
            int64_t dref = cur.rev ? (cur.ref_start - nxt.ref_end) : (nxt.ref_start - cur.ref_end);
            if (dr >= -tol_o) {
                if (dref >= -tol_o) {
                    int64_t dev = dr - dref;
                    if (dev >= lo) {
                        if (dref <= tol_g) {                            // INS, :80-94
                            int64_t at = cur.rev ? cur.ref_start : cur.ref_end;
                            int64_t a = 0, slo = 0, shi = 0;
                            bool has = pi.l_seq > 0;
                            if (cur.rev) { if (pi.read_len < 0) has = false; else a = pi.read_len - nxt.q_start; }
                            else a = cur.q_end;
                            if (has) py_slice(a, dev, pi.l_seq, slo, shi);
                            svim_sig s; memset(&s, 0, sizeof(s));
                            s.type = SVIM_INS; s.flags = SVIM_F_SUPPL; s.contig1 = cur.tid; s.contig2 = -1;
                            s.start = (int32_t)at; s.end = (int32_t)(at + dev);
                            s.seq_off = (uint64_t)slo; s.seq_len = (uint32_t)(shi - slo);
                            s.aln_idx = pi.aln_idx; s.qname_id = pi.qname_id; s.ordinal = ord_sig++;
                            out.sig(s);
                        }
                    } else if (-hi <= dev && dev <= -lo) {
                        if (dr <= tol_g) {                              // DEL, :96-106
                            int64_t at = cur.rev ? nxt.ref_end : cur.ref_end;
                            svim_sig s; memset(&s, 0, sizeof(s));
                            s.type = SVIM_DEL; s.flags = SVIM_F_SUPPL; s.contig1 = cur.tid; s.contig2 = -1;
                            s.start = (int32_t)at; s.end = (int32_t)(at - dev);
                            s.aln_idx = pi.aln_idx; s.qname_id = pi.qname_id; s.ordinal = ord_sig++;
                            out.sig(s);
                            if (p.all_bnds) { emit_bnd(out, true, ct, pi, ord_twin, cur.tid, at - 1, 0, cur.tid, at - dev, 0, SVIM_F_SUPPL); ord_twin++; }
                        }
                    } else if (dev < -hi) {
                        if (dr <= tol_g) junction = true;               // :108-116
                    }
                } else if (dref <= -lo) {                               // tandem duplication, :117-150
                    bool full; Tandem t; t.chr = cur.tid;
                    if (!cur.rev) { t.start = nxt.ref_start; t.end = cur.ref_end; full = nxt.ref_end > cur.ref_start; }
                    else { t.start = cur.ref_start; t.end = nxt.ref_end; full = nxt.ref_start < cur.ref_end; }
                    if (full || dref >= -hi) { t.full = full; t.fwd = !cur.rev; tand[nt++] = t; twin = true; }
                    else junction = true;
                }
            }
        } else if (-tol_o <= dr && dr <= tol_g) {                       // inversions, :152-204
            int64_t a = cur.rev ? cur.ref_start : cur.ref_end;
            int64_t b = cur.rev ? nxt.ref_start : nxt.ref_end;
            int dircode = cur.rev ? 2 : 0;       // right_* : left_*
            int64_t size = 0, s0 = 0, s1 = 0; bool cand = false;
            if (nxt.ref_start - cur.ref_end >= -tol_o) { size = b - a; s0 = a; s1 = b; cand = true; }
            else if (cur.ref_start - nxt.ref_end >= -tol_o) { size = a - b; s0 = b; s1 = a; dircode += 1; cand = true; }
            if (cand) {
                if (lo <= size && size <= hi) {
                    svim_sig s; memset(&s, 0, sizeof(s));
                    s.type = SVIM_INV; s.flags = (uint8_t)(SVIM_F_SUPPL | (dircode << SVIM_F_INVDIR_SHIFT)); s.contig1 = cur.tid; s.contig2 = -1;
                    s.start = (int32_t)s0; s.end = (int32_t)s1;
                    s.aln_idx = pi.aln_idx; s.qname_id = pi.qname_id; s.ordinal = ord_sig++;
                    out.sig(s);
                    twin = true;
                } else if (size > hi) junction = true;
            }
        }
        if (junction) {
            emit_bnd(out, false, ct, pi, ord_sig, cur.tid, p1, d1, nxt.tid, p2, d2, SVIM_F_SUPPL);
            Junction j; j.d1 = d1; j.d2 = d2; j.c1 = cur.tid; j.c2 = nxt.tid; j.p1 = p1; j.p2 = p2; junc[nj++] = j;
        }
        if (twin && p.all_bnds) { emit_bnd(out, true, ct, pi, ord_twin, cur.tid, p1, d1, cur.tid, p2, d2, SVIM_F_SUPPL); ord_twin++; }
    }

    // tandem duplications, :242-272.  The direction of the FIRST group is never refreshed (quirk kept).
    if (nt > 0) {
        int32_t g_chr = tand[0].chr, g_dir = tand[0].fwd, copies = 1, anyfull = tand[0].full;
        int64_t sum_s = tand[0].start, sum_e = tand[0].end;
        for (int k = 1; k <= nt; ++k) {
            bool merge = false;
            if (k < nt) {
                double ms = (double)sum_s / (double)copies, me = (double)sum_e / (double)copies;
                merge = seg_similar(g_chr, ms, me, tand[k].chr, (double)tand[k].start, (double)tand[k].end, 0.3, err) && g_dir == tand[k].fwd;
            }
            if (merge) { sum_s += tand[k].start; sum_e += tand[k].end; copies++; anyfull |= tand[k].full; }
            else {
                double ms = (double)sum_s / (double)copies, me = (double)sum_e / (double)copies;
                svim_sig s; memset(&s, 0, sizeof(s));
                s.type = SVIM_DUP_TAN; s.flags = (uint8_t)(SVIM_F_SUPPL | (anyfull ? SVIM_F_FULLY_COVERED : 0)); s.contig1 = g_chr; s.contig2 = -1;
                s.start = (int32_t)(int64_t)ms; s.end = (int32_t)(int64_t)me;   // int() truncation
                s.copies = (uint16_t)(copies > 65535 ? 65535 : copies);
                s.aln_idx = pi.aln_idx; s.qname_id = pi.qname_id; s.ordinal = ord_sig++;
                if (s.end < s.start) err |= CH_DATA_ERROR;   // assert end >= start
                out.sig(s);
                if (k < nt) { g_chr = tand[k].chr; sum_s = tand[k].start; sum_e = tand[k].end; copies = 1; anyfull = tand[k].full; }
            }
        }
    }

    // interspersed duplications, :274-300
    for (int t = 0; t < nj; ++t) {
        const Junction& T = junc[t];
        for (int b = 0; b < t; ++b) {
            const Junction& B = junc[b];
            if (!(B.d1 == T.d2 && B.d2 == T.d1)) continue;
            if (!seg_similar(ct.rank[B.c1], (double)B.p1, (double)(B.p1 + 1), ct.rank[T.c2], (double)T.p2, (double)(T.p2 + 1), 0.1, err)) continue;
            if (ct.rank[B.c2] != ct.rank[T.c1] || B.d1 != B.d2) continue;
            svim_sig s; memset(&s, 0, sizeof(s));
            s.type = SVIM_DUP_INT; s.flags = SVIM_F_SUPPL; s.contig1 = B.c2; s.contig2 = B.c1;
            s.aln_idx = pi.aln_idx; s.qname_id = pi.qname_id;
            if (B.d1 == 0) {
                int64_t size = T.p1 - B.p2 + 1;
                if (!(lo <= size && size <= hi)) continue;
                s.start = (int32_t)B.p2; s.end = (int32_t)(T.p1 + 1);
                s.pos = (int32_t)(int64_t)((double)(B.p1 + 1 + T.p2) / 2.0);
            } else {
                int64_t size = B.p2 - T.p1;
                if (!(lo <= size && size <= hi)) continue;
                s.start = (int32_t)T.p1; s.end = (int32_t)(B.p2 + 1);
                s.pos = (int32_t)(int64_t)((double)(B.p1 + T.p2 + 1) / 2.0);
            }
            s.ordinal = ord_sig++;
            out.sig(s);
        }
    }
}

// stable insertion sort by (q_start, q_end) — sorted(..., key=...) of SVIM_inter.py:49
SVIM_HD void sort_chain(Seg* c, int n) {
    for (int i = 1; i < n; ++i) {
        Seg x = c[i]; int j = i - 1;
        while (j >= 0 && (c[j].q_start > x.q_start || (c[j].q_start == x.q_start && c[j].q_end > x.q_end))) { c[j + 1] = c[j]; --j; }
        c[j + 1] = x;
    }
}

// ---- svim_aln_soa.cigar8 (include/svimgpu.h): byte = len4 << 4 | op; op nibble 0xF = a length-extension byte whose high nibble
// carries 4 more significant bits, extension bytes precede the operation byte (most significant first), pad = 0x0F.
// k_expand_cigar8 gives every lane 16 bytes per round; these are the per-lane pieces (replayed on the host by tests/hostcheck).
SVIM_HD uint32_t c8_ext_bytes(const uint32_t w[4]) {            // how many of the 16 bytes are extension / pad bytes
    uint32_t n = 0;
    for (int q = 0; q < 4; ++q) {
        uint32_t e = ((w[q] & 0x0F0F0F0Fu) + 0x01010101u) & 0x10101010u;      // bit 4 of a byte set <=> its low nibble is 0xF
        e = (e >> 4) * 0x01010101u;                                            // byte sum in the top byte
        n += e >> 24;
    }
    return n;
}

SVIM_HD uint32_t c8_byte(const uint32_t w[4], int k) { return (w[k >> 2] >> (8 * (k & 3))) & 0xFFu; }

// value accumulated by the run of extension bytes that ends the lane's 16 bytes (0 when the last byte is an operation):
// the operation it belongs to is the first one of the NEXT lane
SVIM_HD uint32_t c8_tail(const uint32_t w[4]) {
    int s = 16;
    while (s > 0 && (c8_byte(w, s - 1) & 15u) == 15u) --s;
    uint32_t acc = 0;
    for (int k = s; k < 16; ++k) acc = (acc << 4) | (c8_byte(w, k) >> 4);
    return acc;
}

// decode the lane's 16 bytes in order: `init` = pending extension bits from the previous lane, operations go to dst[out], dst[out+1], ...
// (nothing at or beyond nc); returns the number of operations seen; *bad is set for a length of 2^28 or more
SVIM_HD uint32_t c8_decode_lane(const uint32_t w[4], uint32_t init, uint64_t out, uint64_t nc, uint32_t* dst, uint32_t* bad) {
    uint32_t acc = init, n = 0;
    for (int k = 0; k < 16; ++k) {
        const uint32_t x = c8_byte(w, k), hi = x >> 4;
        if ((x & 15u) == 15u) { if (acc >> 20) *bad = 1u; acc = (acc << 4) | hi; }
        else {
            if (acc >> 24) *bad = 1u;
            const uint32_t len = (acc << 4) | hi;
            if (out + n < nc) dst[out + n] = (len << 4) | (x & 15u);
            ++n; acc = 0;
        }
    }
    return n;
}


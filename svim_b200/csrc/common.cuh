// Shared definitions for the svimgpu kernels.
//
// Functions marked SVIM_HD hold per-item logic that is identical on host and
// device; tests/hostcheck compiles them with g++ (SVIM_HOST_ONLY) so the branchy
// parts can be exercised without a GPU.  They are not a CPU fallback: the C ABI in
// api.cu only ever launches kernels.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/svimgpu.h"

#ifdef SVIM_HOST_ONLY
#define SVIM_HD inline
#define SVIM_D inline
#else
#define SVIM_HD __host__ __device__ __forceinline__
#define SVIM_D __device__ __forceinline__
#endif

// BAM CIGAR op codes (MIDNSHP=X)
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

// SVIM_intra.py:14-29 — ops that advance pos_ref (N does NOT) / pos_read
#define SVIM_MASK_REF_QUIRK 0x185u   // M D = X
#define SVIM_MASK_READ 0x193u        // M I S = X
// htslib bam_cigar2rlen / bam_cigar2qlen(+H)
#define SVIM_MASK_REF_TRUE 0x18Du    // M D N = X
#define SVIM_MASK_QLEN_H 0x1B3u      // M I S H = X

static const int SVIM_MAX_SEGMENTS = 64;   // primary + SA entries a read may have on the per-thread fast path; longer chains take the large-read pass

// What SVIM_inter.py:28-47 reads from one alignment.
struct Seg {
    int32_t tid;
    int64_t ref_start, ref_end;
    int64_t q_start, q_end;
    int32_t rev;
};

// Running pysam-style summary of a CIGAR (see tools/ref_shim/pysam.py docstring).
struct CigarSummary {
    int64_t ref_len;       // sum M,D,N,=,X
    int64_t qlen_h;        // sum M,I,S,H,=,X  (infer_read_length)
    int64_t qsum_nolead;   // sum M,I,=,X  + S while still 0   (query_alignment_end without SEQ)
    int64_t hard;          // sum H
    int64_t lead_s;        // leading soft clip (hard clips skipped)
    int64_t trail_s;       // soft clips after the last non-clip op, index >= 1
    int32_t in_lead;
    int32_t n_ops;
};

SVIM_HD void cigsum_init(CigarSummary& s) {
    s.ref_len = s.qlen_h = s.qsum_nolead = s.hard = s.lead_s = s.trail_s = 0;
    s.in_lead = 1; s.n_ops = 0;
}

SVIM_HD void cigsum_add(CigarSummary& s, uint32_t op, int64_t len) {
    if (op < 9) {
        if ((SVIM_MASK_REF_TRUE >> op) & 1) s.ref_len += len;
        if ((SVIM_MASK_QLEN_H >> op) & 1) s.qlen_h += len;
    }
    if (op == OP_M || op == OP_I || op == OP_EQ || op == OP_X || (op == OP_S && s.qsum_nolead == 0)) s.qsum_nolead += len;
    if (op == OP_H) s.hard += len;
    if (s.in_lead) {
        if (op == OP_S) s.lead_s += len;
        else if (op != OP_H) s.in_lead = 0;
    }
    if (op == OP_S) { if (s.n_ops >= 1) s.trail_s += len; }
    else if (op != OP_H) s.trail_s = 0;
    s.n_ops++;
}

// query_alignment_start / _end / reference_end / infer_read_length from a summary.
SVIM_HD void cigsum_finish(const CigarSummary& s, int64_t l_seq, int64_t ref_start, int32_t rev, Seg& out, int64_t& read_len) {
    int64_t qas = s.lead_s;
    int64_t qae = (l_seq > 0) ? (l_seq - s.trail_s) : s.qsum_nolead;
    read_len = s.qlen_h > 0 ? s.qlen_h : -1;   // -1 == None
    out.ref_start = ref_start;
    out.ref_end = ref_start + (s.ref_len ? s.ref_len : 1);   // htslib bam_endpos
    out.rev = rev;
    if (rev) { out.q_start = read_len - qae; out.q_end = read_len - qas; }
    else { out.q_start = qas; out.q_end = qae; }
}

// Python slice seq[a:a+n] on a sequence of length L -> [lo,hi)
SVIM_HD void py_slice(int64_t a, int64_t n, int64_t L, int64_t& lo, int64_t& hi) {
    int64_t start = a, stop = a + n;
    if (start < 0) { start += L; if (start < 0) start = 0; } else if (start > L) start = L;
    if (stop < 0) { stop += L; if (stop < 0) stop = 0; } else if (stop > L) stop = L;
    lo = start; hi = stop < start ? start : stop;
}

static_assert(sizeof(svim_sig) == 48, "svim_sig layout");
static_assert(sizeof(svim_csig) == 64, "svim_csig layout");
static_assert(sizeof(svim_cluster) == 72, "svim_cluster layout");

/* svimgpu.h — C ABI of the B200-native SVIM COLLECT -> CLUSTER path.
 *
 * The reference (eldariont/svim v2.0.0) has no FFI: the seam is two Python call
 * sites in its CLI script plus the helpers its tests import.  Each entry point
 * below names the reference interface it stands under (file:line relative to the
 * reference tree).  The Python host mirror (svim_b200/SVIM_COLLECT.py,
 * SVIM_CLUSTER.py, SVIM_clustering.py) binds these through ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative svimgpu_status;
 * svimgpu_last_error() gives text.  All pointers are HOST pointers owned by the
 * caller unless stated otherwise; calls are synchronous (the context's stream is
 * drained before returning).  One context per thread.  Plain C types only.
 */
#ifndef SVIMGPU_H
#define SVIMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svimgpu_ctx svimgpu_ctx;

typedef enum {
    SVIMGPU_OK = 0,
    SVIMGPU_ERR_CUDA = -1,      /* CUDA runtime / driver failure (no device, OOM, launch error) */
    SVIMGPU_ERR_ARG = -2,       /* invalid argument */
    SVIMGPU_ERR_STATE = -3,     /* call sequence violated (e.g. cluster before collect) */
    SVIMGPU_ERR_LIMIT = -4,     /* a documented capacity was exceeded */
    SVIMGPU_ERR_DATA = -5,      /* input the reference itself would raise on (ZeroDivisionError, bad SA ints) */
    SVIMGPU_ERR_NCCL = -6,
    SVIMGPU_ERR_PEER = -7       /* multi-GPU: another rank failed before a collective; every rank returns an error, none keeps a result */
} svimgpu_status;

/* Hot-path options: SVIM_input_parsing.py:279-371 (defaults in comments). */
typedef struct {
    int32_t min_mapq;                    /* 20 */
    int32_t min_sv_size;                 /* 40 */
    int32_t max_sv_size;                 /* 100000 */
    int32_t segment_gap_tolerance;       /* 10 */
    int32_t segment_overlap_tolerance;   /* 5 */
    int32_t all_bnds;                    /* 0 */
    double partition_max_distance;       /* 1000 */
    double position_distance_normalizer; /* 900 */
    double edit_distance_normalizer;     /* 1.0 */
    double cluster_max_distance;         /* 0.5 */
} svim_params;

/* Flattened alignment records — replaces the pysam.AlignedSegment stream of
 * SVIM_COLLECT.py:133.  See svim_b200/records.py for the field semantics. */
typedef struct {
    int64_t n_aln;
    const int32_t* tid;
    const int32_t* pos;
    const uint16_t* flag;
    const uint8_t* mapq;
    const uint32_t* n_cigar;
    const uint64_t* cigar_off;   /* in uint32 words, multiple of 4 */
    const int32_t* l_seq;
    const uint64_t* seq_off;     /* bytes */
    const uint64_t* sa_off;      /* bytes */
    const uint32_t* sa_len;
    const uint32_t* qname_id;
    const uint32_t* cigar; int64_t cigar_words;  /* BAM encoding len<<4|op */
    const uint8_t* seq; int64_t seq_bytes;       /* BAM 4-bit packing */
    const uint8_t* sa; int64_t sa_bytes;         /* SA tag text */
    /* Optional 16-bit packed CIGAR stream — half the PCIe bytes of `cigar` for long reads.  When cigar16 is set the library
     * uploads it instead of `cigar` (which may then be NULL) and expands it to the uint32 words on the device.
     * Word = len << 4 | op for len < 4096; words with op nibble 0xF carry 12 more significant length bits each and precede
     * the op word; every record's stream starts at cigar16_off[i] (uint16 units, multiple of 8) and is padded with 0x000F.
     * bamio_pack_cigar16 (csrc_host/bamio.cpp; AlignmentBatch.pack_cigar16) produces it. */
    const uint16_t* cigar16; int64_t cigar16_words;
    const uint64_t* cigar16_off;                 /* n_aln + 1 entries */
    /* Optional 8-bit packed CIGAR stream, taken in preference to cigar16 / cigar when set: noisy long reads have almost only short
     * operations (91 % under 16 bases on the CLR-like workloads), so one byte per operation carries them: 1.09 bytes per operation
     * on BASELINE configs[1] against 2 and 4.  Byte = len << 4 | op for len < 16; bytes with op nibble 0xF carry 4 more
     * significant length bits each and precede the operation byte (most significant first, no leading zero extension);
     * every record's stream starts at cigar8_off[i] (bytes, multiple of 16) and is padded with 0x0F.  svim_b200's host BAM decoder
     * emits it while decoding (csrc_host/bamio.cpp, bamio_set_pack); bamio_pack_cigar8 is the stand-alone packer. */
    const uint8_t* cigar8; int64_t cigar8_bytes;
    const uint64_t* cigar8_off;                  /* n_aln + 1 entries */
} svim_aln_soa;

/* Signature types, in the reference's clustering call order (SVIM_CLUSTER.py:19-24). */
enum { SVIM_DEL = 0, SVIM_INS = 1, SVIM_INV = 2, SVIM_DUP_TAN = 3, SVIM_BND = 4, SVIM_DUP_INT = 5,
       /* cluster-stage only: a DUP_INT *candidate* (partition_and_cluster_candidates, SVIM_clustering.py:306-372):
        * partitioned like a unilocal record on its source interval (SVCandidate.py:24-36), distance of
        * span_position_distance_intdup_candidates (:110-119), no same-read rules; svim_csig.seq_off carries the
        * IEEE-754 bits of the destination END (candidates store it explicitly). Counted under index 5 in the stats. */
       SVIM_DUP_INT_CAND = 6 };
/* svim_sig.flags */
enum {
    SVIM_F_SUPPL = 1,          /* signature source "suppl" (else "cigar") */
    SVIM_F_FULLY_COVERED = 2,  /* DUP_TAN fully_covered */
    SVIM_F_DIR1_REV = 4,       /* BND direction1 == 'rev' */
    SVIM_F_DIR2_REV = 8,       /* BND direction2 == 'rev' */
    SVIM_F_INVDIR_SHIFT = 4    /* INV: bits 4-6 = 0 left_fwd, 1 left_rev, 2 right_fwd, 3 right_rev, 4 all */
};

/* One SVSignature (SVSignature.py:36-233), 48 bytes.
 *   DEL/INS/INV/DUP_TAN: contig1,start,end            INS: seq_off/seq_len into the INS blob
 *   DUP_INT: source contig1,start,end; destination contig2,pos
 *   BND: contig1,start(=pos1) / contig2,pos(=pos2), already in canonical order */
typedef struct {
    int32_t start, end, pos;
    int32_t contig1, contig2;     /* reference ids (tid) */
    uint32_t aln_idx;             /* emitting record */
    uint32_t qname_id;            /* read identity */
    uint32_t ordinal;             /* emission order inside the record; bit 31 = from segment analysis */
    uint64_t seq_off;
    uint32_t seq_len;
    uint8_t type, flags;
    uint16_t copies;              /* DUP_TAN */
} svim_sig;

/* Cluster-stage signature: what span_position_distance / form_partitions read
 * (SVIM_clustering.py:17-96, SVSignature.py get_key/get_source/get_destination).
 * Coordinates are doubles because the reference's public seam accepts floats
 * (tests/test_clustering.py:15-18); integers below 2^53 are exact. 64 bytes. */
typedef struct {
    double start, end;            /* source interval; BND: pos1, pos1+1 */
    double dpos;                  /* DUP_INT destination start; BND pos2 */
    int32_t contig_a;             /* string-order rank of the source contig (BND contig1) */
    int32_t contig_b;             /* rank of the destination contig (DUP_INT, BND) else -1 */
    uint32_t read_id;             /* equal ids <=> equal read names */
    uint32_t seq_len;             /* INS */
    uint64_t seq_off;             /* INS: offset into the INS blob */
    uint8_t type;
    uint8_t dirs;                 /* BND: bit0 dir1 rev, bit1 dir2 rev; INV: direction code */
    uint16_t copies;
    uint32_t pad[3];
} svim_csig;

/* One SignatureClusterUniLocal / BiLocal (SVSignature.py:236-310). */
typedef struct {
    int64_t start, end;           /* int(round(mean)) */
    int64_t dest_start, dest_end; /* bilocal only */
    double score;
    double std_span, std_pos;     /* NaN = None */
    uint32_t member_off, size;    /* members[member_off .. +size) index the input signature array */
    uint8_t type, dir1_rev, dir2_rev, pad0;
    uint32_t pad1;
} svim_cluster;

typedef struct {
    int64_t n_partitions[6], n_clusters[6], large_partitions[6], duplicate_signatures[6];
    int64_t n_members;            /* total member indices */
    int64_t n_clusters_total;
    int64_t myers_pairs, myers_cells;          /* edit distances read / cells of their full DP matrices */
    int64_t myers_banded_pairs;                /* of those, scheduled on the banded first pass ...            */
    int64_t myers_retry_pairs;                 /* ... and handed over to the unbanded kernels (bound exceeded) */
    int64_t myers_band_cells;                  /* cells the banded first pass computed (wavefront + thread-per-pair) */
    int64_t myers_tpp_pairs, myers_tpp_cells;  /* of those: pairs / cells of the thread-per-pair window kernels  */
    int64_t myers_unbanded_cells;              /* cells the unbanded kernels computed: their own pairs + the hand-overs, full matrices */
} svim_cluster_stats;

typedef struct {
    int64_t n_signatures, n_twin_signatures;   /* main list / --all_bnds extras */
    int64_t ins_bytes, twin_ins_bytes;
    int64_t n_sa_bad_fields;   /* SA entries skipped: != 6 fields (warning at SVIM_COLLECT.py:60-62) */
    int64_t n_no_read_length;  /* segments skipped: infer_read_length() None (SVIM_inter.py:31-34) */
    int64_t n_primaries;       /* read_nr of SVIM_COLLECT.py:150 */
    int64_t n_data_errors;     /* conditions the reference raises on (unknown SA contig, non-integer SA field, ...) */
} svim_collect_stats;

/* ---- lifecycle ---------------------------------------------------------------- */
int svimgpu_create(svimgpu_ctx** ctx, int device, const svim_params* params);
void svimgpu_destroy(svimgpu_ctx* ctx);
const char* svimgpu_last_error(const svimgpu_ctx* ctx);
int svimgpu_set_params(svimgpu_ctx* ctx, const svim_params* params);
const char* svimgpu_version(void);
/* PCI bus id ("0000:1b:00.0") of the context's device, for matching it in NVML (bench.py clock sampling). */
int svimgpu_pci_bus_id(svimgpu_ctx* ctx, char* out, int32_t cap);

/* Contig table: names (for SA rname lookup = bam.get_tid, SVIM_COLLECT.py:79) and
 * string-order ranks (Python `str <` of SVSignature.py:194 and the tuple sort of
 * SVIM_clustering.py:19).  names = concatenated, name_off[n+1]. */
int svimgpu_set_contigs(svimgpu_ctx* ctx, int32_t n_contigs, const char* names, const int32_t* name_off);

/* Reference genome resident on the device (pysam.FastaFile of SVIM_clustering.py:377,
 * fetched at :37-43): concatenated contig bytes, offsets[n+1]. */
int svimgpu_set_genome(svimgpu_ctx* ctx, int32_t n_contigs, const int64_t* offsets, const uint8_t* bytes);

/* Page-lock caller memory so svimgpu_upload_alignments runs at PCIe speed. */
int svimgpu_pin_host(void* p, int64_t bytes);
int svimgpu_unpin_host(void* p);

/* ---- COLLECT: analyze_alignment_file_coordsorted (SVIM_COLLECT.py:132-167) ------ */
/* H2D of the record buffer (pageable or pinned host memory). */
int svimgpu_upload_alignments(svimgpu_ctx* ctx, const svim_aln_soa* soa);
/* CIGAR scan + SA/segment analysis + INS sequence gather on the resident buffer. */
/* ---- BAM file -> resident record buffer on the GPU (the step in front of the path: bam.fetch(until_eof=True), SVIM_COLLECT.py:133) ----
 * One BGZF block of the file: payload offset / inflated offset / payload bytes / inflated bytes (ISIZE).  The host indexes the
 * block headers (svim_b200/csrc_host/bamio.cpp: bamio_open + bamio_blocks) and reads the BAM header; everything else runs on
 * the device: raw DEFLATE, record boundaries, rows, CIGAR / SEQ / SA blobs, read-name ids (first-appearance numbering). */
typedef struct { uint64_t coff, uoff; uint32_t clen, ulen; } svim_bgzf_block;
typedef struct { int64_t n_records, cigar_words, seq_bytes, sa_bytes, names_bytes, n_names, inflated_bytes; } svim_bam_info;
/* file: the whole .bam in host memory (mapped or read); first_record: offset of the first alignment record in the inflated
 * stream; n_ref: contigs in the header.  Leaves the records resident exactly like svimgpu_upload_alignments (svimgpu_collect,
 * svimgpu_genotype run on them).  SVIMGPU_ERR_DATA when the file is malformed or the decoder declines it (a speculative record
 * boundary or a read-name hash did not verify): decode on the host then — the call never guesses. */
int svimgpu_decode_bam(svimgpu_ctx* ctx, const uint8_t* file, int64_t file_bytes, const svim_bgzf_block* blocks, int64_t n_blocks,
                       int64_t first_record, int32_t n_ref, svim_bam_info* info);
/* names: names_bytes of NUL-terminated read names in record order; name_off[n]; rec_of_id[n_names]: first record of every id;
 * qname_id[n] (any pointer may be NULL) */
int svimgpu_fetch_bam_names(svimgpu_ctx* ctx, uint8_t* names, uint64_t* name_off, uint32_t* rec_of_id, uint32_t* qname_id);
/* The resident record buffer back on the host (any pointer may be NULL); sizes from svim_bam_info / the uploaded svim_aln_soa. */
int svimgpu_download_alignments(svimgpu_ctx* ctx, int32_t* tid, int32_t* pos, uint16_t* flag, uint8_t* mapq, uint32_t* n_cigar,
                                uint64_t* cigar_off, int32_t* l_seq, uint64_t* seq_off, uint64_t* sa_off, uint32_t* sa_len,
                                uint32_t* qname_id, uint32_t* cigar, uint8_t* seq, uint8_t* sa);
/* Copy the resident uint32 CIGAR blob back (tests: the device expansion of svim_aln_soa.cigar16 must equal the caller's `cigar`). */
int svimgpu_download_cigar(svimgpu_ctx* ctx, uint32_t* out, int64_t words);
int svimgpu_collect(svimgpu_ctx* ctx, svim_collect_stats* stats);
/* upload + collect */
int svimgpu_collect_host(svimgpu_ctx* ctx, const svim_aln_soa* soa, svim_collect_stats* stats);
/* analyze_alignment_file_querysorted (SVIM_COLLECT.py:96-129): records grouped by read name (consecutive runs of equal
 * qname_id, i.e. bam_iterator :8-41); a read's REAL supplementary records are its segments and are CIGAR-analysed too. */
int svimgpu_collect_host_querysorted(svimgpu_ctx* ctx, const svim_aln_soa* soa, svim_collect_stats* stats);
/* D2H: which = 0 main list (sv_signatures), 1 = translocation_signatures_all_bnds.
 * out_sigs[n_signatures] in the reference's emission order; out_ins[ins_bytes] ASCII. */
int svimgpu_fetch_signatures(svimgpu_ctx* ctx, int which, svim_sig* out_sigs, uint8_t* out_ins);
/* The same lists without a second copy: svimgpu_collect_host[_querysorted] starts a D2H of both lists into pinned memory
 * owned by the context, on a copy stream, so it overlaps CLUSTER.  This call waits for that copy and returns the host
 * pointers (valid until the next collect / exchange / destroy on this context); *sigs = NULL when there is no host copy
 * (lists produced by svimgpu_collect on a resident buffer, or replaced by svimgpu_exchange_signatures). */
int svimgpu_signatures_host(svimgpu_ctx* ctx, int which, const svim_sig** sigs, const uint8_t** ins);

/* ---- CLUSTER: cluster_sv_signatures (SVIM_CLUSTER.py:7-26) ------------------------ */
/* Use the device-resident result of svimgpu_collect as the clustering input. */
int svimgpu_use_collected(svimgpu_ctx* ctx, int which);
/* Or upload an arbitrary signature list (partition_and_cluster seam,
 * SVIM_clustering.py:375; tests/test_clustering.py:53). */
int svimgpu_set_signatures(svimgpu_ctx* ctx, int64_t n, const svim_csig* sigs, const uint8_t* ins_blob, int64_t ins_bytes,
                           const int32_t* contig_rank_to_tid /* nullable; for genome lookup of INS */, int32_t n_ranks);
/* form_partitions + clusters_from_partitions + consolidate_* + final ordering for
 * all six types in one pass. */
int svimgpu_cluster(svimgpu_ctx* ctx, svim_cluster_stats* stats);
/* form_partitions only (SVIM_clustering.py:17-29): key sort + gap split of the selected
 * signatures; read the result with svimgpu_fetch_partitions. */
int svimgpu_partition(svimgpu_ctx* ctx, int64_t* n_partitions);
/* D2H: clusters[n_clusters_total] grouped by type in order DEL, INS, INV, DUP_TAN,
 * BND, DUP_INT (each group in the reference's list order); members[n_members]. */
int svimgpu_fetch_clusters(svimgpu_ctx* ctx, svim_cluster* clusters, uint32_t* members);
/* The same arrays without a second copy: svimgpu_cluster[_sharded] leaves its result in pinned host memory owned by the
 * context; this returns those pointers (valid until the next cluster / destroy on this context). */
int svimgpu_clusters_host(svimgpu_ctx* ctx, const svim_cluster** clusters, const uint32_t** members);
/* form_partitions only (SVIM_clustering.py:17-29): order[n] = signature indices in
 * partition order, part_off[n_partitions+1]; call after svimgpu_partition or svimgpu_cluster. */
int svimgpu_fetch_partitions(svimgpu_ctx* ctx, int64_t* n_partitions, uint32_t* order, uint32_t* part_off);

/* ---- GENOTYPE: genotype(candidates, bam, type, options) (SVIM_genotyping.py:34-93; call sites svim:161-170) ------- */
/* Genotyping options (SVIM_input_parsing.py:404-437); minimum_score (:39) is applied by the caller, which only submits
 * the candidates that pass it. */
typedef struct {
    int32_t min_mapq;               /* 20 */
    int32_t minimum_depth;          /* 4 */
    double homozygous_threshold;    /* 0.8 */
    double heterozygous_threshold;  /* 0.2 */
} svim_geno_params;
/* What genotype() reads from one candidate. 32 bytes. */
typedef struct {
    int64_t start, end;             /* get_source() for DEL/INV, get_destination() for INS/DUP_INT (:43-48) */
    int32_t tid;                    /* contig of that locus (reference id of the record buffer) */
    uint32_t n_variant_reads;       /* distinct read names among candidate.members (:51) = alt_reads */
    uint64_t variant_off;           /* their qname ids: variant_qname_ids[variant_off .. +n), ascending */
} svim_geno_cand;
/* What genotype() writes to one candidate (:78-93). 24 bytes. */
typedef struct {
    double support_fraction;        /* NaN = "." */
    int32_t ref_reads, alt_reads;
    uint8_t genotype;               /* 0 "1/1", 1 "0/1", 2 "0/0", 3 "./." */
    uint8_t status;                 /* 0 ok; what the reference raises: 1 TypeError (reference_end None: record without CIGAR),
                                       2 ValueError (fetch start > stop), 3 ZeroDivisionError (no reads and minimum_depth <= 0) */
    uint16_t pad;
    uint32_t n_fetched;             /* records the region fetch (:49) returned before the 500-alignment cap (:57) stopped it */
} svim_geno_result;
/* Genotype n candidates of one type (SVIM_DEL, SVIM_INV, SVIM_INS, SVIM_DUP_INT) against the alignment records left on the
 * device by the last svimgpu_upload_alignments / svimgpu_collect_host (rows + CIGAR; they must be coordinate-sorted, as
 * bam.fetch needs an index: svim:93-98).  contig_lengths[n_contigs] = bam.get_reference_length (:48). */
int svimgpu_genotype(svimgpu_ctx* ctx, int32_t type, const svim_geno_params* params, int64_t n, const svim_geno_cand* cands,
                     const uint32_t* variant_qname_ids, int64_t n_variant_ids, const int64_t* contig_lengths, int32_t n_contigs,
                     svim_geno_result* out);

/* Cut&paste search of COMBINE, flag_cutpaste_candidates (SVIM_merging.py:12-29): for every interval a (source of a DUP_INT
 * cluster) the closest interval b (source of a deletion cluster) under span_position_distance_clusters
 * (SVIM_clustering.py:99-107; contigs are not compared there).  out_index[n_a] = first index of the minimum (stable sort at
 * :20), -1 when n_b == 0 (the reference raises IndexError); out_distance[n_a] in FP64, the reference's operation order.
 * SVIMGPU_ERR_DATA where the reference raises ZeroDivisionError (both spans zero). */
int svimgpu_closest_source(svimgpu_ctx* ctx, int64_t n_a, const int64_t* a_start, const int64_t* a_end, int64_t n_b,
                           const int64_t* b_start, const int64_t* b_end, double position_distance_normalizer,
                           int64_t* out_index, double* out_distance);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------ */
/* id_bytes: 128-byte ncclUniqueId from svimgpu_nccl_unique_id on rank 0. */
int svimgpu_nccl_unique_id(uint8_t* id_bytes /*128*/);
/* Also probes CUDA IPC between the ranks once; if any rank cannot map a peer's memory all ranks agree to gather the INS blobs
 * instead (see svimgpu_exchange_signatures).  svimgpu_peer_ins_active: 1 = peer-mapped, 0 = gathered. */
int svimgpu_comm_init(svimgpu_ctx* ctx, int nranks, int rank, const uint8_t* id_bytes);
int svimgpu_peer_ins_active(svimgpu_ctx* ctx);
/* allgatherv of the collected signature records of all ranks: after it every rank holds the full lists (record indices made
 * global with aln_base, seq_off pointing into the concatenation of all ranks' INS blobs in rank order).  The INS bytes themselves
 * are not moved: each rank keeps its piece and the others map it (CUDA IPC over NVLink / NVSwitch peer memory).
 * svimgpu_cluster[_sharded] reads from the peers only the sequences its own partitions compare; svimgpu_fetch_signatures and the
 * host mirror of svimgpu_signatures_host assemble the whole blob from the peers on demand.  Consequence: a rank must not start its
 * next collect while another rank may still be reading its piece - the exchange inside svimgpu_cluster_sharded orders the cluster
 * calls; put a barrier (svimgpu_barrier_max) between a fetch of the gathered lists and the next collect.
 * A rank's INS buffer only grows (it is reallocated when a later collect emits more inserted bases than any before; the peers
 * re-map it at the next exchange, and nobody reads the old mapping in between).
 * SVIM_PEER_INS=0 in the environment of every rank: gather the INS blobs as well (every rank holds a private full copy). */
int svimgpu_exchange_signatures(svimgpu_ctx* ctx, uint32_t aln_base, svim_collect_stats* stats);
/* After the exchange the host mirror of svimgpu_signatures_host is restarted for the gathered lists (when the lists came from
 * svimgpu_collect_host).  with_ins = 0: this rank mirrors the gathered RECORDS only (*ins = NULL) - for the ranks of a job that do
 * not build the Signature objects; the default (1) also pulls every rank's INS bytes to this host. */
int svimgpu_mirror_gathered_ins(svimgpu_ctx* ctx, int with_ins);
/* restrict clustering to partitions [p*rank/n, p*(rank+1)/n) and allgatherv the
 * cluster records so every rank ends with the full result. */
int svimgpu_cluster_sharded(svimgpu_ctx* ctx, svim_cluster_stats* stats);
int svimgpu_barrier_max(svimgpu_ctx* ctx, double* value /* in: local, out: max over ranks */);

/* ---- micro entry points used by unit tests (same device code as the pipeline) ----- */
/* analyze_cigar_indel (SVIM_intra.py:8-30): out rows of (pos_ref,pos_read,len,type 1=INS|2=DEL) */
int svimgpu_cigar_indel(svimgpu_ctx* ctx, const uint32_t* cigar, int64_t n, int32_t min_len, int64_t* out, int64_t out_cap, int64_t* n_out);
/* edlib.align(a,b)["editDistance"] (SVIM_clustering.py:45) for n_pairs pairs */
int svimgpu_edit_distance(svimgpu_ctx* ctx, int64_t n_pairs, const uint8_t* blob, const int64_t* a_off, const int32_t* a_len,
                          const int64_t* b_off, const int32_t* b_len, int32_t* out);
/* scipy linkage(y,'average') + fcluster(Z,t,'distance') (SVIM_clustering.py:170-171) */
int svimgpu_linkage_average(svimgpu_ctx* ctx, const double* condensed, int32_t m, double t, double* Z /*(m-1)*4*/, int32_t* T /*m*/);
/* random.seed(1524); random.sample(range(n),100) stream of SVIM_clustering.py:129-134
 * for a list of partition sizes (host RNG, CPython 3.12 algorithm) */
int svimgpu_sample_indices(const int64_t* sizes, int64_t n_sizes, int32_t* out /* 100 per size>100 */);
/* timing of the last collect / cluster call, milliseconds of device time per stage */
int svimgpu_last_timings(svimgpu_ctx* ctx, double* ms, int32_t cap, int32_t* n);
const char* svimgpu_timing_name(int32_t i);
/* CUDA events on the context's stream around an arbitrary sequence of calls (bench.py) */
int svimgpu_timer_start(svimgpu_ctx* ctx);
int svimgpu_timer_stop(svimgpu_ctx* ctx, double* ms);
/* kernels of this library launched so far on this context (CUB/NCCL launches not counted) */
int64_t svimgpu_launch_count(svimgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SVIMGPU_H */

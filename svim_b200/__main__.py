"""`python -m svim_b200 alignment <working_dir> <bam_file> <genome> [options]`

Runs COLLECT -> CLUSTER on the GPU with the option names and defaults of the reference's `svim alignment`
sub-command (SVIM_input_parsing.py:262-371), coordinate- or queryname-sorted input (svim:89-111), and writes the
signature-cluster files the reference writes after CLUSTER (`signatures/{del,ins,inv,dup_tan_source,dup_tan_dest,dup_int,trans}.bed`
and `signatures/all.vcf`, SVIM_CLUSTER.py:29-107).  COMBINE and the final VCF are the
reference's downstream stages: `python -m svim_b200.patch alignment ...` with the reference installed runs the whole
pipeline with COLLECT, CLUSTER, the cut&paste search, the candidate clustering and GENOTYPE rebound to this package.
"""
import argparse
import logging
import os
import sys
import time


def parse(argv):
    ap = argparse.ArgumentParser(prog="svim_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="sub")
    p = sub.add_parser("alignment", help="detect SV signatures and cluster them from an existing alignment")
    p.add_argument("working_dir", type=os.path.abspath)
    p.add_argument("bam_file")
    p.add_argument("genome")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--min_mapq", type=int, default=20)
    p.add_argument("--min_sv_size", type=int, default=40)
    p.add_argument("--max_sv_size", type=int, default=100000)
    p.add_argument("--segment_gap_tolerance", type=int, default=10)
    p.add_argument("--segment_overlap_tolerance", type=int, default=5)
    p.add_argument("--partition_max_distance", type=int, default=1000)
    p.add_argument("--position_distance_normalizer", type=int, default=900)
    p.add_argument("--edit_distance_normalizer", type=float, default=1.0)
    p.add_argument("--cluster_max_distance", type=float, default=0.5)
    p.add_argument("--all_bnds", action="store_true")
    return ap.parse_args(argv)


SVIM_VERSION = "2.0.0"          # version string of the reference this package mirrors (svim:3), written into all.vcf

# signatures/<file>: which list of the 6-tuple (DEL, INS, INV, DUP_TAN, DUP_INT, BND) goes where.  Unilocal clusters have one BED
# line; bilocal ones a source and a destination line, which tandem duplications split over two files and the other two types
# keep together (the file set of the reference's write_signature_clusters_bed, SVIM_CLUSTER.py:29-69).
_UNILOCAL_BEDS = (("del.bed", 0), ("ins.bed", 1), ("inv.bed", 2))
_BILOCAL_BEDS = ((3, ("dup_tan_source.bed", "dup_tan_dest.bed")), (4, ("dup_int.bed", "dup_int.bed")), (5, ("trans.bed", "trans.bed")))

_VCF_HEADER = (
    "##fileformat=VCFv4.3",
    "##source=SVIMV{version}",
    '##ALT=<ID=DEL,Description="Deletion">',
    '##ALT=<ID=INV,Description="Inversion">',
    '##ALT=<ID=DUP,Description="Duplication">',
    '##ALT=<ID=DUP:TANDEM,Description="Tandem Duplication">',
    '##ALT=<ID=INS,Description="Insertion">',
    '##INFO=<ID=END,Number=1,Type=Integer,Description="End position of the variant described in this record">',
    '##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">',
    '##INFO=<ID=SVLEN,Number=.,Type=Integer,Description="Difference in length between REF and ALT alleles">',
    "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO",
)


def write_cluster_beds(working_dir, clusters):
    """signatures/*.bed, file for file what the reference writes after CLUSTER (SVIM_CLUSTER.py:29-69)."""
    out = os.path.join(working_dir, "signatures")
    os.makedirs(out, exist_ok=True)
    lines = {}
    for name, idx in _UNILOCAL_BEDS:
        lines[name] = [c.get_bed_entry() for c in clusters[idx]]
    for idx, (src_file, dst_file) in _BILOCAL_BEDS:
        lines.setdefault(src_file, []); lines.setdefault(dst_file, [])
        for c in clusters[idx]:
            src, dst = c.get_bed_entries()
            lines[src_file].append(src); lines[dst_file].append(dst)
    for name, rows in lines.items():
        with open(os.path.join(out, name), "w") as fh:
            fh.writelines(r + "\n" for r in rows)


def write_cluster_vcf(working_dir, clusters, version=SVIM_VERSION):
    """signatures/all.vcf (SVIM_CLUSTER.py:72-107): DEL, INS, INV and DUP_TAN clusters, sorted by their source triple."""
    out = os.path.join(working_dir, "signatures")
    os.makedirs(out, exist_ok=True)
    records = sorted(((c.get_source(), c.get_vcf_entry()) for idx in (0, 1, 2, 3) for c in clusters[idx]), key=lambda pair: pair[0])
    with open(os.path.join(out, "all.vcf"), "w") as fh:
        fh.writelines(h.format(version=version) + "\n" for h in _VCF_HEADER)
        fh.writelines(str(entry) + "\n" for _src, entry in records)


def main(argv=None):
    options = parse(sys.argv[1:] if argv is None else argv)
    if options.sub != "alignment":
        print("usage: python -m svim_b200 alignment <working_dir> <bam_file> <genome>")
        return 2
    logging.basicConfig(level=logging.DEBUG if options.verbose else logging.INFO, format="%(asctime)s [%(levelname)-7.7s]  %(message)s")
    os.makedirs(options.working_dir, exist_ok=True)
    from .io import read_alignments
    from .SVIM_COLLECT import analyze_alignment_file_coordsorted, analyze_alignment_file_querysorted
    from .SVIM_CLUSTER import cluster_sv_signatures
    t0 = time.perf_counter()
    batch = read_alignments(options.bam_file)
    # svim:89-111: coordinate-sorted and queryname-sorted inputs take different COLLECT paths, anything else is refused
    if batch.sort_order not in ("coordinate", "queryname"):
        logging.error("Input BAM file needs to be coordinate-sorted or queryname-sorted. The given file, however, is unsorted according to its header line.")
        return 1
    logging.info("****************** STEP 1: COLLECT ******************")
    t1 = time.perf_counter()
    if batch.sort_order == "queryname":
        sigs, all_bnds = analyze_alignment_file_querysorted(batch, options)
    else:
        sigs, all_bnds = analyze_alignment_file_coordsorted(batch, options)
    t2 = time.perf_counter()
    for t in ("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND"):
        logging.info("Found {0} signatures of type {1}".format(sum(1 for s in sigs if s.type == t), t))
    logging.info("****************** STEP 2: CLUSTER ******************")
    clusters = cluster_sv_signatures(sigs, options)
    if options.all_bnds:
        extra = cluster_sv_signatures(all_bnds, options)
        clusters = clusters[:5] + (clusters[5] + extra[5],)
    t3 = time.perf_counter()
    write_cluster_beds(options.working_dir, clusters)
    write_cluster_vcf(options.working_dir, clusters)
    logging.info("decode %.2f s, COLLECT %.2f s, CLUSTER %.2f s for %d alignment records (%d signatures, %d clusters)",
                 t1 - t0, t2 - t1, t3 - t2, batch.n, len(sigs), sum(len(c) for c in clusters))
    return 0


if __name__ == "__main__":
    sys.exit(main())

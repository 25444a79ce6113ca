"""Rebind the reference's two hot-path call sites to the B200 path.

    import svim_b200.patch; svim_b200.patch.install()      # before the `svim` script imports its modules
or  python -m svim_b200.patch alignment <workdir> <bam> <genome> [...]   # runs the installed `svim` CLI patched

What is replaced (and nothing else):
  svim.SVIM_COLLECT.analyze_alignment_file_coordsorted   (svim:102)
  svim.SVIM_CLUSTER.cluster_sv_signatures                (svim:132,135)
  svim.SVIM_genotyping.genotype                          (svim:161-170)  - on the record rows COLLECT left in HBM
  svim.SVIM_merging.flag_cutpaste_candidates             (SVIM_COMBINE.py:399)  - the O(#DUP_INT x #DEL) closest-deletion search
  svim.SVIM_clustering.partition_and_cluster_candidates  (SVIM_COMBINE.py:476)  - COMBINE-stage twin of the clustering pipeline
COMBINE / VCF keep running on the objects returned here (same attribute surface).
"""
import runpy
import shutil
import sys


def install():
    import svim.SVIM_COLLECT as ref_collect      # the reference package must be importable
    import svim.SVIM_CLUSTER as ref_cluster
    from . import SVIM_COLLECT, SVIM_CLUSTER
    from .io import read_alignments

    def bam_path(bam, options):
        # `bam` is the pysam.AlignmentFile the CLI opened (svim:91; in `reads` mode svim:80,87 open the aligner's output and
        # there is no options.bam_file): decode the same file into the flattened buffer
        name = getattr(bam, "filename", None)
        if isinstance(name, bytes):
            name = name.decode()
        if not name:
            name = getattr(options, "bam_file", None)
        if not name:
            raise ValueError("svim_b200.patch: cannot tell which alignment file the AlignmentFile object reads")
        return name

    def analyze_alignment_file_coordsorted(bam, options):
        return SVIM_COLLECT.analyze_alignment_file_coordsorted(read_alignments(bam_path(bam, options)), options)

    ref_collect.analyze_alignment_file_coordsorted = analyze_alignment_file_coordsorted
    ref_cluster.cluster_sv_signatures = SVIM_CLUSTER.cluster_sv_signatures

    import svim.SVIM_genotyping as ref_genotyping
    from . import SVIM_genotyping, runtime

    def genotype(candidates, bam, type, options):
        # the flattened buffer of the same file is still resident from COLLECT; `bam` (pysam) is not read again
        batch = getattr(runtime.context(), "resident", None)
        return SVIM_genotyping.genotype(candidates, batch if batch is not None else read_alignments(bam_path(bam, options)), type, options)

    ref_genotyping.genotype = genotype

    # COMBINE binds these two by `from ... import` (SVIM_COMBINE.py:13-14): rebind the defining modules before it is imported,
    # and its own globals when it already was
    import svim.SVIM_merging as ref_merging
    import svim.SVIM_clustering as ref_clustering
    from . import SVIM_merging, SVIM_clustering
    ref_merging.flag_cutpaste_candidates = SVIM_merging.flag_cutpaste_candidates
    ref_clustering.partition_and_cluster_candidates = SVIM_clustering.partition_and_cluster_candidates
    combine = sys.modules.get("svim.SVIM_COMBINE")
    if combine is not None:
        combine.flag_cutpaste_candidates = SVIM_merging.flag_cutpaste_candidates
        combine.partition_and_cluster_candidates = SVIM_clustering.partition_and_cluster_candidates


def main():
    install()
    script = shutil.which("svim")
    if script is None:
        raise SystemExit("the reference's `svim` script is not on PATH")
    sys.argv = [script] + sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

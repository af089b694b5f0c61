"""GPU suite (-m gpu): the drop-in binary.  oracle/_ref/bin/graphtyper_gtb is the UNMODIFIED reference with two functions
replaced by integration/gtb_pool_reader.cpp (index_graph -> gtb_region_begin, parallel_reader_genotype_only -> one
gtb_submit_bam_records per pool + the reference's own pool finalisation).  `graphtyper_gtb genotype --vcf` and
`graphtyper_gtb genotype_sv` must write the same VCF, byte for byte after decompression, as the stock binary on the same
inputs: CLI, graph construction, htslib I/O, per-region iteration, VCF merge and writing are the reference's own code in both."""
import glob
import gzip
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from graphtyper_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def _need(*names):
    for n in names:
        if not os.path.exists(os.path.join(BIN, n)):
            pytest.skip(f"oracle/_ref/bin/{n} not built (needs /root/reference at build time)")


def _vcf_text(out_dir):
    files = sorted(glob.glob(os.path.join(out_dir, "**", "*.vcf.gz"), recursive=True))
    assert files, f"no VCF under {out_dir}"
    text = {}
    for f in files:
        with gzip.open(f, "rt") as fh:
            text[os.path.relpath(f, out_dir)] = fh.read()
    return text


def _run_both(args, tmp, extra_env=None, expect_in_log=None):
    outs = {}
    for exe in ("graphtyper", "graphtyper_gtb"):
        out = os.path.join(tmp, "out_" + exe)
        env = dict(os.environ, TMPDIR=tmp, **(extra_env or {}))
        r = subprocess.run([os.path.join(BIN, exe)] + args + [f"--output={out}", "--verbose"], capture_output=True, text=True,
                           env=env, timeout=900)
        assert r.returncode == 0, f"{exe} failed:\n{r.stderr[-3000:]}"
        outs[exe] = (_vcf_text(out), r.stderr)
    ref, gtb = outs["graphtyper"][0], outs["graphtyper_gtb"][0]
    assert sorted(ref) == sorted(gtb)
    n_records = 0
    for k in ref:
        assert ref[k] == gtb[k], f"{k}: the drop-in binary's VCF differs from the reference's"
        n_records += sum(1 for line in ref[k].splitlines() if line and not line.startswith("#"))
    assert "CPU path" not in outs["graphtyper_gtb"][1], outs["graphtyper_gtb"][1][-2000:]
    if expect_in_log:
        assert expect_in_log in outs["graphtyper_gtb"][1], outs["graphtyper_gtb"][1][-3000:]
    # N4: the pools' calls reach the merge through memory (integration/gtb_vcf_store.cpp) -- with --no_cleanup the
    # reference leaves its cereal + gzip batch files <tmp>/graphtyper_*/it1/<first sample>/<n> behind, the drop-in none
    left = {}
    for exe in ("graphtyper", "graphtyper_gtb"):
        t2 = os.path.join(tmp, "keep_" + exe)
        os.makedirs(t2)
        r = subprocess.run([os.path.join(BIN, exe)] + args + [f"--output={os.path.join(t2, 'out')}", "--no_cleanup"],
                           capture_output=True, text=True, env=dict(os.environ, TMPDIR=t2), timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        left[exe] = [f for f in glob.glob(os.path.join(t2, "graphtyper_*", "it*", "*", "*")) if os.path.basename(f).isdigit()]
    assert left["graphtyper"] and not left["graphtyper_gtb"], left
    return n_records


def test_genotype_vcf_cli_is_byte_identical():
    """`genotype --vcf` on a 60 kb contig (two regions), 3 samples in 2 pools, 10 sites / kb incl. indels."""
    _need("graphtyper", "graphtyper_gtb", "bgzip", "tabix")
    tmp = tempfile.mkdtemp(prefix="gtb_dropin_")
    try:
        ds = synth.make_dataset(length=60000, n_sites=600, n_samples=3, seed=511, coverage=20, err=0.004, lowmapq_rate=0.05,
                                unpaired_rate=0.02, improper_rate=0.03)
        man = synth.write_dataset(ds, tmp, region_size=60000)
        subprocess.run([os.path.join(BIN, "bgzip"), "-f", "-k", man["vcf"]], check=True)
        subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", man["vcf"] + ".gz"], check=True)
        sams = os.path.join(tmp, "sams.txt")
        with open(sams, "w") as f:
            f.write("\n".join(man["regions"][0]["sams"]) + "\n")
        n = _run_both(["genotype", man["fasta"], f"--sams={sams}", f"--region={man['contig']}:1-60000", f"--vcf={man['vcf']}.gz",
                       "--no_bamshrink", "--threads=2"], tmp)
        assert n > 400
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_genotype_vcf_cli_on_indexed_bams_decodes_on_the_device():
    """`genotype --vcf` from indexed BAM files: the drop-in reader hands the files' COMPRESSED bytes to the device
    (gtb_submit_bgzf: the chunks of the region iterators; inflate, filters, merge order, parsing on the GPU) -- the log must say
    so for every pool, and the final VCF must still be the reference's, byte for byte."""
    _need("graphtyper", "graphtyper_gtb", "bgzip", "tabix", "sam2bam")
    tmp = tempfile.mkdtemp(prefix="gtb_dropin_bgzf_")
    try:
        ds = synth.make_dataset(length=60000, n_sites=600, n_samples=3, seed=611, coverage=20, err=0.003, lowmapq_rate=0.05,
                                unpaired_rate=0.02, improper_rate=0.03)
        man = synth.write_dataset(ds, tmp, region_size=60000)
        subprocess.run([os.path.join(BIN, "bgzip"), "-f", "-k", man["vcf"]], check=True)
        subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", man["vcf"] + ".gz"], check=True)
        bams = []
        for sam in man["regions"][0]["sams"]:
            bam = sam[:-4] + ".bam"
            subprocess.run([os.path.join(BIN, "sam2bam"), sam, bam], check=True)
            bams.append(bam)
        lst = os.path.join(tmp, "bams.txt")
        with open(lst, "w") as f:
            f.write("\n".join(bams) + "\n")
        n = _run_both(["genotype", man["fasta"], f"--sams={lst}", f"--region={man['contig']}:1-60000", f"--vcf={man['vcf']}.gz",
                       "--no_bamshrink", "--threads=2"], tmp, expect_in_log="records decoded on the device")
        assert n > 400
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_genotype_sv_cli_is_byte_identical():
    """`genotype_sv` on a 30 kb window: <DEL>/<INS>/<DUP> records, 2 samples from indexed BAMs, orphaned mates (leftover
    reads), the SV read filter and the coverage-bin cap of the pool loop, AGGREGATED / BREAKPOINT / COVERAGE records."""
    _need("graphtyper", "graphtyper_gtb", "bgzip", "tabix", "sam2bam")
    tmp = tempfile.mkdtemp(prefix="gtb_dropin_sv_")
    try:
        L, seed = 30000, 71
        ref = synth.make_reference(L, seed)
        sites = synth.make_sv_sites(ref, 6, seed=seed + 1, max_size=600, p_del=0.5, p_ins=0.3)
        gts = synth.make_genotypes(len(sites), 2, seed + 2)
        fa = os.path.join(tmp, "ref.fa")
        synth.write_fasta(fa, ref)
        vcf = os.path.join(tmp, "sv.vcf")
        synth.write_sv_vcf(vcf, sites, "chr1", L)
        subprocess.run([os.path.join(BIN, "bgzip"), "-f", vcf], check=True)
        subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", vcf + ".gz"], check=True)
        rng = np.random.default_rng(seed + 3)
        bams = []
        for k in range(2):
            rs = synth.simulate_reads(ref, sites, gts[k], f"SAMP{k + 1}", seed + 10 + k, coverage=20, err=0.004, lowmapq_rate=0.05)
            keep = np.nonzero((rng.random(len(rs)) > 0.06) & (rs.pos >= 0) & (rs.mpos >= 0))[0]
            sam = os.path.join(tmp, f"s{k}.sam")
            synth.write_sam(sam, rs.subset(keep), "chr1", L)
            bam = os.path.join(tmp, f"s{k}.bam")
            subprocess.run([os.path.join(BIN, "sam2bam"), sam, bam], check=True)
            bams.append(bam)
        lst = os.path.join(tmp, "bams.txt")
        with open(lst, "w") as f:
            f.write("\n".join(bams) + "\n")
        n = _run_both(["genotype_sv", fa, vcf + ".gz", f"--sams={lst}", f"--region=chr1:1-{L}", "--threads=2"], tmp)
        assert n >= 6
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

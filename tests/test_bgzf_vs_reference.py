"""The BGZF path's record batch against the reference's own reader, on real BGZF files: synthetic pools are written as SAM,
converted to BAM by the vendored htslib (oracle/_ref/bin/sam2bam), read by the compiled reference through HtsParallelReader
(gt_probe dumps every record it hands to the pool loop, in its order), and decoded + selected + ordered by
gtb_debug_bgzf_host (the CPU run of the source functions the CUDA kernels call).  Records, order, sample and read-group
columns must be identical -- including the order of exact duplicates inside one file.  Needs the compiled reference
(oracle/_ref: this container, and the GPU box, where it travels prebuilt)."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import bgzf_cases as cases
from graphtyper_b200 import abi, bgzf, engine, gtba, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
pytestmark = pytest.mark.skipif(not (os.path.exists(os.path.join(BIN, "gt_probe")) and os.path.exists(os.path.join(BIN, "sam2bam"))),
                                reason="compiled reference (oracle/_ref) not available")

def doctor_sams(sams, seed=7):
    """Makes the pools harder for the order: every 7th record gets the duplicate flag (filtered by the pool loop AFTER the
    merge, so it still takes part in the per-file sort and in the heap), and one position per file becomes a pile-up of 45
    unpaired copies of a read, a third of them with a changed base (more than 16 records of one position: std::sort switches
    from insertion sort to introsort, and the order of exact duplicates is whatever its pivots make it)."""
    rng = np.random.default_rng(seed)
    for path in sams:
        lines = open(path).read().split("\n")
        head = [l for l in lines if l.startswith("@")]
        body = [l for l in lines if l and not l.startswith("@")]
        out = []
        pile_at = len(body) // 2
        for k, l in enumerate(body):
            f = l.split("\t")
            if k % 7 == 3:
                f[1] = str(int(f[1]) | 0x400)
            out.append("\t".join(f))
            if k == pile_at:
                for c in range(45):
                    g = list(f)
                    g[0] = f"{f[0]}_pile{c}"
                    g[1] = "0"
                    g[6], g[7], g[8] = "*", "0", "0"
                    if c % 3 == 0:
                        seq = list(g[9])
                        at = int(rng.integers(0, len(seq)))
                        seq[at] = "ACGT"[("ACGT".index(seq[at]) + 1 + c % 3) % 4] if seq[at] in "ACGT" else "A"
                        g[9] = "".join(seq)
                    out.append("\t".join(g))
        open(path, "w").write("\n".join(head + out) + "\n")


CASES = {
    "one_file": dict(length=6000, n_sites=40, n_samples=1, seed=131, coverage=40, err=0.0, read_len=100),
    "twelve_files": dict(length=2500, n_sites=25, n_samples=12, seed=171, coverage=45, err=0.001, read_len=100, unpaired_rate=0.05),
    "six_files_many_ties": dict(length=3000, n_sites=30, n_samples=6, seed=141, coverage=60, err=0.0, read_len=100),
    "pileup_and_flags": dict(length=5000, n_sites=40, n_samples=3, seed=151, coverage=30, err=0.0, read_len=100),
    "three_files": dict(length=8000, n_sites=120, n_samples=3, seed=121, coverage=16, err=0.004, n_rate=0.002, lowmapq_rate=0.1,
                        unpaired_rate=0.05, improper_rate=0.08, flip_rate=0.5),
}


def strip_padding(bam):
    """htslib pads read names to four bytes in memory (l_qname counts the padding): the file layout has none."""
    core = bam.core.copy()
    data = bytearray()
    off = np.zeros(len(bam) + 1, np.uint64)
    raw = bam.data.tobytes()
    for k in range(len(bam)):
        d = raw[int(bam.data_off[k]):int(bam.data_off[k + 1])]
        lq = int(core[k]["l_qname"])
        name = d[:lq].split(b"\0")[0] + b"\0"
        core[k]["l_qname"] = len(name)
        data += name + d[lq:]
        off[k + 1] = len(data)
    return abi.HostBamBatch(core, np.frombuffer(bytes(data), np.uint8), off, bam.sample, bam.rg)


def build_case(name, tmp):
    """Dataset -> SAM -> BAM (vendored htslib) -> gt_probe; returns (probe prefix, the reference's record batch in file layout,
    the pool's files as whole-file BGZF segments)."""
    ds = synth.make_dataset(**CASES[name])
    man = synth.write_dataset(ds, tmp, region_size=50000)
    subprocess.run([os.path.join(BIN, "bgzip"), "-f", "-k", man["vcf"]], check=True)
    subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", man["vcf"] + ".gz"], check=True)
    reg = man["regions"][0]
    if name == "pileup_and_flags":
        doctor_sams(reg["sams"])
    bams = []
    for sam in reg["sams"]:
        out = sam[:-4] + ".bam"
        subprocess.run([os.path.join(BIN, "sam2bam"), sam, out], check=True)
        bams.append(out)
    pre = os.path.join(tmp, "probe")
    subprocess.run([os.path.join(BIN, "gt_probe"), "--ref", man["fasta"], "--vcf", man["vcf"] + ".gz", "--region",
                    f"{man['contig']}:{reg['begin']}-{reg['end']}", "--sams", ",".join(bams), "--out", pre], check=True,
                   stdout=subprocess.DEVNULL)
    want = strip_padding(abi.HostBamBatch.from_probe(gtba.load(pre + ".reads.gtba")))
    files, offs = [], []
    for fi, path in enumerate(bams):
        raw = open(path, "rb").read()
        blocks = bgzf.scan_blocks(raw)
        stream = bgzf.inflate_file(raw)
        hlen = bgzf.bam_header_length(stream)
        files.append((raw, [(bgzf.voffset_of(blocks, hlen), len(raw) << 16, True)], fi, fi))
        offs.append([b[0] for b in blocks] + [len(raw)])
    return pre, want, bgzf.HostBgzfFiles(files, offs)


@pytest.mark.parametrize("name", sorted(CASES))
def test_record_batch_equals_the_reference_reader(name):
    tmp = tempfile.mkdtemp(prefix="gtb_bgzf_ref_")
    try:
        pre, want, files = build_case(name, tmp)
        got = engine.bgzf_host(files, bgzf.query(0, 0, 0, whole_file=True))  # gt_probe reads the files without a region
        # exact duplicates within a file exist in these pools: the order below is the reference's own
        n_ties = 0
        for k in range(1, len(want)):
            a, b = want.core[k - 1], want.core[k]
            if a["pos"] == b["pos"] and a["l_qseq"] == b["l_qseq"] and want.sample[k - 1] == want.sample[k]:
                n_ties += 1
        assert n_ties > 0
        cases.assert_batches_equal(got, want, name)
        # htslib never lets a record straddle two BGZF blocks: every file's per-block record walks stitch (and the emulation
        # has checked them against the serial walk record by record)
        assert engine.bgzf_host_stitched() == files.n_files
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

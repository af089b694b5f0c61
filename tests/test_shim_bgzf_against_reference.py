"""The drop-in reader's compressed-bytes hand-off, end to end on the CPU: gtb_shim::BgzfPool::collect on indexed BAM files exactly
as the reference's own HtsReader opened them (the chunks of the REAL .bai index for a region, everything behind the header for
region "."), decoded by the CPU run of the device's source functions (gtb_debug_bgzf_host), against what
HtsParallelReader::read_record + the pool loop's filters hand out on the same files -- record by record, in order
(oracle/ref_build/shim_probe.cpp --bgzf).  Only where the reference was compiled (oracle/_ref)."""
import os
import shutil
import subprocess
import tempfile

import pytest

import oracle
from graphtyper_b200 import synth

EXE = oracle.ref_binary("shim_probe")
SAM2BAM = oracle.ref_binary("sam2bam")
pytestmark = pytest.mark.skipif(EXE is None or SAM2BAM is None, reason="compiled reference (oracle/_ref) not available")


@pytest.fixture(scope="module")
def bams():
    tmp = tempfile.mkdtemp(prefix="gtb_shim_bgzf_")
    ds = synth.make_dataset(length=200000, n_sites=300, n_samples=3, seed=711, coverage=12, err=0.003, lowmapq_rate=0.05,
                            unpaired_rate=0.03, improper_rate=0.04)
    man = synth.write_dataset(ds, tmp, region_size=200000)
    out = []
    for sam in man["regions"][0]["sams"]:
        bam = sam[:-4] + ".bam"
        subprocess.run([SAM2BAM, sam, bam], check=True)
        out.append(bam)
    yield man["contig"], out
    shutil.rmtree(tmp, ignore_errors=True)


def run(region, paths, *extra):
    r = subprocess.run([EXE, "--bgzf", region, ",".join(paths), *extra], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM BGZF PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    return dict(kv.split("=") for kv in r.stdout.split("SHIM BGZF PASS", 1)[1].split())


@pytest.mark.parametrize("span", [(1, 200000), (50001, 100000), (120500, 121000), (199000, 200000), (1, 300)])
def test_region_chunks_of_the_real_index(bams, span):
    """Region iterators: the files are several BGZF blocks long, the chunks start inside blocks and end before the files do."""
    contig, paths = bams
    got = run(f"{contig}:{span[0]}-{span[1]}", paths)
    assert got["whole_file"] == "0" and int(got["files"]) == 3
    if span == (50001, 100000):
        assert int(got["records"]) > 2000
        whole = run(".", paths)
        assert int(got["compressed_bytes"]) < int(whole["compressed_bytes"])  # only the region's chunks were read


def test_whole_files_and_single_file_pools(bams):
    contig, paths = bams
    got = run(".", paths)
    assert got["whole_file"] == "1" and int(got["records"]) > 10000
    one = run(".", paths[:1])
    assert int(one["files"]) == 1 and 0 < int(one["records"]) < int(got["records"])
    run(f"{contig}:70000-90000", paths[1:2])


def test_sv_read_filter_after_the_merge(bams):
    contig, paths = bams
    plain = run(f"{contig}:1-200000", paths)
    sv = run(f"{contig}:1-200000", paths, "--sv")
    assert int(sv["records"]) <= int(plain["records"])

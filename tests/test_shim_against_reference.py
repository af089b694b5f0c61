"""CPU: the reference-side shim (integration/gtb_shim.hpp) compiled against the unmodified reference objects and libgtb200.so
(oracle/ref_build/shim_probe.cpp).  Only where the reference was compiled (oracle/_ref)."""
import os
import shutil
import subprocess
import tempfile

import pytest

import oracle
from graphtyper_b200 import synth

EXE = oracle.ref_binary("shim_probe")


@pytest.mark.skipif(EXE is None or oracle.ref_binary("bgzip") is None, reason="compiled reference not present")
@pytest.mark.parametrize("kw", [dict(length=3000, n_sites=7, n_samples=1, seed=5, coverage=30),
                                dict(length=7000, n_sites=90, n_samples=2, seed=91, coverage=14, err=0.004, complex_sites=True)],
                         ids=["tiny", "complex"])
def test_shim_flattens_and_indexes_like_the_reference(kw):
    """gyper::graph -> gtb_shim::flatten -> gtb_region_begin (host-only context) -> gtb_index_export equals the reference's own
    index_graph (keys, labels, bucket order) inside ONE process; the pool's records are gathered through the reference's
    HtsParallelReader; the compute entry points refuse to run without a device."""
    tmp = tempfile.mkdtemp(prefix="gtb_shim_")
    try:
        ds = synth.make_dataset(**kw)
        man = synth.write_dataset(ds, tmp, region_size=50000)
        subprocess.run([oracle.ref_binary("bgzip"), "-f", "-k", man["vcf"]], check=True)
        subprocess.run([oracle.ref_binary("tabix"), "-f", "-p", "vcf", man["vcf"] + ".gz"], check=True)
        reg = man["regions"][0]
        out = subprocess.run([EXE, man["fasta"], man["vcf"] + ".gz", f"{man['contig']}:{reg['begin']}-{reg['end']}",
                              ",".join(reg["sams"])], capture_output=True, text=True, cwd=tmp)
        assert out.returncode == 0 and out.stdout.startswith("SHIM PASS"), out.stdout + out.stderr
        fields = dict(f.split("=") for f in out.stdout.split()[2:])
        assert int(fields["keys"]) > 1000 and int(fields["records"]) > 100 and int(fields["samples"]) == kw["n_samples"]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


@pytest.mark.gpu
@pytest.mark.skipif(EXE is None or oracle.ref_binary("bgzip") is None, reason="compiled reference not present")
def test_shim_genotypes_like_the_reference_in_one_process():
    """The drop-in, end to end: the reference's graph and records go through the shim and the C ABI to the GPU, and every
    HapSample / per-bubble statistic that comes back equals what the reference's own genotype_only + VcfWriter loop produced
    in the same process."""
    kw = dict(length=8000, n_sites=160, n_samples=3, seed=21, coverage=12, err=0.01, n_rate=0.002, lowmapq_rate=0.1,
              unpaired_rate=0.05, improper_rate=0.08, flip_rate=0.5)
    tmp = tempfile.mkdtemp(prefix="gtb_shim_gpu_")
    try:
        ds = synth.make_dataset(**kw)
        man = synth.write_dataset(ds, tmp, region_size=50000)
        subprocess.run([oracle.ref_binary("bgzip"), "-f", "-k", man["vcf"]], check=True)
        subprocess.run([oracle.ref_binary("tabix"), "-f", "-p", "vcf", man["vcf"] + ".gz"], check=True)
        reg = man["regions"][0]
        out = subprocess.run([EXE, man["fasta"], man["vcf"] + ".gz", f"{man['contig']}:{reg['begin']}-{reg['end']}",
                              ",".join(reg["sams"]), "--gpu"], capture_output=True, text=True, cwd=tmp)
        assert out.returncode == 0 and "SHIM GPU PASS" in out.stdout and "SHIM PASS" in out.stdout, out.stdout + out.stderr
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

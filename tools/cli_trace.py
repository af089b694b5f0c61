"""Timeline of one drop-in CLI run (profiling aid): --vverbose log lines of graphtyper_gtb over the 20 regions."""
import sys, os, subprocess, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, oracle
from graphtyper_b200 import synth
ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
tmp = tempfile.mkdtemp(prefix="gtb_clit_")
B = lambda n: oracle.ref_binary(n)
fa = os.path.join(tmp, "ref.fa"); synth.write_fasta(fa, ref)
vcf = os.path.join(tmp, "sites.vcf"); synth.write_vcf(vcf, sites, "chr1", len(ref))
subprocess.run([B("bgzip"), "-f", vcf], check=True); subprocess.run([B("tabix"), "-f", "-p", "vcf", vcf + ".gz"], check=True)
sam = os.path.join(tmp, "all.sam"); synth.write_sam(sam, rs, "chr1", len(ref))
bam = os.path.join(tmp, "all.bam"); subprocess.run([B("sam2bam"), sam, bam], check=True)
rf = os.path.join(tmp, "regions.txt"); open(rf, "w").write("".join(f"chr1:{b}-{e}\n" for b, e in regions))
for name in ("graphtyper_gtb", "graphtyper_gtb", "graphtyper"):
    out = os.path.join(tmp, "out"); shutil.rmtree(out, ignore_errors=True)
    log = os.path.join(tmp, "log.txt")
    t0 = time.perf_counter()
    subprocess.run([B(name), "genotype", fa, f"--sam={bam}", f"--region_file={rf}", f"--vcf={vcf}.gz", "--threads=1", f"--output={out}",
                    "--vverbose", f"--log={log}"], env=dict(os.environ, TMPDIR=tmp), capture_output=True)
    print(name, "wall", round(time.perf_counter() - t0, 3))
    lines = open(log).read().splitlines()
    keys = ("bamshrink", "Constructing", "gtb200", "Got ", "Merging", "Finished", "Running", "Num of dup", "Writing calls", "Read ")
    sel = [l for l in lines if any(k in l for k in keys)]
    print("\n".join(l[:150] for l in sel[:40]))
    print("... total log lines", len(lines))
shutil.rmtree(tmp, ignore_errors=True)

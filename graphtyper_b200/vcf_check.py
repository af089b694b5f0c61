"""Harness: compares per-bubble calls derived from accumulators with the records of a VCF the reference CLI wrote.

`graphtyper genotype --vcf` writes one record per input site that has a called alt allele (sites whose every alt is "bad"
are dropped at the merge, src/typer/vcf_operations.cpp:600-640); FORMAT = GT:AD:MD:DP:GQ:PL (src/typer/vcf.cpp:1031).  For a
one-sample pool of biallelic bubbles the sample column is Vcf::add_haplotype's SampleCall (src/typer/vcf.cpp:1507-1611) as
it is: GT = first zero PL, AD = coverage, MD = ambiguous depth, DP = AD + MD (SampleCall::get_depth), GQ, PL.
Test infrastructure (used by bench.py's parity counter and by tests); not on the product path."""
from __future__ import annotations

import gzip
from typing import Dict, List, Tuple

import numpy as np


def read_vcf_records(path: str) -> List[Tuple[int, str, str, str, str]]:
    """(pos, ref, alt, format, sample column) of every record of a (b)gzipped one-sample VCF."""
    out = []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith("#"):
                continue
            c = line.rstrip("\n").split("\t")
            out.append((int(c[1]), c[3], c[4], c[8], c[9]))
    return out


def left_normalize(ref_seq: np.ndarray, pos: int, ref: bytes, alt: bytes) -> Tuple[int, bytes, bytes]:
    """The usual VCF normalisation (what the reference's merge does to an indel record, Variant::normalize,
    src/typer/variant.cpp): drop a shared last base, extend both alleles to the left when one runs empty, then drop shared
    leading bases.  ref_seq: the contig as ASCII bytes; pos 1-based."""
    ref, alt = bytearray(ref), bytearray(alt)
    while True:
        if ref and alt and ref[-1] == alt[-1] and (len(ref) > 1 or len(alt) > 1 or pos > 1):
            if len(ref) == 1 and len(alt) == 1:
                break
            ref.pop()
            alt.pop()
            if not ref or not alt:
                pos -= 1
                base = int(ref_seq[pos - 1])
                ref.insert(0, base)
                alt.insert(0, base)
            continue
        break
    while len(ref) > 1 and len(alt) > 1 and ref[0] == alt[0]:
        ref.pop(0)
        alt.pop(0)
        pos += 1
    return pos, bytes(ref), bytes(alt)


def sample_columns(acc, phred: np.ndarray, gt: np.ndarray, gq: np.ndarray, binned: np.ndarray, sample: int = 0) -> Dict[int, str]:
    """bubble id (absolute position of the bubble's var nodes) -> "GT:AD:MD:DP:GQ:PL" as the reference prints it
    (GQ and PL go through the reference's output binning table, vcf.cpp:1107-1113)."""
    NS = acc.n_samples
    out = {}
    for b in range(acc.n_bubbles):
        cnum = int(acc.n_alleles[b])
        tri = cnum * (cnum + 1) // 2
        s0 = int(acc.score_off[b]) * NS + sample * tri
        c0 = int(acc.cov_off[b]) * NS + sample * cnum
        ad = acc.gt_coverage[c0:c0 + cnum].astype(np.int64)
        md = int(acc.ambiguous_depth[b * NS + sample])
        pl = phred[s0:s0 + tri]
        g = gt[(b * NS + sample) * 2:(b * NS + sample) * 2 + 2]
        out[int(acc.bubble_id[b])] = (f"{g[0]}/{g[1]}:{','.join(map(str, ad))}:{md}:{int(ad.sum()) + md}:"
                                      f"{min(int(binned[int(gq[b * NS + sample])]), 99)}:"
                                      f"{','.join(str(int(binned[int(v)])) for v in pl)}")
    return out


def check_region(vcf_path: str, acc, phred, gt, gq, binned: np.ndarray, ref_seq: np.ndarray, sites,
                 contig_offset: int = 0) -> Tuple[int, List[str]]:
    """Every record of the reference's VCF against the calls derived from `acc`: returns (#records checked, mismatches).
    `sites`: the input records (pos / ref / alt) the graph was built from -- a record of the output is matched to its bubble
    through the normalised (pos, REF, ALT) of the input site, so REF / ALT are checked as well."""
    cols = sample_columns(acc, phred, gt, gq, binned)
    by_norm = {}
    for s in sites:
        by_norm[left_normalize(ref_seq, s.pos, s.ref, s.alt)] = s.pos
    bad = []
    n = 0
    for pos, ref, alt, fmt, col in read_vcf_records(vcf_path):
        n += 1
        if fmt != "GT:AD:MD:DP:GQ:PL":
            bad.append(f"pos {pos}: unexpected FORMAT {fmt}")
            continue
        site_pos = by_norm.get((pos, ref.encode(), alt.encode()))
        if site_pos is None:
            bad.append(f"pos {pos} {ref}>{alt}: not an input site")
            continue
        mine = cols.get(site_pos + contig_offset)
        if mine is None:
            bad.append(f"pos {pos}: no bubble at {site_pos}")
        elif mine != col:
            bad.append(f"pos {pos}: reference {col} != {mine}")
    return n, bad

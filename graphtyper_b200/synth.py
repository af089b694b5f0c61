"""Seeded synthetic inputs for the genotyping hot path.

Generates, deterministically from a seed:
  * a uniform random ACGT reference contig (``chr1``),
  * a biallelic SNP/indel VCF with sites on a 12-bp grid (85 % SNP, 8 % deletions of
    1-6 bp, 7 % insertions of 1-6 bp), >= 200 bp from the contig ends,
  * per-sample diploid genotypes uniform over {0/0, 0/1, 0/1, 1/1},
  * paired 150-bp reads (fragment U[320,480], 30x, 0.2 % substitution error, Q40,
    MAPQ 60, flags 99/147 or 83/163), coordinate sorted, with CIGARs derived from
    the haplotype -> reference coordinate map.

The same arrays feed (a) SAM/FASTA/VCF text files consumed by the compiled reference
(``oracle/_ref/bin/graphtyper``) and (b) the read batches handed to the C-ABI, so both
sides see identical records in identical order.

This module is harness code (tests + bench); it is not part of the product path.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Tuple

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


@dataclasses.dataclass
class Site:
    pos: int  # 1-based VCF POS
    ref: bytes
    alt: bytes
    more_alts: Tuple[bytes, ...] = ()  # further ALT alleles of a multi-allelic record (allele index 2, 3, ...)
    info: str = "."                    # VCF INFO (GT_ID / GT_ANTI_HAPLOTYPE events, constructor.cpp:1540-1590)

    def allele(self, a: int) -> bytes:
        return self.ref if a == 0 else (self.alt if a == 1 else self.more_alts[a - 2])

    @property
    def n_alleles(self) -> int:
        return 2 + len(self.more_alts)


@dataclasses.dataclass
class ReadSet:
    """Column store of SAM records for one sample, coordinate sorted."""

    name_id: np.ndarray  # int64 pair id (QNAME = f"{prefix}{id}")
    flag: np.ndarray  # uint16
    pos: np.ndarray  # int64, 0-based leftmost aligned ref position
    mapq: np.ndarray  # uint8
    mpos: np.ndarray  # int64 0-based
    isize: np.ndarray  # int64
    seq: np.ndarray  # uint8 [n, L] ASCII
    cigar: List[str]
    as_tag: np.ndarray  # int32 (AS:i)
    xs_tag: np.ndarray  # int32 (XS:i, -1 = absent)
    sample: str
    name_prefix: str

    def __len__(self) -> int:
        return len(self.flag)

    def subset(self, idx: np.ndarray) -> "ReadSet":
        return ReadSet(
            self.name_id[idx], self.flag[idx], self.pos[idx], self.mapq[idx], self.mpos[idx],
            self.isize[idx], self.seq[idx], [self.cigar[i] for i in idx], self.as_tag[idx],
            self.xs_tag[idx], self.sample, self.name_prefix)


def make_reference(length: int, seed: int = 11) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return BASES[rng.integers(0, 4, size=length)]


def make_sites(ref: np.ndarray, n_sites: int, seed: int = 12, grid: int = 12, margin: int = 200,
               p_snp: float = 0.85, p_del: float = 0.08) -> List[Site]:
    rng = np.random.default_rng(seed)
    L = len(ref)
    first = (margin // grid + 1) * grid
    last = ((L - margin) // grid) * grid
    grid_pos = np.arange(first, last, grid)
    if n_sites > len(grid_pos):
        raise ValueError("too many sites for the grid")
    chosen = np.sort(rng.choice(grid_pos, size=n_sites, replace=False))
    kind = rng.random(n_sites)
    sites: List[Site] = []
    for p, k in zip(chosen.tolist(), kind.tolist()):
        r = int(ref[p - 1])
        if k < p_snp:
            alts = [b for b in b"ACGT" if b != r]
            a = alts[int(rng.integers(0, 3))]
            sites.append(Site(p, bytes([r]), bytes([a])))
        elif k < p_snp + p_del:
            n = int(rng.integers(1, 7))
            sites.append(Site(p, bytes(ref[p - 1:p + n]), bytes([r])))
        else:
            n = int(rng.integers(1, 7))
            ins = BASES[rng.integers(0, 4, size=n)]
            sites.append(Site(p, bytes([r]), bytes([r]) + bytes(ins)))
    return sites


def make_sites_complex(ref: np.ndarray, n_sites: int, seed: int = 12, grid: int = 24, margin: int = 200) -> List[Site]:
    """Sites that exercise what the biallelic generator cannot: multi-allelic SNPs and indels (bubbles with 3-4 alleles),
    records overlapping an earlier deletion, adjacent records (empty reference nodes between bubbles) and haplotype
    events (INFO GT_ID / GT_ANTI_HAPLOTYPE -> VarNode::events / anti_events, used by the index build,
    src/index/indexer.cpp:114-131).  n_sites counts anchor positions; companions add more records."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    first = (margin // grid + 1) * grid
    last = ((L - margin) // grid) * grid
    grid_pos = np.arange(first, last, grid)
    if n_sites > len(grid_pos):
        raise ValueError("too many sites for the grid")
    chosen = np.sort(rng.choice(grid_pos, size=n_sites, replace=False))
    sites: List[Site] = []
    event_id = 0

    def other(base: int, k: int) -> List[bytes]:
        alts = [b for b in b"ACGT" if b != base]
        order = rng.permutation(3)[:k]
        return [bytes([alts[int(i)]]) for i in order]

    def rand_seq(n: int) -> bytes:
        return bytes(BASES[rng.integers(0, 4, size=n)])

    for p in chosen.tolist():
        r = int(ref[p - 1])
        k = rng.random()
        if k < 0.25:  # multi-allelic SNP
            a = other(r, int(rng.integers(2, 4)))
            sites.append(Site(p, bytes([r]), a[0], tuple(a[1:])))
        elif k < 0.35:  # two insertions of different length at one anchor
            i1, i2 = rand_seq(int(rng.integers(1, 4))), rand_seq(int(rng.integers(4, 8)))
            sites.append(Site(p, bytes([r]), bytes([r]) + i1, (bytes([r]) + i2,)))
        elif k < 0.43:  # deletion and a shorter deletion / SNP in one record
            n = int(rng.integers(3, 7))
            refa = bytes(ref[p - 1:p + n])
            snp = bytearray(refa)
            snp[-1] = other(snp[-1], 1)[0][0]
            sites.append(Site(p, refa, bytes([r]), (refa[:2], bytes(snp))))
        elif k < 0.53:  # deletion, then a separate SNP record inside the deleted span
            n = int(rng.integers(3, 7))
            sites.append(Site(p, bytes(ref[p - 1:p + n]), bytes([r])))
            q = p + int(rng.integers(1, n + 1))
            sites.append(Site(q, bytes([int(ref[q - 1])]), other(int(ref[q - 1]), 1)[0]))
        elif k < 0.63:  # adjacent SNP records
            sites.append(Site(p, bytes([r]), other(r, 1)[0]))
            sites.append(Site(p + 1, bytes([int(ref[p])]), other(int(ref[p]), 1)[0]))
        elif k < 0.78:  # event pair within one k-mer: the later alt may not follow the earlier alt (or the reverse order)
            event_id += 1
            q = p + int(rng.integers(2, 12))
            first_info, second_info = (f"GT_ANTI_HAPLOTYPE={event_id}", f"GT_ID={event_id}")
            if rng.random() < 0.3:
                first_info, second_info = second_info, first_info
            sites.append(Site(p, bytes([r]), other(r, 1)[0], (), first_info))
            sites.append(Site(q, bytes([int(ref[q - 1])]), other(int(ref[q - 1]), 1)[0], (), second_info))
        elif k < 0.83:  # STR-like site: many insertion alleles (5-14 alts) -> the 181-combination limit of the index
            n_alt = int(rng.integers(5, 15))
            unit = rand_seq(int(rng.integers(1, 4)))
            alts = []
            for j in range(n_alt):
                a = bytes([r]) + unit * (j + 1)
                if rng.random() < 0.4:
                    a = a + rand_seq(1)
                if a not in alts:
                    alts.append(a)
            sites.append(Site(p, bytes([r]), alts[0], tuple(alts[1:])))
        elif k < 0.9:
            sites.append(Site(p, bytes([r]), other(r, 1)[0]))
        else:
            n = int(rng.integers(1, 7))
            if rng.random() < 0.5:
                sites.append(Site(p, bytes(ref[p - 1:p + n]), bytes([r])))
            else:
                sites.append(Site(p, bytes([r]), bytes([r]) + rand_seq(n)))
    sites.sort(key=lambda s: s.pos)
    return sites


def make_genotypes_complex(sites: List[Site], n_samples: int, seed: int = 13) -> np.ndarray:
    """[n_samples, n_sites, 2] allele indices, uniform over each site's alleles (reference twice as likely)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_samples, len(sites), 2), dtype=np.int8)
    for i, s in enumerate(sites):
        choices = [0] + list(range(s.n_alleles))
        out[:, i, :] = rng.choice(choices, size=(n_samples, 2))
    return out


def make_genotypes(n_sites: int, n_samples: int, seed: int = 13) -> np.ndarray:
    """[n_samples, n_sites, 2] allele indices; uniform over {0/0, 0/1, 0/1, 1/1}, hets randomly phased."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=(n_samples, n_sites))
    phase = rng.integers(0, 2, size=(n_samples, n_sites))
    out = np.zeros((n_samples, n_sites, 2), dtype=np.int8)
    het = (g == 1) | (g == 2)
    out[..., 0] = np.where(het, phase, (g == 3))
    out[..., 1] = np.where(het, 1 - phase, (g == 3))
    return out


def build_haplotype(ref: np.ndarray, sites: List[Site], alleles: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (hap_seq uint8, hap_refpos int64) ; hap_refpos = 0-based ref position or -1 for inserted bases."""
    pieces = []
    pos_pieces = []
    cur = 0  # 0-based ref cursor
    for s, a in zip(sites, alleles.tolist()):
        if not a:
            continue
        p0 = s.pos - 1
        if p0 < cur:
            continue  # overlapped by an allele already applied (complex fixtures: a deletion spanning a later record)
        pieces.append(ref[cur:p0])
        pos_pieces.append(np.arange(cur, p0, dtype=np.int64))
        alt = np.frombuffer(s.allele(int(a)), dtype=np.uint8)
        pieces.append(alt)
        pp = np.full(len(alt), -1, dtype=np.int64)
        n_al = min(len(s.ref), len(alt))  # anchor (and SNP base) stay aligned
        pp[:n_al] = np.arange(p0, p0 + n_al)
        pos_pieces.append(pp)
        cur = p0 + len(s.ref)
    pieces.append(ref[cur:])
    pos_pieces.append(np.arange(cur, len(ref), dtype=np.int64))
    return np.concatenate(pieces), np.concatenate(pos_pieces)


def _cigar_for(refpos: np.ndarray) -> Tuple[int, str]:
    """CIGAR + leftmost ref position for one read given per-base ref positions (-1 = inserted)."""
    n = len(refpos)
    aligned = refpos >= 0
    if not aligned.any():
        return -1, f"{n}S"
    first = int(np.argmax(aligned))
    last = n - 1 - int(np.argmax(aligned[::-1]))
    ops: List[Tuple[str, int]] = []
    if first > 0:
        ops.append(("S", first))
    i = first
    prev_ref = None
    while i <= last:
        if refpos[i] >= 0:
            j = i
            while j + 1 <= last and refpos[j + 1] == refpos[j] + 1:
                j += 1
            if prev_ref is not None and refpos[i] > prev_ref + 1:
                ops.append(("D", int(refpos[i] - prev_ref - 1)))
            ops.append(("M", j - i + 1))
            prev_ref = int(refpos[j])
            i = j + 1
        else:
            j = i
            while j + 1 <= last and refpos[j + 1] < 0:
                j += 1
            ops.append(("I", j - i + 1))
            i = j + 1
    if last < n - 1:
        ops.append(("S", n - 1 - last))
    # merge adjacent equal ops (M D M stays; M M cannot happen)
    return int(refpos[first]), "".join(f"{c}{o}" for o, c in ops)


def simulate_reads(ref: np.ndarray, sites: List[Site], gt: np.ndarray, sample: str, seed: int,
                   coverage: float = 30.0, read_len: int = 150, frag_lo: int = 320, frag_hi: int = 480,
                   err: float = 0.002, n_rate: float = 0.0, name_prefix: Optional[str] = None,
                   lowmapq_rate: float = 0.0, unpaired_rate: float = 0.0, improper_rate: float = 0.0,
                   flip_rate: float = 0.0) -> ReadSet:
    """gt: [n_sites, 2] allele indices for this sample."""
    rng = np.random.default_rng(seed)
    haps = [build_haplotype(ref, sites, gt[:, h]) for h in range(2)]
    n_pairs = int(len(ref) * coverage / (2 * read_len))
    hap_idx = rng.integers(0, 2, size=n_pairs)
    frag_len = rng.integers(frag_lo, frag_hi + 1, size=n_pairs)
    u = rng.random(n_pairs)
    rev = rng.integers(0, 2, size=n_pairs).astype(bool)  # fragment strand

    L = read_len
    seqs = np.empty((2 * n_pairs, L), dtype=np.uint8)
    pos = np.empty(2 * n_pairs, dtype=np.int64)
    cig: List[Optional[str]] = [None] * (2 * n_pairs)
    nerr = np.zeros(2 * n_pairs, dtype=np.int32)
    flag = np.empty(2 * n_pairs, dtype=np.uint16)
    offs = np.arange(L)
    for h in range(2):
        hs, hp = haps[h]
        sel = np.nonzero(hap_idx == h)[0]
        fl = frag_len[sel]
        st = (u[sel] * (len(hs) - fl)).astype(np.int64)
        # left read covers [st, st+L), right read covers [st+fl-L, st+fl)
        for side in range(2):
            a = st if side == 0 else st + fl - L
            idx = a[:, None] + offs[None, :]
            s = hs[idx]
            rp = hp[idx]
            row = 2 * sel + side  # row 2k = left read, 2k+1 = right read of pair k
            # substitution errors
            e = rng.random(s.shape) < err
            shift = rng.integers(1, 4, size=s.shape)
            code = np.searchsorted(BASES, s)  # A0 C1 G2 T3 (BASES sorted ascending)
            s = np.where(e, BASES[(code + shift) % 4], s)
            if n_rate > 0:
                nm = rng.random(s.shape) < n_rate
                s = np.where(nm, np.uint8(ord("N")), s)
            seqs[row] = s
            nerr[row] = e.sum(axis=1)
            simple = (rp[:, 0] >= 0) & (rp[:, -1] - rp[:, 0] == L - 1) & (rp.min(axis=1) >= 0)
            pos[row] = np.where(simple, rp[:, 0], -1)
            for k in np.nonzero(~simple)[0].tolist():
                p, c = _cigar_for(rp[k])
                pos[row[k]] = p
                cig[row[k]] = c
            for k in np.nonzero(simple)[0].tolist():
                cig[row[k]] = f"{L}M"
    # flags: fragment forward: left read = R1 fwd (99), right = R2 rev (147);
    #        fragment reverse: left read = R2 fwd (163), right = R1 rev (83)
    left = np.arange(n_pairs) * 2
    right = left + 1
    flag[left] = np.where(rev, 163, 99)
    flag[right] = np.where(rev, 83, 147)
    mpos = np.empty_like(pos)
    mpos[left] = pos[right]
    mpos[right] = pos[left]
    isz = np.empty_like(pos)
    # TLEN: leftmost to rightmost aligned base; use fragment length approximation via positions
    span = pos[right] + L - pos[left]
    isz[left] = span
    isz[right] = -span
    mapq = np.full(2 * n_pairs, 60, dtype=np.uint8)
    if lowmapq_rate > 0:
        lm = rng.random(n_pairs) < lowmapq_rate
        mapq[left[lm]] = 10
        mapq[right[lm]] = 10
    # stress options: unpaired singles, improper pairs (both orientations get aligned), strand-flipped SEQ
    if unpaired_rate > 0:
        up = rng.random(n_pairs) < unpaired_rate
        flag[left[up]] = 0
        flag[right[up]] = 16
    if improper_rate > 0:
        ip = (rng.random(n_pairs) < improper_rate) & ((flag[left] & 1) != 0)
        first_left = (flag[left] & 64) != 0
        flag[left[ip]] = np.where(first_left[ip], 65, 129)
        flag[right[ip]] = np.where(first_left[ip], 129, 65)
        if flip_rate > 0:
            fl = ip & (rng.random(n_pairs) < flip_rate)
            rows = np.concatenate([left[fl], right[fl]])
            seqs[rows] = _COMP[seqs[rows][:, ::-1]]
    as_tag = (L - 5 * nerr).astype(np.int32)
    xs_tag = np.where(rng.random(2 * n_pairs) < 0.3, rng.integers(20, 100, size=2 * n_pairs), -1).astype(np.int32)
    name_id = np.repeat(np.arange(n_pairs, dtype=np.int64), 2)
    if unpaired_rate > 0:
        name_id[right[up]] += n_pairs  # singles get their own names
    # drop pairs where either mate has no aligned base (cannot happen with <=6 bp insertions) and sort
    order = np.lexsort((np.arange(2 * n_pairs), pos))
    prefix = name_prefix if name_prefix is not None else f"{sample}_r"
    rs = ReadSet(name_id[order], flag[order], pos[order], mapq[order], mpos[order], isz[order],
                 seqs[order], [cig[i] for i in order], as_tag[order], xs_tag[order], sample, prefix)
    return rs


# ----------------------------------------------------------------------------- file writers

def write_fasta(path: str, ref: np.ndarray, contig: str = "chr1", width: int = 60) -> None:
    L = len(ref)
    with open(path, "wb") as f:
        f.write(f">{contig}\n".encode())
        full = (L // width) * width
        if full:
            body = ref[:full].reshape(-1, width)
            nl = np.full((body.shape[0], 1), 10, dtype=np.uint8)
            f.write(np.concatenate([body, nl], axis=1).tobytes())
        if L > full:
            f.write(ref[full:].tobytes() + b"\n")
    # .fai
    with open(path + ".fai", "w") as f:
        f.write(f"{contig}\t{L}\t{len(contig) + 2}\t{width}\t{width + 1}\n")


def write_vcf(path: str, sites: List[Site], contig: str = "chr1", contig_len: int = 0,
              samples: Optional[List[str]] = None, gts: Optional[np.ndarray] = None) -> None:
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n")
        f.write(f"##contig=<ID={contig},length={contig_len}>\n")
        if samples:
            f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        hdr = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"
        if samples:
            hdr += "\tFORMAT\t" + "\t".join(samples)
        f.write(hdr + "\n")
        for i, s in enumerate(sites):
            alts = ",".join(a.decode() for a in (s.alt,) + tuple(s.more_alts))
            line = f"{contig}\t{s.pos}\t.\t{s.ref.decode()}\t{alts}\t.\t.\t{s.info}"
            if samples:
                line += "\tGT\t" + "\t".join(f"{int(gts[k, i, 0])}/{int(gts[k, i, 1])}" for k in range(len(samples)))
            f.write(line + "\n")


def sam_header(contig: str, contig_len: int, sample: str) -> str:
    return (f"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{contig}\tLN:{contig_len}\n"
            f"@RG\tID:{sample}\tSM:{sample}\tPL:ILLUMINA\n")


def write_sam(path: str, rs: ReadSet, contig: str, contig_len: int) -> None:
    n = len(rs)
    L = rs.seq.shape[1] if n else 0
    qual = "I" * L
    seq_strs = rs.seq.tobytes().decode("ascii")
    with open(path, "w") as f:
        f.write(sam_header(contig, contig_len, rs.sample))
        out = []
        for i in range(n):
            tags = f"RG:Z:{rs.sample}\tAS:i:{int(rs.as_tag[i])}"
            if rs.xs_tag[i] >= 0:
                tags += f"\tXS:i:{int(rs.xs_tag[i])}"
            out.append(
                f"{rs.name_prefix}{int(rs.name_id[i])}\t{int(rs.flag[i])}\t{contig}\t{int(rs.pos[i]) + 1}\t"
                f"{int(rs.mapq[i])}\t{rs.cigar[i]}\t=\t{int(rs.mpos[i]) + 1}\t{int(rs.isize[i])}\t"
                f"{seq_strs[i * L:(i + 1) * L]}\t{qual}\t{tags}\n")
        f.write("".join(out))


# ----------------------------------------------------------------------------- dataset

@dataclasses.dataclass
class Dataset:
    ref: np.ndarray
    sites: List[Site]
    gts: np.ndarray  # [n_samples, n_sites, 2]
    samples: List[str]
    reads: List[ReadSet]
    contig: str = "chr1"


def make_dataset(length: int, n_sites: int, n_samples: int = 1, seed: int = 11, coverage: float = 30.0,
                 err: float = 0.002, n_rate: float = 0.0, lowmapq_rate: float = 0.0, unpaired_rate: float = 0.0,
                 improper_rate: float = 0.0, flip_rate: float = 0.0, read_len: int = 150,
                 complex_sites: bool = False) -> Dataset:
    ref = make_reference(length, seed)
    if complex_sites:
        sites = make_sites_complex(ref, n_sites, seed + 1)
        gts = make_genotypes_complex(sites, n_samples, seed + 2)
    else:
        sites = make_sites(ref, n_sites, seed + 1)
        gts = make_genotypes(n_sites, n_samples, seed + 2)
    samples = [f"SAMP{k + 1}" for k in range(n_samples)]
    reads = [simulate_reads(ref, sites, gts[k], samples[k], seed + 100 + k, coverage=coverage, err=err,
                            n_rate=n_rate, lowmapq_rate=lowmapq_rate, unpaired_rate=unpaired_rate,
                            improper_rate=improper_rate, flip_rate=flip_rate, read_len=read_len)
             for k in range(n_samples)]
    return Dataset(ref, sites, gts, samples, reads)


def split_regions(length: int, region_size: int = 50000) -> List[Tuple[int, int]]:
    """1-based inclusive [begin, end] regions exactly as `graphtyper genotype` chops them."""
    out = []
    b = 1
    while b <= length:
        e = min(b + region_size - 1, length)
        out.append((b, e))
        b = e + 1
    return out


def reads_for_region(rs: ReadSet, begin1: int, end1: int) -> np.ndarray:
    """Indices of records of pairs whose leftmost mate starts inside [begin1, end1] (each pair in one region)."""
    left = np.minimum(rs.pos, rs.mpos) + 1
    return np.nonzero((left >= begin1) & (left <= end1))[0]


def write_dataset(ds: Dataset, out_dir: str, region_size: int = 50000) -> dict:
    """Writes ref.fa(+.fai), sites.vcf and per-sample per-region SAM files. Returns a manifest."""
    os.makedirs(out_dir, exist_ok=True)
    L = len(ds.ref)
    fa = os.path.join(out_dir, "ref.fa")
    write_fasta(fa, ds.ref, ds.contig)
    vcf = os.path.join(out_dir, "sites.vcf")
    write_vcf(vcf, ds.sites, ds.contig, L)
    regions = split_regions(L, region_size)
    man = {"fasta": fa, "vcf": vcf, "contig": ds.contig, "length": L, "regions": []}
    for (b, e) in regions:
        sams = []
        for rs in ds.reads:
            idx = reads_for_region(rs, b, e)
            p = os.path.join(out_dir, f"{rs.sample}.{b}-{e}.sam")
            write_sam(p, rs.subset(idx), ds.contig, L)
            sams.append(p)
        man["regions"].append({"begin": b, "end": e, "sams": sams})
    return man


# ----------------------------------------------------------------------------- structural variants (genotype_sv)

def make_sv_sites(ref: np.ndarray, n_sites: int, seed: int = 52, min_size: int = 50, max_size: int = 1000,
                  margin: int = 1500, spacing: int = 1500, p_del: float = 0.6, p_ins: float = 0.25) -> List[Site]:
    """Non-overlapping <DEL> / <INS> / <DUP> sites. `Site.ref`/`Site.alt` hold the resolved sequences (used to build
    the sample haplotypes); write_sv_vcf() writes the symbolic records graphtyper's SV constructor parses
    (INFO keys of src/graph/constructor.cpp:1301-1313)."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    n_slots = (L - 2 * margin) // (max_size + spacing)
    if n_sites > n_slots:
        raise ValueError("too many SV sites")
    slots = np.sort(rng.choice(n_slots, size=n_sites, replace=False))
    sites: List[Site] = []
    for sl in slots.tolist():
        p = margin + sl * (max_size + spacing) + int(rng.integers(0, spacing // 2))
        size = int(rng.integers(min_size, max_size + 1))
        anchor = bytes(ref[p - 1:p])
        k = rng.random()
        if k < p_del:
            s = Site(p, bytes(ref[p - 1:p + size]), anchor)
            s.sv = ("DEL", size, None)
        elif k < p_del + p_ins:
            ins = bytes(BASES[rng.integers(0, 4, size=size)])
            s = Site(p, anchor, anchor + ins)
            s.sv = ("INS", size, ins)
        else:  # tandem duplication of the `size` bases that follow the anchor
            dup = bytes(ref[p:p + size])
            s = Site(p, anchor, anchor + dup)
            s.sv = ("DUP", size, None)
        sites.append(s)
    return sites


def write_sv_vcf(path: str, sites: List[Site], contig: str = "chr1", contig_len: int = 0) -> None:
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n")
        f.write(f"##contig=<ID={contig},length={contig_len}>\n")
        for k, d in (("SVTYPE", "String"), ("SVSIZE", "Integer"), ("SVLEN", "Integer"), ("END", "Integer"),
                     ("SEQ", "String")):
            f.write(f'##INFO=<ID={k},Number=1,Type={d},Description="{k}">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for s in sites:
            typ, size, seq = s.sv
            ref_base = s.ref[:1].decode()
            if typ == "DEL":
                info = f"SVTYPE=DEL;SVSIZE={size};SVLEN=-{size};END={s.pos + size}"
            elif typ == "INS":
                info = f"SVTYPE=INS;SVSIZE={size};SVLEN={size};END={s.pos};SEQ={seq.decode()}"
            else:
                info = f"SVTYPE=DUP;SVSIZE={size};SVLEN={size};END={s.pos + size}"
            f.write(f"{contig}\t{s.pos}\t.\t{ref_base}\t<{typ}>\t.\t.\t{info}\n")


# ---------------------------------------------------------------------------------------------------------------------
# (read, haplotype window) pairs for the discovery re-alignment kernel (gtb_sw_align_batch)
_ACGT = np.frombuffer(b"ACGT", np.uint8)


def make_sw_pairs(n_pairs: int, seed: int = 1, min_db: int = 60, max_db: int = 520, max_query: int = 151,
                  lowercase_rate: float = 0.03, min_query: int = 20) -> Tuple[List[bytes], List[bytes]]:
    """Seeded read/window pairs shaped like realign_to_indels' input (src/typer/caller.cpp:1930-2007: a read against
    reference window +- 50..100 bp with one indel haplotype spliced in): the read is a fragment of the window with
    substitutions, small indels, occasional N, overhangs past either window end (exercises clipping) and random
    tails; windows sometimes hold N runs or lower-case bases."""
    rng = np.random.default_rng(seed)

    def mutate(s: bytes) -> bytes:
        out = bytearray()
        psub = rng.choice([0.0, 0.01, 0.03, 0.1])
        pins = rng.choice([0, 0.003, 0.02])
        pdel = rng.choice([0, 0.003, 0.02])
        i = 0
        while i < len(s):
            r = rng.random()
            if r < pdel:
                i += int(rng.integers(1, 8))
                continue
            if r < pdel + pins:
                out += bytes(_ACGT[rng.integers(0, 4, size=int(rng.integers(1, 8)))])
            c = s[i]
            if rng.random() < psub:
                c = int(_ACGT[rng.integers(0, 4)])
            if rng.random() < 0.002:
                c = ord("N")
            out.append(c)
            i += 1
        return bytes(out)

    queries: List[bytes] = []
    windows: List[bytes] = []
    for _ in range(n_pairs):
        n = int(rng.integers(min_db, max_db))
        d = bytes(_ACGT[rng.integers(0, 4, size=n)])
        length = int(rng.integers(min_query, max_query + 1))
        st = int(rng.integers(-30, n - 10))
        frag = d[max(st, 0):max(st, 0) + length]
        if st < 0:
            frag = bytes(_ACGT[rng.integers(0, 4, size=-st)]) + frag
        if rng.random() < 0.2:
            frag = frag + bytes(_ACGT[rng.integers(0, 4, size=int(rng.integers(1, 40)))])
        q = mutate(frag)[:max_query]
        if len(q) < 5:
            q = frag[:20] + b"ACGTA"
        if rng.random() < 0.05:
            d = d[:n // 2] + b"N" * int(rng.integers(1, 5)) + d[n // 2:]
        if rng.random() < lowercase_rate:
            a = int(rng.integers(0, len(d)))
            d = d[:a] + d[a:a + 20].lower() + d[a + 20:]
        queries.append(q)
        windows.append(d)
    return queries, windows

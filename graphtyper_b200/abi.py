"""ctypes mirror of include/gtb200.h (the C-ABI structs) plus array marshalling helpers.

Used by the Python harness (tests, bench, smoke) to call either the product library
(graphtyper_b200/csrc -> libgtb200.so) or, in tests only, the oracle library
(oracle/_build/libgtb_oracle.so).  Both take exactly the same plain-pointer structs.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

SEQ_STRIDE = 76
INVALID_ID = 0xFFFFFFFF
SPECIAL_START = 0xD0000000

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class GraphView(C.Structure):
    _fields_ = [
        ("n_ref", C.c_uint32), ("n_var", C.c_uint32), ("is_sv_graph", C.c_int32), ("reserved", C.c_int32),
        ("ref_order", u32p), ("ref_seq_off", u64p), ("ref_var_off", u32p),
        ("var_order", u32p), ("var_seq_off", u64p), ("var_out_ref", u32p),
        ("seq", u8p), ("seq_len", C.c_uint64),
        ("var_ev_off", u32p), ("var_ev", i64p), ("var_aev_off", u32p), ("var_aev", i64p),
        ("n_special", C.c_uint32), ("actual_poses", u32p), ("ref_reach_poses", u32p),
        ("n_sp_keys", C.c_uint32), ("sp_keys", u32p), ("sp_off", u32p), ("sp_list", u32p),
    ]


class Label(C.Structure):
    _fields_ = [("start", C.c_uint32), ("end", C.c_uint32), ("var_id", C.c_uint32)]


class ReadBatch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("seq_stride", C.c_uint32),
        ("seq4", u8p), ("lseq", u16p), ("flag", u16p), ("mapq", u8p), ("isize", i32p),
        ("same_tid", u8p), ("score_diff", u8p), ("clipped", u8p), ("sample", i32p),
        ("mate", i32p), ("dup_of", i32p), ("leftover", u8p),
    ]


class Accumulators(C.Structure):
    _fields_ = [
        ("n_bubbles", C.c_uint32), ("n_samples", C.c_uint32),
        ("bubble_id", u32p), ("n_alleles", u32p), ("score_off", u64p), ("cov_off", u64p),
        ("log_score", u16p), ("gt_coverage", u16p), ("max_log_score", u16p),
        ("ambiguous_depth", u8p), ("ambiguous_depth_alt", u8p), ("alt_proper_pair_depth", u8p),
        ("saturated", u32p),
        ("vs_clipped_reads", u64p), ("vs_mapq_squared", u64p),
        ("pa_clipped_bp", u64p), ("pa_mapq_squared", u64p), ("pa_score_diff", u64p), ("pa_mismatches", u64p),
        ("read_strand", u32p),
        ("depth_size", C.c_uint32), ("reference_offset", C.c_uint32), ("ref_depth", u16p),
    ]


class SubmitStats(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_alignments", C.c_uint64), ("n_oriented", C.c_uint64),
                ("n_pairs_scored", C.c_uint64), ("n_singles_scored", C.c_uint64),
                ("n_capacity_overflow", C.c_uint64), ("kernel_launches", C.c_uint64)]


class Connection(C.Structure):
    _fields_ = [("sample", C.c_uint32), ("hap1", C.c_uint16), ("allele1", C.c_uint16), ("hap2", C.c_uint16),
                ("allele2", C.c_uint16), ("count", C.c_uint32)]


class PhaseSupportEntry(C.Structure):
    _fields_ = [("hap1", C.c_uint16), ("allele1", C.c_uint16), ("hap2", C.c_uint16), ("allele2", C.c_uint16),
                ("flags", C.c_int8), ("pad", C.c_uint8 * 7)]


CONNECTION_DTYPE = np.dtype([("sample", np.uint32), ("hap1", np.uint16), ("allele1", np.uint16), ("hap2", np.uint16),
                             ("allele2", np.uint16), ("count", np.uint32)])
PHASE_DTYPE = np.dtype([("hap1", np.uint16), ("allele1", np.uint16), ("hap2", np.uint16), ("allele2", np.uint16),
                        ("flags", np.int8), ("pad", np.uint8, (7,))])
assert CONNECTION_DTYPE.itemsize == C.sizeof(Connection) == 16 and PHASE_DTYPE.itemsize == C.sizeof(PhaseSupportEntry) == 16


def connections_as_table(c: np.ndarray) -> np.ndarray:
    """uint32 [n, 6]: sample hap1 allele1 hap2 allele2 count."""
    return np.stack([c[k].astype(np.uint32) for k in ("sample", "hap1", "allele1", "hap2", "allele2", "count")], axis=1) \
        if len(c) else np.zeros((0, 6), np.uint32)


def phase_as_table(p: np.ndarray) -> np.ndarray:
    """uint32 [n, 5]: hap1 allele1 hap2 allele2 flags."""
    return np.stack([p[k].astype(np.uint32) for k in ("hap1", "allele1", "hap2", "allele2", "flags")], axis=1) \
        if len(p) else np.zeros((0, 5), np.uint32)


def phase_support(lib, fn_name: str, acc, conn: np.ndarray) -> np.ndarray:
    """Calls <lib>.<fn_name>(acc, n_conn, conn, &n, out) with the count-then-fill protocol."""
    fn = getattr(lib, fn_name)
    fn.argtypes = [C.POINTER(Accumulators), C.c_uint64, C.c_void_p, u64p, C.c_void_p]
    conn = np.ascontiguousarray(conn, dtype=CONNECTION_DTYPE)
    n = C.c_uint64(0)
    rc = fn(C.byref(acc.view), len(conn), conn.ctypes.data, C.byref(n), None)
    if rc:
        raise RuntimeError(f"{fn_name} rc={rc}")
    out = np.zeros(n.value, PHASE_DTYPE)
    rc = fn(C.byref(acc.view), len(conn), conn.ctypes.data, C.byref(n), out.ctypes.data)
    if rc:
        raise RuntimeError(f"{fn_name} rc={rc}")
    return out


class BamCore(C.Structure):
    _fields_ = [("pos", C.c_int64), ("mpos", C.c_int64), ("isize", C.c_int64), ("tid", C.c_int32), ("mtid", C.c_int32),
                ("l_qseq", C.c_int32), ("n_cigar", C.c_uint32), ("flag", C.c_uint16), ("l_qname", C.c_uint16),
                ("mapq", C.c_uint8), ("reserved", C.c_uint8 * 3)]


BAM_CORE_DTYPE = np.dtype([("pos", np.int64), ("mpos", np.int64), ("isize", np.int64), ("tid", np.int32), ("mtid", np.int32),
                           ("l_qseq", np.int32), ("n_cigar", np.uint32), ("flag", np.uint16), ("l_qname", np.uint16),
                           ("mapq", np.uint8), ("reserved", np.uint8, (3,))])
assert BAM_CORE_DTYPE.itemsize == C.sizeof(BamCore) == 48


class BamBatch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("reserved", C.c_uint32), ("core", C.c_void_p), ("data", u8p), ("data_off", u64p),
                ("sample", i32p), ("rg", i32p)]


class HostBamBatch:
    """Owns the arrays behind a BamBatch: records as htslib holds them (core fields + bam1_t::data)."""

    def __init__(self, core: np.ndarray, data: np.ndarray, data_off: np.ndarray, sample: np.ndarray, rg: np.ndarray):
        self.core = np.ascontiguousarray(core, dtype=BAM_CORE_DTYPE)
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.data_off = np.ascontiguousarray(data_off, dtype=np.uint64)
        self.sample = np.ascontiguousarray(sample, dtype=np.int32)
        self.rg = np.ascontiguousarray(rg, dtype=np.int32)
        v = BamBatch()
        v.n_reads = len(self.core)
        v.core = self.core.ctypes.data
        v.data = _ptr(self.data, u8p)
        v.data_off = _ptr(self.data_off, u64p)
        v.sample = _ptr(self.sample, i32p)
        v.rg = _ptr(self.rg, i32p)
        self.view = v

    def __len__(self) -> int:
        return len(self.core)

    @classmethod
    def from_probe(cls, d: Dict[str, np.ndarray]) -> "HostBamBatch":
        """From the records dumped by oracle/_ref/bin/gt_probe (<out>.reads.gtba: bam_data / bam_off + core columns)."""
        n = len(d["flag"])
        core = np.zeros(n, BAM_CORE_DTYPE)
        core["pos"], core["mpos"], core["isize"] = d["pos"], d["mpos"], d["isize"]
        core["tid"], core["mtid"], core["l_qseq"] = d["tid"], d["mtid"], d["lseq"]
        core["n_cigar"] = np.diff(d["cigar_off"].astype(np.int64))
        core["flag"], core["l_qname"], core["mapq"] = d["flag"], d["bam_lqname"], d["mapq"]
        return cls(core, d["bam_data"], d["bam_off"], d["sample"], d["rg"])


def _ptr(a: Optional[np.ndarray], typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


class HostGraph:
    """Owns the numpy arrays behind a GraphView (keeps them alive)."""

    FIELDS = ["ref_order", "ref_seq_off", "ref_var_off", "var_order", "var_seq_off", "var_out_ref", "seq",
              "var_ev_off", "var_ev", "var_aev_off", "var_aev", "actual_poses", "ref_reach_poses",
              "sp_keys", "sp_off", "sp_list"]
    DTYPES = {"ref_order": np.uint32, "ref_seq_off": np.uint64, "ref_var_off": np.uint32, "var_order": np.uint32,
              "var_seq_off": np.uint64, "var_out_ref": np.uint32, "seq": np.uint8, "var_ev_off": np.uint32,
              "var_ev": np.int64, "var_aev_off": np.uint32, "var_aev": np.int64, "actual_poses": np.uint32,
              "ref_reach_poses": np.uint32, "sp_keys": np.uint32, "sp_off": np.uint32, "sp_list": np.uint32}

    def __init__(self, arrays: Dict[str, np.ndarray], is_sv_graph: bool = False):
        self.a = {k: np.ascontiguousarray(arrays[k], dtype=self.DTYPES[k]) for k in self.FIELDS}
        self.is_sv_graph = bool(is_sv_graph)
        a = self.a
        v = GraphView()
        v.n_ref = len(a["ref_order"])
        v.n_var = len(a["var_order"])
        v.is_sv_graph = 1 if self.is_sv_graph else 0
        v.ref_order = _ptr(a["ref_order"], u32p)
        v.ref_seq_off = _ptr(a["ref_seq_off"], u64p)
        v.ref_var_off = _ptr(a["ref_var_off"], u32p)
        v.var_order = _ptr(a["var_order"], u32p)
        v.var_seq_off = _ptr(a["var_seq_off"], u64p)
        v.var_out_ref = _ptr(a["var_out_ref"], u32p)
        v.seq = _ptr(a["seq"], u8p)
        v.seq_len = len(a["seq"])
        v.var_ev_off = _ptr(a["var_ev_off"], u32p)
        v.var_ev = _ptr(a["var_ev"], i64p)
        v.var_aev_off = _ptr(a["var_aev_off"], u32p)
        v.var_aev = _ptr(a["var_aev"], i64p)
        v.n_special = len(a["actual_poses"])
        v.actual_poses = _ptr(a["actual_poses"], u32p)
        v.ref_reach_poses = _ptr(a["ref_reach_poses"], u32p)
        v.n_sp_keys = len(a["sp_keys"])
        v.sp_keys = _ptr(a["sp_keys"], u32p)
        v.sp_off = _ptr(a["sp_off"], u32p)
        v.sp_list = _ptr(a["sp_list"], u32p)
        self.view = v

    @classmethod
    def from_gtba(cls, d: Dict[str, np.ndarray]) -> "HostGraph":
        return cls(d, is_sv_graph=bool(d["meta"][0]))

    @property
    def n_bubbles(self) -> int:
        return len(self.a["ref_order"]) - 1


class HostBatch:
    """Owns the arrays behind a ReadBatch."""

    def __init__(self, seq4, lseq, flag, mapq, isize, same_tid, score_diff, clipped, sample, mate, dup_of,
                 seq_stride: int = SEQ_STRIDE, leftover=None):
        n = len(lseq)
        self.seq4 = np.ascontiguousarray(seq4, dtype=np.uint8).reshape(n, seq_stride)
        self.lseq = np.ascontiguousarray(lseq, dtype=np.uint16)
        self.flag = np.ascontiguousarray(flag, dtype=np.uint16)
        self.mapq = np.ascontiguousarray(mapq, dtype=np.uint8)
        self.isize = np.ascontiguousarray(isize, dtype=np.int32)
        self.same_tid = np.ascontiguousarray(same_tid, dtype=np.uint8)
        self.score_diff = np.ascontiguousarray(score_diff, dtype=np.uint8)
        self.clipped = np.ascontiguousarray(clipped, dtype=np.uint8)
        self.sample = np.ascontiguousarray(sample, dtype=np.int32)
        self.mate = np.ascontiguousarray(mate, dtype=np.int32)
        self.dup_of = np.ascontiguousarray(dup_of, dtype=np.int32)
        if leftover is None:
            leftover = compute_leftover(self.flag, self.mate)
        self.leftover = np.ascontiguousarray(leftover, dtype=np.uint8)
        b = ReadBatch()
        b.n_reads = n
        b.seq_stride = seq_stride
        b.seq4 = _ptr(self.seq4, u8p)
        b.lseq = _ptr(self.lseq, u16p)
        b.flag = _ptr(self.flag, u16p)
        b.mapq = _ptr(self.mapq, u8p)
        b.isize = _ptr(self.isize, i32p)
        b.same_tid = _ptr(self.same_tid, u8p)
        b.score_diff = _ptr(self.score_diff, u8p)
        b.clipped = _ptr(self.clipped, u8p)
        b.sample = _ptr(self.sample, i32p)
        b.mate = _ptr(self.mate, i32p)
        b.dup_of = _ptr(self.dup_of, i32p)
        b.leftover = _ptr(self.leftover, u8p)
        self.view = b

    def __len__(self) -> int:
        return len(self.lseq)

    def nbytes_h2d(self) -> int:
        return sum(x.nbytes for x in (self.seq4, self.lseq, self.flag, self.mapq, self.isize, self.same_tid,
                                      self.score_diff, self.clipped, self.sample, self.mate, self.dup_of,
                                      self.leftover))


class HostAccumulators:
    """Caller-owned accumulator buffers sized from (n_bubbles, n_scores, n_cov, n_samples)."""

    def __init__(self, n_bubbles: int, n_scores: int, n_cov: int, n_samples: int, depth_size: int = 0):
        NB, NS = n_bubbles, n_samples
        self.ref_depth = np.zeros(depth_size * NS, np.uint16)
        self.bubble_id = np.zeros(NB, np.uint32)
        self.n_alleles = np.zeros(NB, np.uint32)
        self.score_off = np.zeros(NB + 1, np.uint64)
        self.cov_off = np.zeros(NB + 1, np.uint64)
        self.log_score = np.zeros(n_scores * NS, np.uint16)
        self.gt_coverage = np.zeros(n_cov * NS, np.uint16)
        self.max_log_score = np.zeros(NB * NS, np.uint16)
        self.ambiguous_depth = np.zeros(NB * NS, np.uint8)
        self.ambiguous_depth_alt = np.zeros(NB * NS, np.uint8)
        self.alt_proper_pair_depth = np.zeros(NB * NS, np.uint8)
        self.saturated = np.zeros(NB * NS, np.uint32)
        self.vs_clipped_reads = np.zeros(NB, np.uint64)
        self.vs_mapq_squared = np.zeros(NB, np.uint64)
        self.pa_clipped_bp = np.zeros(n_cov, np.uint64)
        self.pa_mapq_squared = np.zeros(n_cov, np.uint64)
        self.pa_score_diff = np.zeros(n_cov, np.uint64)
        self.pa_mismatches = np.zeros(n_cov, np.uint64)
        self.read_strand = np.zeros(n_cov * 4, np.uint32)
        a = Accumulators()
        a.n_bubbles = NB
        a.n_samples = NS
        for name, typ in Accumulators._fields_[2:]:
            if name in ("depth_size", "reference_offset"):
                continue
            setattr(a, name, _ptr(getattr(self, name), typ))
        a.depth_size = depth_size
        if depth_size == 0:
            a.ref_depth = C.cast(None, u16p)
        self.view = a
        self.n_samples = NS
        self.n_bubbles = NB

    ARRAYS = ["bubble_id", "n_alleles", "score_off", "cov_off", "log_score", "gt_coverage", "max_log_score",
              "ambiguous_depth", "ambiguous_depth_alt", "alt_proper_pair_depth", "saturated",
              "vs_clipped_reads", "vs_mapq_squared", "pa_clipped_bp", "pa_mapq_squared", "pa_score_diff",
              "pa_mismatches", "read_strand"]

    def as_dict(self) -> Dict[str, np.ndarray]:
        d = {k: getattr(self, k) for k in self.ARRAYS}
        if len(self.ref_depth):
            d["ref_depth"] = self.ref_depth
        return d


# ----------------------------------------------------------------------------- read preparation (host)

_NIB = np.full(256, 15, dtype=np.uint8)
for _c, _v in zip(b"=ACMGRSVTWYHKDBN", range(16)):
    _NIB[_c] = _v
    _NIB[ord(chr(_c).lower())] = _v


def pack_seq4(seq_ascii: np.ndarray, seq_stride: int = SEQ_STRIDE) -> np.ndarray:
    """ASCII [n, L] -> BAM 4-bit packed [n, seq_stride] (high nibble first), as sam_parse1 stores it."""
    n, L = seq_ascii.shape
    nib = _NIB[seq_ascii]
    if L % 2:
        nib = np.concatenate([nib, np.zeros((n, 1), np.uint8)], axis=1)
    packed = (nib[:, 0::2] << 4) | nib[:, 1::2]
    out = np.zeros((n, seq_stride), np.uint8)
    out[:, :packed.shape[1]] = packed
    return out


def score_diff_from_tags(as_tag: np.ndarray, xs_tag: np.ndarray) -> np.ndarray:
    """get_score_diff (src/typer/alignment.cpp:140-325): as/xs = -1 when the tag is absent."""
    a = as_tag.astype(np.int64)
    x = xs_tag.astype(np.int64)
    zero = (a == -1) | (a < x)
    x = np.where(x == -1, 0, x)
    d = np.minimum(a - x, 255)
    return np.where(zero, 0, d).astype(np.uint8)


def compute_leftover(flag: np.ndarray, mate: np.ndarray) -> np.ndarray:
    """Paired records still waiting in the read-name map when the pool ends (only used on SV graphs,
    hts_parallel_reader.cpp:719-772): paired, not the second arrival of a pair, never referenced as a mate."""
    n = len(flag)
    has_partner = np.zeros(n, bool)
    m = mate[mate >= 0]
    has_partner[m] = True
    return (((flag & 1) != 0) & (mate < 0) & ~has_partner).astype(np.uint8)


def compute_dup_of(pos: np.ndarray, lseq: np.ndarray, seq4: np.ndarray, tid: Optional[np.ndarray] = None) -> np.ndarray:
    """equal_pos_seq chain (include/graphtyper/utilities/hts_utils.hpp:110-128, hts_parallel_reader.cpp:666-684)."""
    n = len(pos)
    dup = np.full(n, -1, np.int32)
    if n < 2:
        return dup
    eq = (pos[1:] == pos[:-1]) & (lseq[1:] == lseq[:-1])
    if tid is not None:
        eq &= tid[1:] == tid[:-1]
    idx = np.nonzero(eq)[0]
    if len(idx):
        nb = (lseq[idx + 1].astype(np.int64) + 1) // 2
        same = np.ones(len(idx), bool)
        # compare packed bytes over (l_qseq+1)/2 (all reads usually share one length)
        for L in np.unique(nb):
            m = nb == L
            same[m] = (seq4[idx[m] + 1, :L] == seq4[idx[m], :L]).all(axis=1)
        eq[idx] = same
    is_dup = np.concatenate([[False], eq])
    # root = last non-dup record before i
    root = np.where(~is_dup, np.arange(n), 0)
    root = np.maximum.accumulate(root)
    dup[is_dup] = root[is_dup]
    return dup


def compute_mates(name_id: np.ndarray, flag: np.ndarray, rg: Optional[np.ndarray] = None) -> np.ndarray:
    """Read-name map semantics of genotype_only (hts_parallel_reader.cpp:270-337): within a read group the
    1st paired record of a name waits, the 2nd pairs with it (and removes it), the 3rd waits again ..."""
    n = len(name_id)
    mate = np.full(n, -1, np.int32)
    paired = (flag & 1) != 0
    idx = np.nonzero(paired)[0]
    if len(idx) == 0:
        return mate
    key = name_id[idx].astype(np.int64)
    if rg is not None:
        key = key * (int(rg.max()) + 1) + rg[idx]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    # rank within each run of equal keys
    start = np.concatenate([[True], ks[1:] != ks[:-1]])
    run_start = np.maximum.accumulate(np.where(start, np.arange(len(ks)), 0))
    rank = np.arange(len(ks)) - run_start
    second = (rank % 2) == 1
    sec_pos = np.nonzero(second)[0]
    mate[idx[order[sec_pos]]] = idx[order[sec_pos - 1]]
    return mate


def batch_from_readsets(readsets: Sequence, region_idx: Optional[Sequence[np.ndarray]] = None,
                        flag_filter: int = 3840) -> HostBatch:
    """Builds one pool's batch from synth.ReadSet objects (one per sample), merged in the reference's k-way
    merge order: (pos, then the 4-bit sequence bytes), ties keeping file order
    (Cmp_gt_pair_bam1_t_fun, include/graphtyper/utilities/hts_utils.hpp:60-108)."""
    cols = {k: [] for k in ("pos", "flag", "mapq", "isize", "seq", "as", "xs", "sample", "name")}
    for s, rs in enumerate(readsets):
        idx = np.arange(len(rs)) if region_idx is None else region_idx[s]
        keep = (rs.flag[idx] & flag_filter) == 0
        idx = idx[keep]
        cols["pos"].append(rs.pos[idx])
        cols["flag"].append(rs.flag[idx])
        cols["mapq"].append(rs.mapq[idx])
        cols["isize"].append(rs.isize[idx])
        cols["seq"].append(rs.seq[idx])
        cols["as"].append(rs.as_tag[idx])
        cols["xs"].append(rs.xs_tag[idx])
        cols["sample"].append(np.full(len(idx), s, np.int32))
        # names are unique per sample: offset ids per sample so the name map never collides across RGs
        cols["name"].append(rs.name_id[idx].astype(np.int64) * len(readsets) + s)
    cat = {k: np.concatenate(v) for k, v in cols.items()}
    n = len(cat["pos"])
    L = cat["seq"].shape[1] if n else 0
    seq4 = pack_seq4(cat["seq"]) if n else np.zeros((0, SEQ_STRIDE), np.uint8)
    if n:
        order = merge_order(cat["pos"], np.full(n, L), seq4, cat["sample"])
        for k in cat:
            cat[k] = cat[k][order]
        seq4 = seq4[order]
    lseq = np.full(n, L, np.uint16)
    dup = compute_dup_of(cat["pos"], lseq, seq4)
    mate = compute_mates(cat["name"], cat["flag"])
    sd = score_diff_from_tags(cat["as"], cat["xs"])
    return HostBatch(seq4, lseq, cat["flag"], cat["mapq"], np.clip(cat["isize"], -2**31, 2**31 - 1),
                     np.ones(n, np.uint8), sd, np.zeros(n, np.uint8), cat["sample"], mate, dup)


def bam_batch_from_readsets(readsets: Sequence, region_idx: Optional[Sequence[np.ndarray]] = None,
                            flag_filter: int = 3840) -> "HostBamBatch":
    """The same pool as batch_from_readsets (same records, same merge order), but as raw records the way htslib holds them
    after reading the SAM/BAM written by synth.write_sam: qname (NUL + padding to 4) | cigar | 4-bit seq | qual | aux
    (RG:Z, AS, XS with the smallest integer type that fits, like sam_parse1)."""
    import re
    cols = {k: [] for k in ("pos", "mpos", "flag", "mapq", "isize", "seq", "as", "xs", "sample", "name", "cig")}
    names: List[bytes] = []
    for s, rs in enumerate(readsets):
        idx = np.arange(len(rs)) if region_idx is None else region_idx[s]
        idx = idx[(rs.flag[idx] & flag_filter) == 0]
        for k, v in (("pos", rs.pos), ("mpos", rs.mpos), ("flag", rs.flag), ("mapq", rs.mapq), ("isize", rs.isize), ("seq", rs.seq),
                     ("as", rs.as_tag), ("xs", rs.xs_tag)):
            cols[k].append(v[idx])
        cols["sample"].append(np.full(len(idx), s, np.int32))
        cols["name"].append(np.arange(len(names), len(names) + len(idx)))
        cols["cig"].append(np.arange(len(names), len(names) + len(idx)))
        names += [(f"{rs.name_prefix}{int(rs.name_id[i])}".encode(), rs.cigar[i], rs.sample.encode()) for i in idx]
    cat = {k: np.concatenate(v) for k, v in cols.items()}
    n = len(cat["pos"])
    L = cat["seq"].shape[1] if n else 0
    seq4 = pack_seq4(cat["seq"]) if n else np.zeros((0, SEQ_STRIDE), np.uint8)
    order = merge_order(cat["pos"], np.full(n, L), seq4, cat["sample"]) if n else np.zeros(0, np.int64)
    ops = {c: i for i, c in enumerate("MIDNSHP=X")}

    def int_tag(tag: bytes, v: int) -> bytes:
        if 0 <= v < 256:
            return tag + b"C" + int(v).to_bytes(1, "little")
        if -128 <= v < 0:
            return tag + b"c" + int(v).to_bytes(1, "little", signed=True)
        if 0 <= v < 65536:
            return tag + b"S" + int(v).to_bytes(2, "little")
        return tag + b"i" + int(v).to_bytes(4, "little", signed=True)

    core = np.zeros(n, BAM_CORE_DTYPE)
    off = np.zeros(n + 1, np.uint64)
    parts: List[bytes] = []
    nb = (L + 1) // 2
    for j, o in enumerate(order):
        qn, cigar, sample = names[cat["name"][o]]
        q = qn + b"\0"
        q += b"\0" * ((4 - len(q) % 4) % 4)
        cg = b"".join(((int(m) << 4) | ops[c]).to_bytes(4, "little") for m, c in re.findall(r"(\d+)([MIDNSHP=X])", cigar))
        aux = b"RGZ" + sample + b"\0" + int_tag(b"AS", int(cat["as"][o]))
        if cat["xs"][o] >= 0:
            aux += int_tag(b"XS", int(cat["xs"][o]))
        rec = q + cg + seq4[o, :nb].tobytes() + b"\x28" * L + aux
        parts.append(rec)
        off[j + 1] = off[j] + len(rec)
        core["l_qname"][j], core["n_cigar"][j] = len(q), len(cg) // 4
    for k, src in (("pos", "pos"), ("mpos", "mpos"), ("isize", "isize"), ("flag", "flag"), ("mapq", "mapq")):
        core[k] = cat[src][order]
    core["l_qseq"] = L
    return HostBamBatch(core, np.frombuffer(b"".join(parts), np.uint8), off, cat["sample"][order], np.zeros(n, np.int32)
                        if len(readsets) == 1 else cat["sample"][order])


def merge_order(pos: np.ndarray, lseq: np.ndarray, seq4: np.ndarray, file_index: np.ndarray) -> np.ndarray:
    """Order of HtsParallelReader's heap merge: ascending (tid,) pos, l_qseq, sequence bytes; equal records keep
    heap order.  Within one file same-position records are std::sort-ed descending and popped from the back
    (HtsReader::get_next_read_in_order, src/utilities/hts_reader.cpp:166-303), so fully equal records come out
    in reverse file order; across files the tie order is the heap's (approximated by file index)."""
    n = len(pos)
    nb = seq4.shape[1]
    keys = [-np.arange(n), file_index]
    for c in range(nb - 1, -1, -1):
        keys.append(seq4[:, c])
    keys.append(lseq)
    keys.append(pos)
    return np.lexsort(tuple(keys))


def batch_from_probe(d: Dict[str, np.ndarray]) -> HostBatch:
    """Batch from the record columns dumped by oracle/_ref/bin/gt_probe (<out>.reads.gtba)."""
    n = len(d["flag"])
    seq4 = np.zeros((n, SEQ_STRIDE), np.uint8)
    off = d["seq_off"].astype(np.int64)
    ln = (off[1:] - off[:-1])
    for L in np.unique(ln):
        m = np.nonzero(ln == L)[0]
        gather = off[m][:, None] + np.arange(L)[None, :]
        seq4[m[:, None], np.arange(L)[None, :]] = d["seq4"][gather]
    names = d["names"].tobytes()
    noff = d["name_off"].astype(np.int64)
    ids: Dict[bytes, int] = {}
    name_id = np.empty(n, np.int64)
    for i in range(n):
        nm = names[noff[i]:noff[i + 1]]
        name_id[i] = ids.setdefault(nm, len(ids))
    lseq = d["lseq"].astype(np.uint16)
    dup = compute_dup_of(d["pos"], lseq, seq4, d["tid"])
    assert np.array_equal(dup >= 0, d["isdup"] != 0), "dup detection differs from the reference loop"
    mate = compute_mates(name_id, d["flag"], d["rg"].astype(np.int64))
    return HostBatch(seq4, lseq, d["flag"], d["mapq"], np.clip(d["isize"], -2**31, 2**31 - 1),
                     (d["tid"] == d["mtid"]).astype(np.uint8), d["score_diff"], np.zeros(n, np.uint8),
                     d["sample"], mate, dup)


def shard_batch(batch: HostBatch, n_shards: int) -> List[HostBatch]:
    """Splits one pool's records over `n_shards` GPUs for the "one sample on N GPUs" mode.
    Records connected by a mate link or a duplicate link stay together (union-find), so every shard reproduces
    exactly the alignments and pairings of the unsharded stream; accumulators are additive, so the shards' widened
    accumulators sum (NCCL all-reduce) to the unsharded result."""
    n = len(batch)
    parent = np.arange(n)

    def find(x: int) -> int:
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i in range(n):
        for j in (int(batch.mate[i]), int(batch.dup_of[i])):
            if j >= 0:
                a, b = find(i), find(j)
                if a != b:
                    parent[max(a, b)] = min(a, b)
    roots = np.array([find(i) for i in range(n)])
    uniq, comp = np.unique(roots, return_inverse=True)
    shard_of = comp % n_shards
    out = []
    for s in range(n_shards):
        idx = np.nonzero(shard_of == s)[0]
        remap = np.full(n, -1, np.int64)
        remap[idx] = np.arange(len(idx))
        mate = np.where(batch.mate[idx] >= 0, remap[np.maximum(batch.mate[idx], 0)], -1)
        dup = np.where(batch.dup_of[idx] >= 0, remap[np.maximum(batch.dup_of[idx], 0)], -1)
        out.append(HostBatch(batch.seq4[idx], batch.lseq[idx], batch.flag[idx], batch.mapq[idx], batch.isize[idx],
                             batch.same_tid[idx], batch.score_diff[idx], batch.clipped[idx], batch.sample[idx],
                             mate, dup))
    return out

"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's genotyping hot path (oracle/gtb_oracle.cpp) and the recipe that
compiles the unmodified reference itself into oracle/_ref/ (oracle/ref_build/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (graphtyper_b200/) never does; it fails loudly when its CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libgtb_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "bin")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "gtb_oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "gtb200.h")
    if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, src]
    subprocess.run(cmd, check=True)
    return LIB


PAW_LIB = os.path.join(BUILD, "libpaw_oracle.so")


def build_paw_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "paw_oracle.cpp")
    if not force and os.path.exists(PAW_LIB) and os.path.getmtime(PAW_LIB) >= os.path.getmtime(src):
        return PAW_LIB
    os.makedirs(BUILD, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", PAW_LIB, src], check=True)
    return PAW_LIB


def build_reference(jobs: int = 8) -> Optional[str]:
    """Compiles the reference into oracle/_ref (only possible where /root/reference exists)."""
    if not os.path.isdir("/root/reference/src"):
        return None
    subprocess.run(["make", "-C", os.path.join(HERE, "ref_build"), f"-j{jobs}"], check=True,
                   stdout=subprocess.DEVNULL)
    return REF_BIN


def binned_pl() -> Optional[np.ndarray]:
    """The reference's PL / GQ output binning table (vcf.cpp:1107-1113), dumped by the reference build."""
    p = os.path.join(HERE, "_ref", "gen", "binned_pl.txt")
    if not os.path.exists(p):
        return None
    return np.loadtxt(p, dtype=np.int64)


def ref_binary(name: str) -> Optional[str]:
    p = os.path.join(REF_BIN, name)
    return p if os.path.exists(p) else None


class Oracle:
    """Thin ctypes wrapper over libgtb_oracle.so with numpy in/out."""

    def __init__(self):
        from graphtyper_b200 import abi
        self.abi = abi
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.gto_last_error.restype = C.c_char_p
        L.gto_index_build.argtypes = [C.POINTER(abi.GraphView), C.POINTER(C.c_void_p)]
        L.gto_index_size.argtypes = [C.c_void_p, abi.u64p, abi.u64p]
        L.gto_index_export.argtypes = [C.c_void_p, abi.u64p, abi.u32p, C.POINTER(abi.Label)]
        L.gto_index_free.argtypes = [C.c_void_p]
        L.gto_pool_run.argtypes = [C.POINTER(abi.GraphView), C.c_void_p, C.c_int, C.POINTER(abi.ReadBatch), C.c_int,
                                   C.POINTER(C.c_void_p)]
        L.gto_result_free.argtypes = [C.c_void_p]
        L.gto_result_stats.argtypes = [C.c_void_p, C.POINTER(abi.SubmitStats)]
        L.gto_result_accum_sizes.argtypes = [C.c_void_p, abi.u32p, abi.u64p, abi.u64p]
        L.gto_result_accum.argtypes = [C.c_void_p, C.POINTER(abi.Accumulators)]
        L.gto_result_ref_depth_size.argtypes = [C.c_void_p, abi.u32p, abi.u32p]
        L.gto_result_seed_sizes.argtypes = [C.c_void_p, abi.u64p, abi.u64p, abi.u64p]
        L.gto_result_seeds.argtypes = [C.c_void_p, abi.u32p, abi.u32p, abi.u32p, C.POINTER(abi.Label)]
        L.gto_result_path_sizes.argtypes = [C.c_void_p, abi.u64p, abi.u64p, abi.u64p, abi.u64p]
        L.gto_result_paths.argtypes = [C.c_void_p, abi.u32p, abi.u32p, abi.u32p, abi.u32p, abi.u32p, abi.u16p]
        L.gto_calls_from_accumulators.argtypes = [C.POINTER(abi.Accumulators), abi.u8p, abi.u16p, abi.u8p]
        L.gto_parse_bam.argtypes = [C.POINTER(abi.BamBatch), C.c_int, abi.u8p, abi.u16p, abi.u16p, abi.u8p, abi.i32p, abi.u8p,
                                    abi.u8p, abi.i32p, abi.i32p, abi.u8p]
        L.gto_set_connections.argtypes = [C.c_int]
        L.gto_set_connections.restype = None
        L.gto_result_connections_size.argtypes = [C.c_void_p, abi.u64p]
        L.gto_result_connections.argtypes = [C.c_void_p, C.c_void_p]

    # -- index
    def index_build(self, graph) -> C.c_void_p:
        h = C.c_void_p()
        rc = self.lib.gto_index_build(C.byref(graph.view), C.byref(h))
        if rc:
            raise RuntimeError(self.lib.gto_last_error().decode())
        return h

    def index_export(self, h) -> Dict[str, np.ndarray]:
        nk, nl = C.c_uint64(), C.c_uint64()
        self.lib.gto_index_size(h, C.byref(nk), C.byref(nl))
        keys = np.zeros(nk.value, np.uint64)
        off = np.zeros(nk.value + 1, np.uint32)
        labels = np.zeros(nl.value * 3, np.uint32)
        self.lib.gto_index_export(h, keys.ctypes.data_as(self.abi.u64p), off.ctypes.data_as(self.abi.u32p),
                                  C.cast(labels.ctypes.data, C.POINTER(self.abi.Label)))
        return {"keys": keys, "label_off": off, "labels": labels}

    def index_free(self, h) -> None:
        self.lib.gto_index_free(h)

    # -- pool
    def pool_run(self, graph, index, n_samples: int, batch, tap: bool = True):
        r = C.c_void_p()
        rc = self.lib.gto_pool_run(C.byref(graph.view), index, n_samples, C.byref(batch.view), 1 if tap else 0,
                                   C.byref(r))
        if rc:
            raise RuntimeError(f"oracle pool_run rc={rc}: {self.lib.gto_last_error().decode()}")
        return r

    def result_free(self, r) -> None:
        self.lib.gto_result_free(r)

    def result_stats(self, r):
        s = self.abi.SubmitStats()
        self.lib.gto_result_stats(r, C.byref(s))
        return s

    def result_accum(self, r, n_samples: int):
        nb, ns, nc = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self.lib.gto_result_accum_sizes(r, C.byref(nb), C.byref(ns), C.byref(nc))
        ds, ro = C.c_uint32(), C.c_uint32()
        self.lib.gto_result_ref_depth_size(r, C.byref(ds), C.byref(ro))
        acc = self.abi.HostAccumulators(nb.value, ns.value, nc.value, n_samples, depth_size=ds.value)
        self.lib.gto_result_accum(r, C.byref(acc.view))
        return acc

    def result_seeds(self, r) -> Dict[str, np.ndarray]:
        nu, ns, nl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.gto_result_seed_sizes(r, C.byref(nu), C.byref(ns), C.byref(nl))
        unit_record = np.zeros(nu.value, np.uint32)
        nslots = np.zeros(nu.value * 4, np.uint32)
        nlabels = np.zeros(ns.value, np.uint32)
        labels = np.zeros(nl.value * 3, np.uint32)
        self.lib.gto_result_seeds(r, unit_record.ctypes.data_as(self.abi.u32p), nslots.ctypes.data_as(self.abi.u32p),
                                  nlabels.ctypes.data_as(self.abi.u32p),
                                  C.cast(labels.ctypes.data, C.POINTER(self.abi.Label)))
        return {"unit_record": unit_record, "nslots": nslots, "nlabels": nlabels, "labels": labels}

    def result_paths(self, r) -> Dict[str, np.ndarray]:
        nu, np_, nv, nn = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.gto_result_path_sizes(r, C.byref(nu), C.byref(np_), C.byref(nv), C.byref(nn))
        out = {"gp_npaths": np.zeros(nu.value * 2, np.uint32), "gp_longest": np.zeros(nu.value * 2, np.uint32),
               "p_fields": np.zeros(np_.value * 6, np.uint32), "v_order": np.zeros(nv.value, np.uint32),
               "v_nnum": np.zeros(nv.value, np.uint32), "v_nums": np.zeros(nn.value, np.uint16)}
        a = self.abi
        self.lib.gto_result_paths(r, out["gp_npaths"].ctypes.data_as(a.u32p), out["gp_longest"].ctypes.data_as(a.u32p),
                                  out["p_fields"].ctypes.data_as(a.u32p), out["v_order"].ctypes.data_as(a.u32p),
                                  out["v_nnum"].ctypes.data_as(a.u32p), out["v_nums"].ctypes.data_as(a.u16p))
        return out

    def parse_bam(self, bam, is_sv: bool = False):
        """Records (abi.HostBamBatch) -> abi.HostBatch with every derived column (score_diff, mate, dup_of, leftover)."""
        a = self.abi
        n = len(bam)
        seq4 = np.zeros((n, a.SEQ_STRIDE), np.uint8)
        lseq, flag = np.zeros(n, np.uint16), np.zeros(n, np.uint16)
        mapq, same, sd, left = (np.zeros(n, np.uint8) for _ in range(4))
        isize, mate, dup = (np.zeros(n, np.int32) for _ in range(3))
        rc = self.lib.gto_parse_bam(C.byref(bam.view), 1 if is_sv else 0, seq4.ctypes.data_as(a.u8p), lseq.ctypes.data_as(a.u16p),
                                    flag.ctypes.data_as(a.u16p), mapq.ctypes.data_as(a.u8p), isize.ctypes.data_as(a.i32p),
                                    same.ctypes.data_as(a.u8p), sd.ctypes.data_as(a.u8p), mate.ctypes.data_as(a.i32p),
                                    dup.ctypes.data_as(a.i32p), left.ctypes.data_as(a.u8p))
        if rc:
            raise RuntimeError(f"oracle parse_bam rc={rc}: {self.lib.gto_last_error().decode()}")
        return a.HostBatch(seq4, lseq, flag, mapq, isize, same, sd, np.zeros(n, np.uint8), bam.sample, mate, dup, leftover=left)

    def set_connections(self, on: bool) -> None:
        """Phasing connections (vcf_writer.cpp:587-637) in the following pool_run calls."""
        self.lib.gto_set_connections(1 if on else 0)

    def result_connections(self, r) -> np.ndarray:
        n = C.c_uint64()
        self.lib.gto_result_connections_size(r, C.byref(n))
        out = np.zeros(n.value, self.abi.CONNECTION_DTYPE)
        self.lib.gto_result_connections(r, out.ctypes.data)
        return out

    def phase_support(self, acc, conn: np.ndarray) -> np.ndarray:
        return self.abi.phase_support(self.lib, "gto_phase_support", acc, conn)

    def calls(self, acc):
        phred = np.zeros(len(acc.log_score), np.uint8)
        gt = np.zeros(acc.n_bubbles * acc.n_samples * 2, np.uint16)
        gq = np.zeros(acc.n_bubbles * acc.n_samples, np.uint8)
        a = self.abi
        self.lib.gto_calls_from_accumulators(C.byref(acc.view), phred.ctypes.data_as(a.u8p), gt.ctypes.data_as(a.u16p),
                                             gq.ctypes.data_as(a.u8p))
        return phred, gt, gq


class PawOracle:
    """Scalar restatement of paw::pairwise_alignment + clipping (oracle/paw_oracle.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(build_paw_oracle())
        self.lib.gto_sw_align_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_char_p,
                                                C.POINTER(C.c_int32), C.c_void_p]

    def align(self, queries, windows) -> np.ndarray:
        """int32 [n, 5]: score, database_begin, database_end, clip_begin, clip_end."""
        n = len(queries)
        q_off = np.zeros(n + 1, np.int32)
        d_off = np.zeros(n + 1, np.int32)
        if n:
            np.cumsum([len(q) for q in queries], out=q_off[1:])
            np.cumsum([len(d) for d in windows], out=d_off[1:])
        out = np.zeros((n, 5), np.int32)
        self.lib.gto_sw_align_batch(n, b"".join(queries), q_off.ctypes.data_as(C.POINTER(C.c_int32)), b"".join(windows),
                                    d_off.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.c_void_p))
        return out


def paw_reference(queries, windows) -> Optional[np.ndarray]:
    """The compiled, unmodified paw (oracle/_ref/bin/paw_probe) on the same pairs; None when it is not built."""
    exe = ref_binary("paw_probe")
    if exe is None:
        return None
    import tempfile
    with tempfile.NamedTemporaryFile("wb", suffix=".tsv", delete=False) as f:
        f.write(b"".join(q + b"\t" + d + b"\n" for q, d in zip(queries, windows)))
        path = f.name
    try:
        txt = subprocess.run([exe, path], check=True, capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    return np.array([[int(x) for x in line.split()] for line in txt.strip().split("\n")], np.int32).reshape(-1, 5)

"""Python binding of libgtb200.so (the C ABI of include/gtb200.h) for the harness: tests, bench, smoke.

The product is the shared library; this file only marshals numpy arrays through ctypes.  It raises loudly
when the CUDA library is missing or when a compute entry point is used without a device -- there is no CPU
fallback anywhere on this path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTB_LIB") or os.path.join(HERE, "libgtb200.so")  # GTB_LIB: tuning builds (tools/)

EXPORTS = [
    "gtb_last_error", "gtb_version", "gtb_create", "gtb_destroy", "gtb_region_begin", "gtb_region_begin_multi", "gtb_region_end",
    "gtb_index_size", "gtb_index_export", "gtb_pool_begin", "gtb_submit_reads", "gtb_accumulator_sizes", "gtb_ref_depth_size",
    "gtb_pool_finish", "gtb_pool_finish_multi", "gtb_pool_reset_multi", "gtb_submit_reads_multi", "gtb_debug_enable", "gtb_debug_seed_sizes", "gtb_debug_seeds",
    "gtb_debug_path_sizes", "gtb_debug_paths", "gtb_calls_from_accumulators", "gtb_scan_calls", "gtb_merge_varstats", "gtb_replay_last",
    "gtb_last_timing", "gtb_last_kernel_timing", "gtb_pool_reset", "gtb_set_chunks", "gtb_host_alloc", "gtb_host_free", "gtb_nccl_unique_id", "gtb_nccl_init", "gtb_allreduce_accumulators", "gtb_allreduce_accumulators_multi", "gtb_debug_counters",
    "gtb_sw_align_batch", "gtb_sw_last_timing", "gtb_sw_replay_last", "gtb_set_index_build",
    "gtb_set_connections", "gtb_connections_size", "gtb_connections", "gtb_phase_support", "gtb_last_prep_timing",
    "gtb_submit_bam_records", "gtb_submit_bam_records_multi", "gtb_debug_bam_columns", "gtb_merge_connections", "gtb_sample_depths",
    "gtb_last_chain_timing", "gtb_region_attach", "gtb_allreduce_varstats", "gtb_scan_calls_multi",
    "gtb_allreduce_varstats_multi", "gtb_submit_bgzf", "gtb_debug_bgzf_records", "gtb_debug_bgzf_host", "gtb_debug_bgzf_stitched",
]


class GtbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgtb200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.gtb_last_error.restype = C.c_char_p
    L.gtb_version.restype = C.c_char_p
    L.gtb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.gtb_destroy.argtypes = [vp]
    L.gtb_destroy.restype = None
    L.gtb_region_begin.argtypes = [vp, C.c_int, C.POINTER(abi.GraphView)]
    L.gtb_region_begin_multi.argtypes = [vp, C.c_int, abi.i32p, C.POINTER(abi.GraphView)]
    L.gtb_region_end.argtypes = [vp, C.c_int]
    L.gtb_region_attach.argtypes = [vp, C.c_int, vp, C.c_int]
    L.gtb_index_size.argtypes = [vp, C.c_int, abi.u64p, abi.u64p]
    L.gtb_index_export.argtypes = [vp, C.c_int, abi.u64p, abi.u32p, C.POINTER(abi.Label)]
    L.gtb_pool_begin.argtypes = [vp, C.c_int, C.c_int]
    L.gtb_submit_reads.argtypes = [vp, C.c_int, C.POINTER(abi.ReadBatch), C.POINTER(abi.SubmitStats)]
    L.gtb_submit_reads_multi.argtypes = [vp, C.c_int, abi.i32p, C.POINTER(abi.ReadBatch), C.POINTER(abi.SubmitStats)]
    L.gtb_accumulator_sizes.argtypes = [vp, C.c_int, abi.u32p, abi.u64p, abi.u64p]
    L.gtb_ref_depth_size.argtypes = [vp, C.c_int, abi.u32p, abi.u32p]
    L.gtb_pool_finish.argtypes = [vp, C.c_int, C.POINTER(abi.Accumulators)]
    L.gtb_pool_finish_multi.argtypes = [vp, C.c_int, abi.i32p, C.POINTER(abi.Accumulators)]
    L.gtb_pool_reset_multi.argtypes = [vp, C.c_int, abi.i32p]
    L.gtb_debug_enable.argtypes = [vp, C.c_int]
    L.gtb_debug_seed_sizes.argtypes = [vp, C.c_int, abi.u64p, abi.u64p, abi.u64p]
    L.gtb_debug_seeds.argtypes = [vp, C.c_int, abi.u32p, abi.u32p, abi.u32p, C.POINTER(abi.Label)]
    L.gtb_debug_path_sizes.argtypes = [vp, C.c_int, abi.u64p, abi.u64p, abi.u64p, abi.u64p]
    L.gtb_debug_paths.argtypes = [vp, C.c_int, abi.u32p, abi.u32p, abi.u32p, abi.u32p, abi.u32p, abi.u16p]
    L.gtb_calls_from_accumulators.argtypes = [C.POINTER(abi.Accumulators), abi.u8p, abi.u16p, abi.u8p]
    dp = C.POINTER(C.c_double)
    L.gtb_scan_calls.argtypes = [C.POINTER(abi.Accumulators), abi.u8p, abi.u64p, abi.u64p, dp]
    L.gtb_merge_varstats.argtypes = [C.c_uint32, C.c_uint64, abi.u64p, abi.u64p, dp, abi.u64p, abi.u64p, dp]
    L.gtb_replay_last.argtypes = [vp, C.POINTER(abi.SubmitStats)]
    fp = C.POINTER(C.c_float)
    L.gtb_last_timing.argtypes = [vp, fp, fp, fp, fp]
    L.gtb_last_kernel_timing.argtypes = [vp, fp, fp, fp, fp, abi.u64p]
    L.gtb_pool_reset.argtypes = [vp, C.c_int]
    L.gtb_set_chunks.argtypes = [vp, C.c_int]
    L.gtb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.gtb_host_free.argtypes = [vp]
    L.gtb_nccl_unique_id.argtypes = [abi.u8p]
    L.gtb_nccl_init.argtypes = [vp, C.c_int, C.c_int, abi.u8p]
    L.gtb_allreduce_accumulators.argtypes = [vp, C.c_int, vp]
    L.gtb_allreduce_accumulators_multi.argtypes = [vp, C.c_int, abi.i32p, vp]
    L.gtb_debug_counters.argtypes = [vp, abi.u64p]
    L.gtb_allreduce_varstats.argtypes = [vp, C.c_uint64, C.c_uint64, abi.u64p, abi.u64p, C.POINTER(C.c_double), vp]
    L.gtb_sw_align_batch.argtypes = [vp, C.c_int, abi.u8p, abi.i32p, abi.u8p, abi.i32p, C.c_void_p]
    L.gtb_sw_last_timing.argtypes = [vp, fp, fp, fp]
    L.gtb_sw_replay_last.argtypes = [vp]
    L.gtb_set_index_build.argtypes = [vp, C.c_int]
    L.gtb_last_prep_timing.argtypes = [vp, fp]
    L.gtb_last_chain_timing.argtypes = [vp, fp, fp, abi.u64p, abi.u64p, abi.u64p]
    L.gtb_submit_bam_records.argtypes = [vp, C.c_int, C.POINTER(abi.BamBatch), C.POINTER(abi.SubmitStats)]
    L.gtb_submit_bam_records_multi.argtypes = [vp, C.c_int, abi.i32p, C.POINTER(abi.BamBatch), C.POINTER(abi.SubmitStats)]
    L.gtb_debug_bam_columns.argtypes = [vp, C.c_uint32, abi.u8p, abi.u16p, abi.u16p, abi.u8p, abi.i32p, abi.u8p, abi.u8p, abi.i32p,
                                        abi.i32p, abi.u8p]
    L.gtb_set_connections.argtypes = [vp, C.c_int]
    L.gtb_submit_bgzf.argtypes = [vp, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(abi.SubmitStats)]
    L.gtb_debug_bgzf_stitched.argtypes = [vp, abi.u32p]
    L.gtb_debug_bgzf_records.argtypes = [vp, abi.u32p, abi.u64p, C.c_void_p, abi.u8p, abi.u64p, abi.i32p, abi.i32p]
    L.gtb_debug_bgzf_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, abi.u32p, abi.u64p, C.c_void_p, abi.u8p, abi.u64p, abi.i32p,
                                      abi.i32p, abi.u64p, abi.u8p]
    L.gtb_connections_size.argtypes = [vp, C.c_int, abi.u64p]
    L.gtb_connections.argtypes = [vp, C.c_int, C.c_void_p]
    _lib = L
    return L


class PinnedArena:
    """Page-locked host memory from gtb_host_alloc, handed out as numpy arrays (freed with the arena)."""

    def __init__(self, nbytes: int):
        self.lib = load_library()
        self.ptr = C.c_void_p()
        rc = self.lib.gtb_host_alloc(nbytes, C.byref(self.ptr))
        if rc != 0:
            raise GtbError(rc, self.lib.gtb_last_error().decode())
        self.nbytes = nbytes
        self.used = 0

    def take(self, arr: np.ndarray, skew: int = 0) -> np.ndarray:
        """Copy of `arr` living in the arena (256-byte aligned; `skew` elements further for tests of unaligned columns)."""
        off = (self.used + 255) // 256 * 256 + skew * arr.dtype.itemsize
        if off + arr.nbytes > self.nbytes:
            raise MemoryError("pinned arena exhausted")
        buf = (C.c_uint8 * arr.nbytes).from_address(self.ptr.value + off)
        out = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
        out[...] = arr
        self.used = off + arr.nbytes
        return out

    def close(self) -> None:
        if self.ptr:
            self.lib.gtb_host_free(self.ptr)
            self.ptr = C.c_void_p()


def pin_batches(batches: Sequence[abi.HostBatch], skew: int = 0) -> Tuple[List[abi.HostBatch], "PinnedArena"]:
    """Re-homes the batch columns in page-locked memory (what a production caller fills directly): the bases are then DMA-ed
    and the small columns read by the device straight from these arrays (no host staging)."""
    total = sum(b.nbytes_h2d() + 16 * 256 for b in batches) + 4096
    arena = PinnedArena(total)
    out = []
    for b in batches:
        t = lambda a: arena.take(a, skew)
        out.append(abi.HostBatch(arena.take(b.seq4), t(b.lseq), t(b.flag), t(b.mapq), t(b.isize), t(b.same_tid),
                                 t(b.score_diff), t(b.clipped), t(b.sample), t(b.mate), t(b.dup_of),
                                 leftover=t(b.leftover)))
    return out, arena


def pin_bam_batches(bams: Sequence[abi.HostBamBatch]) -> Tuple[List[abi.HostBamBatch], "PinnedArena"]:
    """Raw-record batches re-homed in page-locked memory."""
    total = sum(b.core.nbytes + b.data.nbytes + b.data_off.nbytes + b.sample.nbytes + b.rg.nbytes + 8 * 256 for b in bams) + 4096
    arena = PinnedArena(total)
    return [abi.HostBamBatch(arena.take(b.core), arena.take(b.data), arena.take(b.data_off), arena.take(b.sample), arena.take(b.rg))
            for b in bams], arena


def bgzf_host(files, query, want_inflated: bool = False):
    """gtb_debug_bgzf_host: decode + selection + merge order of gtb_submit_bgzf computed serially on the CPU from the same
    source functions (parity infrastructure).  Returns the HostBamBatch (and the inflated bytes)."""
    lib = load_library()
    n, nd, ni = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)

    def check(rc):
        if rc != 0:
            raise GtbError(rc, lib.gtb_last_error().decode())
    check(lib.gtb_debug_bgzf_host(files.n_files, C.addressof(files.files), C.addressof(query), C.byref(n), C.byref(nd), None, None,
                                  None, None, None, C.byref(ni), None))
    core = np.zeros(n.value, abi.BAM_CORE_DTYPE)
    data = np.zeros(max(1, nd.value), np.uint8)
    off = np.zeros(n.value + 1, np.uint64)
    smp, rg = np.zeros(n.value, np.int32), np.zeros(n.value, np.int32)
    infl = np.zeros(max(1, ni.value), np.uint8)
    check(lib.gtb_debug_bgzf_host(files.n_files, C.addressof(files.files), C.addressof(query), C.byref(n), C.byref(nd),
                                  core.ctypes.data, abi._ptr(data, abi.u8p), abi._ptr(off, abi.u64p), abi._ptr(smp, abi.i32p),
                                  abi._ptr(rg, abi.i32p), C.byref(ni), abi._ptr(infl, abi.u8p)))
    batch = abi.HostBamBatch(core, data[:nd.value], off, smp, rg)
    return (batch, infl[:ni.value]) if want_inflated else batch


def bgzf_host_stitched() -> int:
    n = C.c_uint32(0)
    load_library().gtb_debug_bgzf_stitched(None, C.byref(n))
    return n.value


class Context:
    """One gtb_ctx: a device (or host-only when device < 0) with resident regions."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        self._check(self.lib.gtb_create(device, C.byref(self.h)))
        self.device = device
        self._graphs: Dict[int, abi.HostGraph] = {}
        self._samples: Dict[int, int] = {}

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise GtbError(rc, self.lib.gtb_last_error().decode())

    def close(self) -> None:
        if self.h:
            self.lib.gtb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- regions / index
    def region_begin(self, region_id: int, graph: abi.HostGraph) -> None:
        self._check(self.lib.gtb_region_begin(self.h, region_id, C.byref(graph.view)))
        self._graphs[region_id] = graph

    def region_begin_multi(self, region_ids: Sequence[int], graphs: Sequence[abi.HostGraph]) -> None:
        """Host index builds of all regions run in parallel threads inside the library."""
        n = len(region_ids)
        ids = (C.c_int32 * n)(*region_ids)
        arr = (abi.GraphView * n)(*[g.view for g in graphs])
        self._check(self.lib.gtb_region_begin_multi(self.h, n, ids, arr))
        for r, g in zip(region_ids, graphs):
            self._graphs[r] = g

    def region_attach(self, region_id: int, owner: "Context", owner_region_id: int) -> None:
        """Shares graph + index of a region another context (same device) built; pools stay per context (one context per
        pool thread, gtb_region_attach)."""
        self._check(self.lib.gtb_region_attach(self.h, region_id, owner.h, owner_region_id))
        self._graphs[region_id] = owner._graphs[owner_region_id]

    def region_end(self, region_id: int) -> None:
        self._check(self.lib.gtb_region_end(self.h, region_id))
        self._graphs.pop(region_id, None)
        self._samples.pop(region_id, None)

    def index_export(self, region_id: int) -> Dict[str, np.ndarray]:
        nk, nl = C.c_uint64(), C.c_uint64()
        self._check(self.lib.gtb_index_size(self.h, region_id, C.byref(nk), C.byref(nl)))
        keys = np.zeros(nk.value, np.uint64)
        off = np.zeros(nk.value + 1, np.uint32)
        labels = np.zeros(nl.value * 3, np.uint32)
        self._check(self.lib.gtb_index_export(self.h, region_id, keys.ctypes.data_as(abi.u64p),
                                              off.ctypes.data_as(abi.u32p),
                                              C.cast(labels.ctypes.data, C.POINTER(abi.Label))))
        return {"keys": keys, "label_off": off, "labels": labels}

    # -- pools
    def pool_begin(self, region_id: int, n_samples: int) -> None:
        self._check(self.lib.gtb_pool_begin(self.h, region_id, n_samples))
        self._samples[region_id] = n_samples

    def set_index_build(self, on_device: bool) -> None:
        """Device-side (default) or host-side construction of the region k-mer indexes."""
        self._check(self.lib.gtb_set_index_build(self.h, 1 if on_device else 0))

    def set_connections(self, on: int = 1) -> None:
        """Phasing connections (HapSample::connections) for pools begun afterwards; `on` > 1 = table slots per record."""
        self._check(self.lib.gtb_set_connections(self.h, int(on)))

    def connections(self, region_id: int) -> np.ndarray:
        """Structured array (abi.CONNECTION_DTYPE) sorted by (sample, hap1, allele1, hap2, allele2)."""
        n = C.c_uint64()
        self._check(self.lib.gtb_connections_size(self.h, region_id, C.byref(n)))
        out = np.zeros(n.value, abi.CONNECTION_DTYPE)
        if n.value:
            self._check(self.lib.gtb_connections(self.h, region_id, out.ctypes.data))
        return out

    def phase_support(self, acc: abi.HostAccumulators, conn: np.ndarray) -> np.ndarray:
        """The `ph` map (hts_parallel_reader.cpp:782-893) as a structured array (abi.PHASE_DTYPE); pure host function."""
        return abi.phase_support(self.lib, "gtb_phase_support", acc, conn)

    def merge_connections(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """Sum of two sorted connection lists of one pool (read-sharded ranks); pure host function."""
        fn = self.lib.gtb_merge_connections
        fn.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, abi.u64p, C.c_void_p]
        a = np.ascontiguousarray(a, dtype=abi.CONNECTION_DTYPE)
        b = np.ascontiguousarray(b, dtype=abi.CONNECTION_DTYPE)
        n = C.c_uint64(0)
        self._check(fn(len(a), a.ctypes.data, len(b), b.ctypes.data, C.byref(n), None))
        out = np.zeros(n.value, abi.CONNECTION_DTYPE)
        self._check(fn(len(a), a.ctypes.data, len(b), b.ctypes.data, C.byref(n), out.ctypes.data))
        return out

    def set_chunks(self, n: int) -> None:
        self._check(self.lib.gtb_set_chunks(self.h, n))

    def pool_reset(self, region_id: int) -> None:
        self._check(self.lib.gtb_pool_reset(self.h, region_id))

    def submit(self, region_id: int, batch: abi.HostBatch) -> abi.SubmitStats:
        st = abi.SubmitStats()
        self._check(self.lib.gtb_submit_reads(self.h, region_id, C.byref(batch.view), C.byref(st)))
        return st

    def submit_multi(self, region_ids: Sequence[int], batches: Sequence[abi.HostBatch]) -> abi.SubmitStats:
        n = len(region_ids)
        ids = (C.c_int32 * n)(*region_ids)
        arr = (abi.ReadBatch * n)(*[b.view for b in batches])
        st = abi.SubmitStats()
        self._check(self.lib.gtb_submit_reads_multi(self.h, n, ids, arr, C.byref(st)))
        return st

    def submit_bam(self, region_id: int, bam: abi.HostBamBatch) -> abi.SubmitStats:
        """Raw htslib records of one pool (abi.HostBamBatch): parsed, paired and de-duplicated on the device."""
        st = abi.SubmitStats()
        self._check(self.lib.gtb_submit_bam_records(self.h, region_id, C.byref(bam.view), C.byref(st)))
        return st

    def submit_bam_multi(self, region_ids: Sequence[int], bams: Sequence[abi.HostBamBatch]) -> abi.SubmitStats:
        n = len(region_ids)
        ids = (C.c_int32 * n)(*region_ids)
        arr = (abi.BamBatch * n)(*[b.view for b in bams])
        st = abi.SubmitStats()
        self._check(self.lib.gtb_submit_bam_records_multi(self.h, n, ids, arr, C.byref(st)))
        return st

    def submit_bgzf(self, region_id: int, files, query) -> abi.SubmitStats:
        """Compressed BGZF segments of the pool's BAM files (graphtyper_b200.bgzf.HostBgzfFiles + bgzf.query): inflated,
        scanned, filtered, merged, parsed and genotyped on the device."""
        st = abi.SubmitStats()
        self._check(self.lib.gtb_submit_bgzf(self.h, region_id, files.n_files, C.addressof(files.files), C.addressof(query),
                                             C.byref(st)))
        return st

    def debug_bgzf_stitched(self) -> int:
        """Files of the last submit_bgzf whose record boundaries came from the per-block walks (no serial walk)."""
        n = C.c_uint32(0)
        self._check(self.lib.gtb_debug_bgzf_stitched(self.h, C.byref(n)))
        return n.value

    def debug_bgzf_records(self) -> abi.HostBamBatch:
        """The record batch the last submit_bgzf built on the device (parity tap)."""
        n, nd = C.c_uint32(0), C.c_uint64(0)
        self._check(self.lib.gtb_debug_bgzf_records(self.h, C.byref(n), C.byref(nd), None, None, None, None, None))
        core = np.zeros(n.value, abi.BAM_CORE_DTYPE)
        data = np.zeros(max(1, nd.value), np.uint8)
        off = np.zeros(n.value + 1, np.uint64)
        smp, rg = np.zeros(n.value, np.int32), np.zeros(n.value, np.int32)
        if n.value:
            self._check(self.lib.gtb_debug_bgzf_records(self.h, C.byref(n), C.byref(nd), core.ctypes.data, abi._ptr(data, abi.u8p),
                                                        abi._ptr(off, abi.u64p), abi._ptr(smp, abi.i32p), abi._ptr(rg, abi.i32p)))
        return abi.HostBamBatch(core, data[:nd.value], off, smp, rg)

    def debug_bam_columns(self, n: int) -> Dict[str, np.ndarray]:
        """Per-record columns the device derived in the last submit_bam (parity tap)."""
        out = {"seq4": np.zeros((n, abi.SEQ_STRIDE), np.uint8), "lseq": np.zeros(n, np.uint16), "flag": np.zeros(n, np.uint16),
               "mapq": np.zeros(n, np.uint8), "isize": np.zeros(n, np.int32), "same_tid": np.zeros(n, np.uint8),
               "score_diff": np.zeros(n, np.uint8), "mate": np.zeros(n, np.int32), "dup_of": np.zeros(n, np.int32),
               "leftover": np.zeros(n, np.uint8)}
        a = abi
        self._check(self.lib.gtb_debug_bam_columns(
            self.h, n, out["seq4"].ctypes.data_as(a.u8p), out["lseq"].ctypes.data_as(a.u16p), out["flag"].ctypes.data_as(a.u16p),
            out["mapq"].ctypes.data_as(a.u8p), out["isize"].ctypes.data_as(a.i32p), out["same_tid"].ctypes.data_as(a.u8p),
            out["score_diff"].ctypes.data_as(a.u8p), out["mate"].ctypes.data_as(a.i32p), out["dup_of"].ctypes.data_as(a.i32p),
            out["leftover"].ctypes.data_as(a.u8p)))
        return out

    def replay(self) -> abi.SubmitStats:
        st = abi.SubmitStats()
        self._check(self.lib.gtb_replay_last(self.h, C.byref(st)))
        return st

    def last_timing(self) -> Tuple[float, float, float, float]:
        a, b, c, d = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        self._check(self.lib.gtb_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def last_kernel_timing(self) -> Dict[str, float]:
        a, b, c, d, n = C.c_float(), C.c_float(), C.c_float(), C.c_float(), C.c_uint64()
        self._check(self.lib.gtb_last_kernel_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(n)))
        p = C.c_float()
        self._check(self.lib.gtb_last_prep_timing(self.h, C.byref(p)))
        return {"prep_kernels": p.value, "probe_kernel": a.value, "chain_kernel": b.value, "slow_kernel": c.value,
                "score_kernel": d.value, "n_slow_tasks": int(n.value)}

    T0_REASONS = ["sv_graph", "labels>8", "bubbles>4", "path_splits", "chains>2", "no_full_chain", "end_in_bubble",
                  "walk_capacity", "walk_splits", "special_pos", "path_pool"]

    def last_chain_timing(self) -> Dict[str, object]:
        """The two tiers of the chaining stage in the last submit/replay (gtb_last_chain_timing)."""
        f, g = C.c_float(), C.c_float()
        n, ng = C.c_uint64(), C.c_uint64()
        r = np.zeros(16, dtype=np.uint64)
        self._check(self.lib.gtb_last_chain_timing(self.h, C.byref(f), C.byref(g), C.byref(n), C.byref(ng), abi._ptr(r, abi.u64p)))
        return {"chain_kernel": f.value, "chain_general_kernel": g.value, "tasks": int(n.value), "general_tasks": int(ng.value),
                "general_reasons": {k: int(v) for k, v in zip(self.T0_REASONS, r) if v}}

    def pool_finish(self, region_id: int) -> abi.HostAccumulators:
        acc = self.alloc_accumulators(region_id)
        self._check(self.lib.gtb_pool_finish(self.h, region_id, C.byref(acc.view)))
        return acc

    def alloc_accumulators(self, region_id: int) -> abi.HostAccumulators:
        nb, ns, nc = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.gtb_accumulator_sizes(self.h, region_id, C.byref(nb), C.byref(ns), C.byref(nc)))
        ds, ro = C.c_uint32(), C.c_uint32()
        self._check(self.lib.gtb_ref_depth_size(self.h, region_id, C.byref(ds), C.byref(ro)))
        return abi.HostAccumulators(nb.value, ns.value, nc.value, self._samples[region_id], depth_size=ds.value)

    def pool_finish_multi(self, region_ids: Sequence[int], out: Optional[List[abi.HostAccumulators]] = None):
        """Accumulators of several regions with one stream synchronisation; `out` buffers are reused when given."""
        n = len(region_ids)
        if out is None:
            out = [self.alloc_accumulators(r) for r in region_ids]
        ids = (C.c_int32 * n)(*region_ids)
        arr = (abi.Accumulators * n)(*[a.view for a in out])
        self._check(self.lib.gtb_pool_finish_multi(self.h, n, ids, arr))
        return out

    def pool_reset_multi(self, region_ids: Sequence[int]) -> None:
        n = len(region_ids)
        ids = (C.c_int32 * n)(*region_ids)
        self._check(self.lib.gtb_pool_reset_multi(self.h, n, ids))

    # -- debug taps
    def debug_enable(self, on: bool = True) -> None:
        self._check(self.lib.gtb_debug_enable(self.h, 1 if on else 0))

    def debug_counters(self) -> List[int]:
        """[0..11] why chain_kernel re-queued tasks for slow_kernel, [12..23] why slow_kernel re-queued for huge_kernel."""
        out = (C.c_uint64 * 24)()
        self._check(self.lib.gtb_debug_counters(self.h, out))
        return [int(x) for x in out]

    def debug_seeds(self, region_id: int) -> Dict[str, np.ndarray]:
        nu, ns, nl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.gtb_debug_seed_sizes(self.h, region_id, C.byref(nu), C.byref(ns), C.byref(nl)))
        unit_record = np.zeros(nu.value, np.uint32)
        nslots = np.zeros(nu.value * 4, np.uint32)
        nlabels = np.zeros(ns.value, np.uint32)
        labels = np.zeros(nl.value * 3, np.uint32)
        self._check(self.lib.gtb_debug_seeds(self.h, region_id, unit_record.ctypes.data_as(abi.u32p),
                                             nslots.ctypes.data_as(abi.u32p), nlabels.ctypes.data_as(abi.u32p),
                                             C.cast(labels.ctypes.data, C.POINTER(abi.Label))))
        return {"unit_record": unit_record, "nslots": nslots, "nlabels": nlabels, "labels": labels}

    def debug_paths(self, region_id: int) -> Dict[str, np.ndarray]:
        nu, np_, nv, nn = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.gtb_debug_path_sizes(self.h, region_id, C.byref(nu), C.byref(np_), C.byref(nv), C.byref(nn)))
        out = {"gp_npaths": np.zeros(nu.value * 2, np.uint32), "gp_longest": np.zeros(nu.value * 2, np.uint32),
               "p_fields": np.zeros(np_.value * 6, np.uint32), "v_order": np.zeros(nv.value, np.uint32),
               "v_nnum": np.zeros(nv.value, np.uint32), "v_nums": np.zeros(nn.value, np.uint16)}
        self._check(self.lib.gtb_debug_paths(self.h, region_id, out["gp_npaths"].ctypes.data_as(abi.u32p),
                                             out["gp_longest"].ctypes.data_as(abi.u32p),
                                             out["p_fields"].ctypes.data_as(abi.u32p),
                                             out["v_order"].ctypes.data_as(abi.u32p),
                                             out["v_nnum"].ctypes.data_as(abi.u32p),
                                             out["v_nums"].ctypes.data_as(abi.u16p)))
        return out

    # -- host finalisation
    def calls(self, acc: abi.HostAccumulators):
        phred = np.zeros(len(acc.log_score), np.uint8)
        gt = np.zeros(acc.n_bubbles * acc.n_samples * 2, np.uint16)
        gq = np.zeros(acc.n_bubbles * acc.n_samples, np.uint8)
        self._check(self.lib.gtb_calls_from_accumulators(C.byref(acc.view), phred.ctypes.data_as(abi.u8p),
                                                         gt.ctypes.data_as(abi.u16p), gq.ctypes.data_as(abi.u8p)))
        return phred, gt, gq

    def sample_depths(self, acc: abi.HostAccumulators):
        """SampleCall::ref_total_depth / alt_total_depth per bubble x sample (sample_call.cpp:34-61)."""
        fn = self.lib.gtb_sample_depths
        fn.argtypes = [C.POINTER(abi.Accumulators), abi.u16p, abi.u16p]
        n = acc.n_bubbles * acc.n_samples
        r, a = np.zeros(n, np.uint16), np.zeros(n, np.uint16)
        self._check(fn(C.byref(acc.view), r.ctypes.data_as(abi.u16p), a.ctypes.data_as(abi.u16p)))
        return r, a

    def scan_calls(self, acc: abi.HostAccumulators, phred: np.ndarray):
        """Variant::scan_calls summary of one pool: (var[NB,9], allele[n_cov,13], ratio[n_cov])."""
        n_cov = int(acc.cov_off[-1])
        var = np.zeros(acc.n_bubbles * 9, np.uint64)
        allele = np.zeros(n_cov * 13, np.uint64)
        ratio = np.zeros(n_cov, np.float64)
        self._check(self.lib.gtb_scan_calls(C.byref(acc.view), phred.ctypes.data_as(abi.u8p), var.ctypes.data_as(abi.u64p),
                                            allele.ctypes.data_as(abi.u64p), ratio.ctypes.data_as(C.POINTER(C.c_double))))
        return var, allele, ratio

    def scan_calls_multi(self, accs: Sequence[abi.HostAccumulators]):
        """Per-pool summaries of several regions in one call: concatenated (var, allele, ratio) rows."""
        n = len(accs)
        arr = (abi.Accumulators * n)(*[a.view for a in accs])
        nb = sum(a.n_bubbles for a in accs)
        na = sum(int(a.cov_off[-1]) for a in accs)
        var, allele, ratio = np.zeros(nb * 9, np.uint64), np.zeros(na * 13, np.uint64), np.zeros(na, np.float64)
        fn = self.lib.gtb_scan_calls_multi
        fn.argtypes = [C.c_int, C.POINTER(abi.Accumulators), abi.u64p, abi.u64p, C.POINTER(C.c_double)]
        self._check(fn(n, arr, var.ctypes.data_as(abi.u64p), allele.ctypes.data_as(abi.u64p),
                       ratio.ctypes.data_as(C.POINTER(C.c_double))))
        return var, allele, ratio

    def merge_varstats(self, var, allele, ratio, var_src, allele_src, ratio_src) -> None:
        """Cross-pool merge VarStats::add_stats of (var_src, allele_src, ratio_src) into (var, allele, ratio), in place."""
        dp = C.POINTER(C.c_double)
        self._check(self.lib.gtb_merge_varstats(len(var) // 9, len(ratio), var.ctypes.data_as(abi.u64p),
                                                allele.ctypes.data_as(abi.u64p), ratio.ctypes.data_as(dp),
                                                var_src.ctypes.data_as(abi.u64p), allele_src.ctypes.data_as(abi.u64p),
                                                ratio_src.ctypes.data_as(dp)))

    # -- multi-GPU
    def allreduce_varstats(self, var: np.ndarray, allele: np.ndarray, ratio: np.ndarray) -> None:
        """VarStats::add_stats over the ranks (sample-sharded runs), in place; rows of any number of regions back to back."""
        self._check(self.lib.gtb_allreduce_varstats(self.h, len(var) // 9, len(ratio), var.ctypes.data_as(abi.u64p),
                                                    allele.ctypes.data_as(abi.u64p),
                                                    ratio.ctypes.data_as(C.POINTER(C.c_double)), None))

    def allreduce_varstats_multi(self, triples) -> None:
        """Merges the (var, allele, ratio) summaries of this rank's pools and reduces them over the ranks; result in triples[0]."""
        n = len(triples)
        dp = C.POINTER(C.c_double)
        V = (abi.u64p * n)(*[t[0].ctypes.data_as(abi.u64p) for t in triples])
        A = (abi.u64p * n)(*[t[1].ctypes.data_as(abi.u64p) for t in triples])
        R = (dp * n)(*[t[2].ctypes.data_as(dp) for t in triples])
        fn = self.lib.gtb_allreduce_varstats_multi
        fn.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(abi.u64p), C.POINTER(abi.u64p), C.POINTER(dp), C.c_void_p]
        self._check(fn(self.h, n, len(triples[0][0]) // 9, len(triples[0][2]), V, A, R, None))

    def nccl_unique_id(self) -> np.ndarray:
        buf = np.zeros(128, np.uint8)
        self._check(self.lib.gtb_nccl_unique_id(buf.ctypes.data_as(abi.u8p)))
        return buf

    def nccl_init(self, n_ranks: int, rank: int, uid: np.ndarray) -> None:
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._check(self.lib.gtb_nccl_init(self.h, n_ranks, rank, uid.ctypes.data_as(abi.u8p)))

    def allreduce(self, region_id: int) -> None:
        self._check(self.lib.gtb_allreduce_accumulators(self.h, region_id, None))

    def allreduce_multi(self, region_ids: Sequence[int]) -> None:
        n = len(region_ids)
        ids = (C.c_int32 * n)(*region_ids)
        self._check(self.lib.gtb_allreduce_accumulators_multi(self.h, n, ids, None))

    # -- discovery re-alignment
    def sw_align(self, queries: Sequence[bytes], databases: Sequence[bytes]) -> np.ndarray:
        """paw::pairwise_alignment + clipping of each (read, haplotype window) pair as realign_to_indels configures it
        (src/typer/caller.cpp:1864-1870,2007).  Returns int32 [n, 5]: score, database_begin, database_end, clip_begin,
        clip_end."""
        q, q_off = pack_sequences(queries)
        d, d_off = pack_sequences(databases)
        return self.sw_align_packed(q, q_off, d, d_off)

    def sw_align_packed(self, q: np.ndarray, q_off: np.ndarray, d: np.ndarray, d_off: np.ndarray) -> np.ndarray:
        n = len(q_off) - 1
        if len(d_off) - 1 != n:
            raise ValueError("queries and databases differ in number")
        out = np.zeros((n, 5), np.int32)
        self._check(self.lib.gtb_sw_align_batch(self.h, n, q.ctypes.data_as(abi.u8p), q_off.ctypes.data_as(abi.i32p),
                                                d.ctypes.data_as(abi.u8p), d_off.ctypes.data_as(abi.i32p),
                                                out.ctypes.data_as(C.c_void_p)))
        return out

    def sw_replay(self) -> None:
        self._check(self.lib.gtb_sw_replay_last(self.h))

    def sw_last_timing(self) -> Dict[str, float]:
        k, h, d = C.c_float(), C.c_float(), C.c_float()
        self._check(self.lib.gtb_sw_last_timing(self.h, C.byref(k), C.byref(h), C.byref(d)))
        return {"kernel_ms": k.value, "h2d_ms": h.value, "d2h_ms": d.value}


def pack_sequences(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenates byte strings: (uint8 bases, int32 offsets[n+1])."""
    off = np.zeros(len(seqs) + 1, np.int32)
    if len(seqs):
        np.cumsum([len(s) for s in seqs], out=off[1:])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, off

"""Harness-side construction of the flat region graph for NON-OVERLAPPING biallelic/multiallelic VCF records.

In a deployment the reference's own host code builds the graph (src/graph/constructor.cpp:1597-1777 stays on
the host) and hands the flattened view to gtb_region_begin.  The bench and the
large-size tests run on a GPU box where the reference sources are absent, so this module restates the simple
case they need -- records that do not overlap or abut, no SV tags, no events -- and tests/test_graph_build.py
pins it against graphs dumped from the compiled reference.

Layout produced (include/gtb200.h gtb_graph_view): ref node 0 starts at the padded region's first base; each
record becomes a bubble whose var nodes carry the VCF alleles verbatim (allele 0 = REF); special positions are
created for every alternative-allele position beyond the reference allele's reach (graph.cpp:384-407).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import abi
from .synth import Site

SPECIAL_START = 0xD0000000


def build_region_graph(ref: np.ndarray, sites: Sequence[Site], begin1: int, end1: int, pad: int = 1000,
                       contig_offset: int = 0) -> abi.HostGraph:
    """ref: uint8 ASCII contig; region [begin1, end1] 1-based inclusive, padded by `pad` and clipped to the contig
    exactly as genotype() does (src/utilities/genotype.cpp:401-402, GenomicRegion::pad)."""
    L = len(ref)
    B = max(1, begin1 - pad)
    E = min(L, end1 + pad)
    recs = [s for s in sites if B <= s.pos <= E and s.pos + len(s.ref) - 1 <= E]
    for a, b in zip(recs, recs[1:]):
        if b.pos <= a.pos + len(a.ref):
            raise ValueError("overlapping or abutting records are not supported by this harness builder")
    ref_order: List[int] = []
    ref_seq_off: List[int] = [0]
    ref_var_off: List[int] = [0]
    var_order: List[int] = []
    var_seq_off: List[int] = []
    var_out_ref: List[int] = []
    pieces: List[bytes] = []
    var_pieces: List[bytes] = []
    actual, rreach = [], []
    sp_keys, sp_off, sp_list = [], [0], []

    cur = B  # 1-based position of the next reference base to place
    nvar = 0
    for r, s in enumerate(recs):
        ref_order.append(contig_offset + cur)
        pieces.append(bytes(ref[cur - 1:s.pos - 1]))
        ref_seq_off.append(ref_seq_off[-1] + (s.pos - cur))
        alleles = [s.ref] + ([s.alt] if isinstance(s.alt, (bytes, bytearray)) else list(s.alt))
        for a in alleles:
            var_order.append(contig_offset + s.pos)
            var_out_ref.append(r + 1)
            var_pieces.append(bytes(a))
        nvar += len(alleles)
        ref_var_off.append(nvar)
        ref_reach = contig_offset + s.pos + len(s.ref) - 1
        max_reach = max(contig_offset + s.pos + len(a) - 1 for a in alleles[1:])
        if max_reach > ref_reach:
            sp_keys.append(ref_reach)
            for reach in range(ref_reach + 1, max_reach + 1):
                sp_list.append(SPECIAL_START + len(actual))
                actual.append(reach)
                rreach.append(ref_reach)
            sp_off.append(len(sp_list))
        cur = s.pos + len(s.ref)
    ref_order.append(contig_offset + cur)
    pieces.append(bytes(ref[cur - 1:E]))
    ref_seq_off.append(ref_seq_off[-1] + (E - cur + 1))
    ref_var_off.append(nvar)

    ref_bytes = b"".join(pieces)
    off = len(ref_bytes)
    for vp in var_pieces:
        var_seq_off.append(off)
        off += len(vp)
    var_seq_off.append(off)
    seq = np.frombuffer(ref_bytes + b"".join(var_pieces), dtype=np.uint8)
    arrays = {
        "ref_order": np.array(ref_order, np.uint32), "ref_seq_off": np.array(ref_seq_off, np.uint64),
        "ref_var_off": np.array(ref_var_off, np.uint32), "var_order": np.array(var_order, np.uint32),
        "var_seq_off": np.array(var_seq_off, np.uint64), "var_out_ref": np.array(var_out_ref, np.uint32),
        "seq": seq, "var_ev_off": np.zeros(nvar + 1, np.uint32), "var_ev": np.zeros(0, np.int64),
        "var_aev_off": np.zeros(nvar + 1, np.uint32), "var_aev": np.zeros(0, np.int64),
        "actual_poses": np.array(actual, np.uint32), "ref_reach_poses": np.array(rreach, np.uint32),
        "sp_keys": np.array(sp_keys, np.uint32), "sp_off": np.array(sp_off, np.uint32),
        "sp_list": np.array(sp_list, np.uint32),
    }
    return abi.HostGraph(arrays, is_sv_graph=False)

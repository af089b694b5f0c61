"""N > 1 host logic on CPU (gloo, world_size 2): one pool's records are sharded over ranks keeping mate / duplicate
groups together, every rank accumulates its shard (oracle stands in for the GPU here), the widened accumulators are
summed with a collective and must equal the unsharded result bit for bit -- the contract of
gtb_allreduce_accumulators."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, fixture_prefixes

ADDITIVE = ["log_score", "gt_coverage", "max_log_score", "ambiguous_depth", "ambiguous_depth_alt",
            "alt_proper_pair_depth", "vs_clipped_reads", "vs_mapq_squared", "pa_clipped_bp", "pa_mapq_squared",
            "pa_score_diff", "pa_mismatches", "read_strand"]


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, pre: str, out_dir: str) -> None:
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from graphtyper_b200 import abi, gtba
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = oracle.Oracle()
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    ns = len(rd["sample_names"].tobytes().split(b"\n")) - 1
    full = abi.batch_from_probe(rd)
    shard = abi.shard_batch(full, world)[rank]
    h = O.index_build(g)
    r = O.pool_run(g, h, ns, shard, tap=False)
    acc = O.result_accum(r, ns).as_dict()
    summed = {}
    for k in ADDITIVE:
        t = torch.from_numpy(acc[k].astype(np.int64))
        dist.all_reduce(t)
        summed[k] = t.numpy()
    n_rec = torch.tensor([len(shard)])
    dist.all_reduce(n_rec)
    if rank == 0:
        np.savez(os.path.join(out_dir, "summed.npz"), n_rec=n_rec.numpy(), **summed)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["mini_stress.r0", "tiny.r0"])
def test_sharded_accumulators_sum_to_unsharded(name, tmp_path, oracle_lib):
    pre = os.path.join(ROOT, "tests", "golden", name)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), pre, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "summed.npz"))
    from graphtyper_b200 import abi, gtba
    import compare
    ref = compare.probe_accum(gtba.load(pre + ".accum.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    assert int(got["n_rec"][0]) == len(rd["flag"])
    for k in ADDITIVE:
        assert np.array_equal(got[k], np.asarray(ref[k]).astype(np.int64)), k

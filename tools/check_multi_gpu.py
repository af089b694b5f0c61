"""Run under torchrun on N GPUs (tests/test_gpu_multi.py does):
 * read sharding: one pool's records sharded over the ranks (abi.shard_batch), every rank accumulates its shard on its GPU,
   ONE grouped NCCL all-reduce, result compared bit-for-bit with the reference's golden accumulators;
 * sample sharding: every rank genotypes the whole pool (as if each rank owned a copy of the samples), the per-variant
   summaries are merged over the ranks with gtb_allreduce_varstats and must equal N host-side merges
   (gtb_merge_varstats = VarStats::add_stats, pinned to the reference on the CPU suite)."""
import glob, os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import compare
from graphtyper_b200 import abi, engine, gtba
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
ctx = engine.Context(lr)
uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{lr}")
if rank == 0:
    uid.copy_(torch.from_numpy(ctx.nccl_unique_id()))
dist.broadcast(uid, 0)
ctx.nccl_init(world, rank, uid.cpu().numpy())
from conftest import fixture_prefixes
pres = fixture_prefixes(include_big=True)
ok = True
for k, pre in enumerate(pres):
    g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
    rd = gtba.load(pre + ".reads.gtba")
    ns = len(rd["sample_names"].tobytes().split(b"\n")) - 1
    shard = abi.shard_batch(abi.batch_from_probe(rd), world)[rank]
    ctx.region_begin(k, g); ctx.pool_begin(k, ns)
    ctx.submit(k, shard)
    ctx.allreduce_multi([k])
    acc = ctx.pool_finish(k)
    try:
        compare.compare_accum(compare.probe_accum(gtba.load(pre + ".accum.gtba")), acc.as_dict(), f"rank{rank}")
        if rank == 0: print("OK  ", os.path.basename(pre), "shard sizes", len(shard))
    except AssertionError as e:
        ok = False; print("FAIL", os.path.basename(pre), e)
    # sample sharding: local summaries -> NCCL sum / max
    full = abi.batch_from_probe(rd)
    ctx.pool_reset(k)
    ctx.submit(k, full)
    acc = ctx.pool_finish(k)
    ph, _, _ = ctx.calls(acc)
    var, allele, ratio = ctx.scan_calls(acc, ph)
    ev, ea, er = var.copy(), allele.copy(), ratio.copy()
    for _ in range(world - 1):
        ctx.merge_varstats(ev, ea, er, var, allele, ratio)
    v1, a1, r1 = var.copy(), allele.copy(), ratio.copy()
    ctx.allreduce_varstats(v1, a1, r1)
    if not (np.array_equal(v1, ev) and np.array_equal(a1, ea) and np.array_equal(r1, er)):
        ok = False; print("FAIL varstats", os.path.basename(pre), "rank", rank)
    # two pools per rank merged while packing (gtb_allreduce_varstats_multi) = 2 * world host merges
    for _ in range(world):
        ctx.merge_varstats(ev, ea, er, var, allele, ratio)
    trip = [(var.copy(), allele.copy(), ratio.copy()), (var.copy(), allele.copy(), ratio.copy())]
    ctx.allreduce_varstats_multi(trip)
    if not (np.array_equal(trip[0][0], ev) and np.array_equal(trip[0][1], ea) and np.array_equal(trip[0][2], er)):
        ok = False; print("FAIL varstats multi", os.path.basename(pre), "rank", rank)
    ctx.region_end(k)
flag = torch.tensor([1 if ok else 0], device=f"cuda:{lr}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
ok = bool(int(flag[0]))
dist.barrier()
if rank == 0: print("multi-GPU sharded parity:", "PASS" if ok else "FAIL")
dist.destroy_process_group()

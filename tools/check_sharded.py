"""Multi-GPU parity: ShardedForward (dst-range partition + all-gather) against the single-GPU forward.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
Exits non-zero on mismatch.  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from graph_hypernetwork_forge.distributed import ShardedForward  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
w = dict(bench.WORKLOADS["c3"], N=200_003, E=1_300_000, R=61, L=3)   # N not divisible by the world size
ok = True
for prec, tol in (("f16", 1.5e-4), ("tf32", 1.5e-4), ("fp32", 1e-4)):
    model = bench.build_model(w, dev, prec)
    with torch.no_grad():                      # O(1) generated weights so that the update matters
        for gen in model.weight_generators:
            for p in gen.log_scales.values():
                p.fill_(-1.5)
    x, ei, _rel, utf8, offsets = bench.make_device_inputs(w, dev)
    single = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, w["N"]))
    errs = []
    variants = [dict(), dict(transport="collective")]
    if prec == "f16":
        variants += [dict(push="kernel"), dict(push="copy", chunks=3), dict(push="copy", balance=True, chunks=2),
                     dict(push="kernel", balance=True),
                     dict(transport="collective", balance=True)]
    for kw in variants:
        kw = dict(kw)
        ranges = None
        if kw.pop("balance", False):
            from graph_hypernetwork_forge.distributed import plan_partition_by_edges
            rowptr = torch.zeros(w["N"] + 1, dtype=torch.int64, device=dev)
            rowptr[1:] = torch.cumsum(torch.bincount(ei[1], minlength=w["N"]), 0)
            ranges = plan_partition_by_edges(rowptr, world)
        sf = ShardedForward(model, w["N"], dist.group.WORLD, ranges=ranges, **kw)
        local = sf.forward_packed(x, ei, utf8, offsets)                        # the rank's own rows
        full = sf.forward_packed(x, ei, utf8, offsets, gather_output=True)     # every row on every rank
        e1 = float((single[sf.lo:sf.hi] - local).abs().max()) if sf.hi > sf.lo else 0.0
        e2 = float((single - full).abs().max())
        errs.append(max(e1, e2))
        if rank == 0:
            print(f"{prec}: world {world} {kw or 'default'} ranges={'edges' if ranges else 'nodes'} transport={sf.transport}"
                  f"  max|single - sharded| = {errs[-1]:.3e}  (tol {tol:g})", flush=True)
        ok = ok and bool(torch.isfinite(full).all())
    err = max(errs)
    ok = ok and err <= tol
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))

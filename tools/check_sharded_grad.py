"""Multi-GPU gradients: `loss.backward()` through ShardedForward (every rank: a loss over its own rows) +
`allreduce_gradients()` against the single-GPU model on the same inputs - every parameter gradient and the gradient of
the node features.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_sharded_grad.py
Exits non-zero on mismatch.  GPU box only."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from graph_hypernetwork_forge.distributed import ShardedForward, plan_partition_by_edges  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
# N divisible by the world size (reduce-scatter of the message term) and not (all-reduce), node- and edge-balanced
for N, balance, entry in ((60_000, False, "packed"), (60_003, False, "packed"), (60_003, True, "packed"),
                          (60_003, False, "ids"), (60_000, False, "list")):
    w = dict(bench.WORKLOADS["c3"], N=N, E=400_000, R=23, L=2)
    model = bench.build_model(w, dev, "f16").train()
    with torch.no_grad():                      # O(1) generated weights so that the update matters
        for gen in model.weight_generators:
            for p in gen.log_scales.values():
                p.fill_(-1.5)
    x, ei, _rel, utf8, offsets = bench.make_device_inputs(w, dev)
    g = torch.Generator(device=dev).manual_seed(7)
    loss_w = torch.randn(N, w["d"], generator=g, device=dev)
    # single GPU: the whole loss
    ref = copy.deepcopy(model)
    x_ref = x.clone().requires_grad_(True)
    out_ref = ref.forward_prepared(x_ref, ref.prepare_packed(ei, utf8, offsets, N))
    (out_ref * loss_w).sum().backward()
    # sharded: each rank its rows
    ranges = None
    if balance:
        rowptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(torch.bincount(ei[1], minlength=N), 0)
        ranges = plan_partition_by_edges(rowptr, world)
    sf = ShardedForward(model, N, dist.group.WORLD, ranges=ranges)
    x_sh = x.clone().requires_grad_(True)
    names = [f"relation_{r:05d}" for r in range(w["R"])]
    if entry == "ids":                         # the ids-in entry: relation ids + vocabulary, whole edge list
        out = sf.forward_ids(x_sh, ei, _rel, names)
    elif entry == "list":                      # the reference's call shape
        out = sf.forward(x_sh, ei, [names[r] for r in _rel.tolist()])
    else:
        out = sf.forward_packed(x_sh, ei, utf8, offsets)
    (out * loss_w[sf.lo:sf.hi]).sum().backward()
    sf.allreduce_gradients()
    gx = x_sh.grad.clone()
    dist.all_reduce(gx)                        # every rank holds the gradient of its own rows only
    errs = {"out": (float((out.detach() - out_ref.detach()[sf.lo:sf.hi]).abs().max()), float(out_ref.abs().max())),
            "grad x": (float((gx - x_ref.grad).abs().max()), float(x_ref.grad.abs().max()))}
    for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        errs["grad " + k] = (float((p.grad - q.grad).abs().max()), float(q.grad.abs().max()))
    worst = max(errs.items(), key=lambda kv: kv[1][0] / max(kv[1][1], 1e-30))
    rel = worst[1][0] / max(worst[1][1], 1e-30)
    # both sides run the f16 engine; they differ in the fp16 scale of the layer inputs (agreed maximum vs LayerNorm
    # bound) and in summation order: 2e-3 of each tensor's maximum (the gradient tests' bound against fp64 is 2e-2)
    good = rel <= 2e-3 and all(v[0] == v[0] for v in errs.values())
    ok = ok and good
    if rank == 0:
        print(f"N={N} ranges={'edges' if balance else 'nodes'} entry={entry} world {world}: worst {worst[0]}: |diff| {worst[1][0]:.3e} "
              f"of max {worst[1][1]:.3e} = {rel:.2e} (bound 2e-3)  out diff {errs['out'][0]:.2e}  "
              f"{'ok' if good else 'MISMATCH'}", flush=True)
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))

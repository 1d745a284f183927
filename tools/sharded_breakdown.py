"""Where a sharded forward spends its time (rank 0 prints): device time per stage with CUDA events, and the host time
of the same call.   torchrun --nproc-per-node N tools/sharded_breakdown.py        GPU box only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from graph_hypernetwork_forge import _native  # noqa: E402
from graph_hypernetwork_forge import distributed as D  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
w = bench.WORKLOADS["c3"]
model = bench.build_model(w, dev, "f16")
x, ei, _rel, utf8, offsets = bench.make_device_inputs(w, dev)
sf = D.ShardedForward(model, w["N"], dist.group.WORLD)

marks = []
orig = {}


def wrap(mod, name, label):
    fn = getattr(mod, name)
    orig[(mod, name)] = fn

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        e1.record()
        marks.append((label, e0, e1, time.perf_counter() - t0))
        return r
    setattr(mod, name, inner)


wrap(_native, "select_edges", "select")
wrap(_native, "dedup_texts", "dedup")
wrap(_native, "linear", "linear")
wrap(D, "gather_rows", "gather(enqueue)")
for _ in range(3):
    sf.forward_packed(x, ei, utf8, offsets)
dist.barrier(); torch.cuda.synchronize()
marks.clear()
t0 = time.perf_counter()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
n = 5
for _ in range(n):
    sf.forward_packed(x, ei, utf8, offsets)
ev1.record()
host = time.perf_counter() - t0
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}: {ev0.elapsed_time(ev1) / n:.3f} ms per forward (device), host enqueue {1e3 * host / n:.3f} ms")
    agg = {}
    for label, e0, e1, h in marks:
        a = agg.setdefault(label, [0, 0.0, 0.0])
        a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += 1e3 * h
    for label, (c, dms, hms) in agg.items():
        print(f"  {label:18s} x{c / n:4.1f}  device {dms / n:7.3f} ms   host {hms / n:7.3f} ms   per forward")
dist.destroy_process_group()

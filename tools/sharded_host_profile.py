"""Host-side cost of one sharded forward at a rank's share of c3 on 8 GPUs (N/8 nodes, E/8 edges), on ONE GPU with a
world-size-1 process group: the per-rank GPU work is small, so the step is bound by the host enqueueing it.
    python tools/sharded_host_profile.py        GPU box only."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from graph_hypernetwork_forge.distributed import ShardedForward  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29571")
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
w = dict(bench.WORKLOADS["c3"])
w["N"], w["E"] = w["N"] // 8, w["E"] // 8
model = bench.build_model(w, dev, "f16")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
sf = ShardedForward(model, w["N"], dist.group.WORLD, transport=os.environ.get("GHF_TRANSPORT"), chunks=1)


def step():
    return sf.forward_packed(x, ei, utf8, offsets)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 30
for _ in range(n):
    step()
torch.cuda.synchronize()
print(f"sharded forward, world 1, N={w['N']} E={w['E']}: {1e3 * (time.perf_counter() - t0) / n:.3f} ms per step (wall)")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
dist.destroy_process_group()

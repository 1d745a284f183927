"""Host and device timeline of one sharded forward at a rank's share of c3 on 8 GPUs (world-size-1 group on ONE GPU):
for every native call, when the host issued it and how long the call blocked; plus device time between the stage marks.
    python tools/sharded_timeline.py        GPU box only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from graph_hypernetwork_forge import _native  # noqa: E402
from graph_hypernetwork_forge.distributed import ShardedForward  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29572")
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
w = dict(bench.WORKLOADS["c3"])
share = int(os.environ.get("GHF_SHARE", "8"))
w["N"], w["E"] = w["N"] // share, w["E"] // share
model = bench.build_model(w, dev, "f16")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
sf = ShardedForward(model, w["N"], dist.group.WORLD, transport=os.environ.get("GHF_TRANSPORT"), chunks=1)

log = []
L = _native.lib()


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        log.append((self.name, t0, time.perf_counter()))
        return r


for name in _native.EXPORTED_SYMBOLS:
    if name not in ("ghf_last_error", "ghf_abi_version", "ghf_device_ok", "ghf_launch_count"):
        setattr(L, name, Timed(name, getattr(L, name)))


def step():
    return sf.forward_packed(x, ei, utf8, offsets)


for _ in range(5):
    step()
torch.cuda.synchronize()
for it in range(3):
    log.clear()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"--- step {it}: host issue {1e3 * (t1 - t0):.3f} ms, until idle {1e3 * (t2 - t0):.3f} ms, device span {e0.elapsed_time(e1):.3f} ms")
    for name, a, b in log:
        print(f"  {1e3 * (a - t0):7.3f} ms  +{1e3 * (b - a):6.3f}  {name}")
dist.destroy_process_group()

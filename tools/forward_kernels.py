"""Per-kernel device time of one warm forward on a bench workload (torch.profiler).
    python tools/forward_kernels.py [c4]         GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, None)
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
for _ in range(3):
    model.forward_packed(x, ei, utf8, offsets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.forward_packed(x, ei, utf8, offsets)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=64))

"""A/B of the weight-gradient kernel on a bench workload: GHF_WGRAD_FLAGS values in one process (same box, same
clocks).   python tools/wgrad_sweep.py [c3] [flags ...]          GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)
import bench  # noqa: E402
from graph_hypernetwork_forge import _native  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
flag_sets = sys.argv[2:] or ["0", "1", "4", "5", "7"]
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, "f16")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
graph = model.prepare_packed(ei, utf8, offsets, w["N"]).graph
h = torch.randn(w["N"], w["d"], device=dev)
g_acc = torch.randn(w["N"], w["d"], device=dev) * 1e-3
h16 = _native.to_f16(h, _native.Shadow(torch.empty(h.shape, dtype=torch.float16, device=dev)))
g16 = _native.to_f16(g_acc, _native.Shadow(torch.empty(h.shape, dtype=torch.float16, device=dev)))
for rnd in range(2):
    for f in flag_sets:
        os.environ["GHF_WGRAD_FLAGS"] = f
        for _ in range(2):
            graph.weight_grad(h, g_acc, _native.PREC_F16, h16=h16, g16=g16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.weight_grad(h, g_acc, _native.PREC_F16, h16=h16, g16=g16)
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} flags {f:>3s}: {e0.elapsed_time(e1) / 5:.3f} ms per weight_grad (incl. bias kernel + 3 memsets)")

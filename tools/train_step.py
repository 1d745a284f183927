"""Forward + backward of one training step on a bench workload: device time of each half (CUDA events) and, with
--kernels, a per-kernel table of the step from torch.profiler.     python tools/train_step.py [c3] [--kernels]
GPU box only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

import bench  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = args[0] if args else "c3"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
prec = os.environ.get("GHF_PRECISION") or ("f16" if w["d"] in (64, 128, 256) else "tf32" if w["d"] == 32 else "fp32")
model = bench.build_model(w, dev, prec).train()
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
prepared = model.prepare_packed(ei, utf8, offsets, w["N"])
loss_w = torch.randn(w["N"], w["d"], device=dev)


def step():
    model.zero_grad(set_to_none=True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = model.forward_prepared(x, prepared)
    loss = (out * loss_w).sum()
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])


for _ in range(2):
    step()
times = [step() for _ in range(5)]
fwd = sorted(t[0] for t in times)[len(times) // 2]
bwd = sorted(t[1] for t in times)[len(times) // 2]
EL = w["E"] * w["L"]
print(f"{wl} {prec}: forward {fwd:.2f} ms, backward {bwd:.2f} ms, step {fwd + bwd:.2f} ms "
      f"= {EL / (fwd + bwd) / 1e6:.3f} G edge-layers/s (fwd+bwd); peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if "--kernels" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))

"""Per-step device times of the whole forward (native single call vs Python-staged), to spot outliers.  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph

import bench  # noqa: E402

w = bench.WORKLOADS["c3"]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, "f16")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
paths = {"native": lambda: model.forward_packed(x, ei, utf8, offsets),
         "python": lambda: model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, w["N"]))}
for rep in range(2):
    for name, fn in paths.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
        evs[0].record()
        for i in range(12):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        print(name, " ".join(f"{evs[i].elapsed_time(evs[i + 1]):.2f}" for i in range(12)), flush=True)

"""Two whole forwards on the bench workload (for `ncu --metrics gpu__time_duration.sum`: the second is the warm one).
    python tools/one_forward.py [c3]        GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph

import bench  # noqa: E402
from graph_hypernetwork_forge import _native  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, "f16" if w["d"] in (64, 128, 256) else "tf32" if w["d"] == 32 else "fp32")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
torch.cuda.synchronize()
for i in range(2):
    _native.launch_count(reset=True)
    if os.environ.get("GHF_ONE_FORWARD_STAGED"):       # stage by stage from Python
        out = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, w["N"]))
    else:                                               # the benched entry: one native call
        out = model.forward_packed(x, ei, utf8, offsets)
    torch.cuda.synchronize()
    print(f"forward {i}: {_native.launch_count()} library launches", flush=True)

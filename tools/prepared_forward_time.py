"""Device time of HyperGNN.forward_prepared on a graph prepared once (bench workload), with the generators on the side
stream and (GHF_NO_SIDE_GENERATORS=1) in stream order.     python tools/prepared_forward_time.py [c3]   GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)
import bench  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, "f16" if w["d"] in (64, 128, 256) else "tf32" if w["d"] == 32 else "fp32")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
prep = model.prepare_packed(ei, utf8, offsets, w["N"])
for mode in ("side", "inline", "side", "inline"):
    if mode == "inline":
        os.environ["GHF_NO_SIDE_GENERATORS"] = "1"
    else:
        os.environ.pop("GHF_NO_SIDE_GENERATORS", None)
    for _ in range(3):
        out = model.forward_prepared(x, prep)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        out = model.forward_prepared(x, prep)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward_prepared, generators {mode}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)

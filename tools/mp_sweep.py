"""Time the message-passing layer alone on the bench workload for several (sb_nodes, unit_edges, flags).
    python tools/mp_sweep.py [--workload c3] "sb,unit,flags" ...      (GPU box only)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph

import bench  # noqa: E402
from graph_hypernetwork_forge import _native  # noqa: E402

args = sys.argv[1:]
wl, prec_name = "c3", "f16"
while args and args[0] in ("--workload", "--precision"):
    if args[0] == "--workload":
        wl = args[1]
    else:
        prec_name = args[1]
    args = args[2:]
PREC = _native.precision_code(prec_name)
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, prec_name)
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
h = _native.linear(x, model.input_proj.weight, model.input_proj.bias, relu=True)
# calibration: device copy bandwidth on this box (read + write bytes), and clocks
import subprocess
a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
print(f"copy bandwidth {10 * 2 * a.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.0f} GB/s", flush=True)
del a, b
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,serial",
                      "--format=csv,noheader"], capture_output=True, text=True).stdout.strip(), flush=True)
packed = None
for spec in args or ["0,0,7"]:
    vals = [int(v) for v in spec.split(",")]
    sb, unit, flags = vals[:3]
    os.environ["GHF_SB_NODES"], os.environ["GHF_UNIT_EDGES"], os.environ["GHF_MP_FLAGS"] = str(sb), str(unit), str(flags)
    prepared = model.prepare_packed(ei, utf8, offsets, w["N"])
    g = prepared.graph
    text = model.text_encoder.encode_packed(prepared.packed)
    wts = model.weight_generators[0](text)
    ln = model.layer_norms[0]
    h16 = None
    if PREC == _native.PREC_F16:                             # as chained from the previous layer
        h16 = _native.to_f16(h, _native.Shadow(torch.empty(h.shape, dtype=torch.float16, device=dev)))
    for _ in range(2):
        g.mp_layer(h, wts["W_msg"], wts["W_self"], wts["bias"], ln.weight, ln.bias, 1e-5, PREC, h16=h16)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    _native.profile_read()
    for _ in range(5):
        g.mp_layer(h, wts["W_msg"], wts["W_self"], wts["bias"], ln.weight, ln.bias, 1e-5, PREC, h16=h16)
    prof, n = _native.profile_read()
    _native.profile_enable(False)
    print(f"sb={g.sb_nodes:6d} unit={g.unit_edges:5d} flags={flags} units={g.num_units:7d} "
          f"contraction {prof['contraction_ms']/n:.3f} ms  epilogue {prof['epilogue_ms']/n:.3f}  prep {prof['prep_ms']/n:.3f}",
          flush=True)
    del prepared, g

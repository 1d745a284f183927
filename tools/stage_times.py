"""Per-stage wall time of one forward on the bench workload (synchronising between stages).  GPU box only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph

import bench  # noqa: E402
from graph_hypernetwork_forge import _native  # noqa: E402
from graph_hypernetwork_forge.models.hypergnn import PackedTexts  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, "f16" if w["d"] in (64, 128, 256) else "tf32" if w["d"] == 32 else "fp32")
PREC = _native.precision_code("f16" if w["d"] in (64, 128, 256) else "tf32" if w["d"] == 32 else "fp32")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)


def timed(name, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc.setdefault(name, []).append(1e3 * (time.perf_counter() - t0))
    return out


acc = {}
for it in range(6):
    packed = timed("dedup", lambda: PackedTexts(None, dev, utf8, offsets), acc)
    text = timed("text_encode", lambda: model.text_encoder.encode_packed(packed), acc)
    h = timed("input_proj", lambda: _native.linear(x, model.input_proj.weight, model.input_proj.bias, relu=True), acc)
    g = timed("graph_build", lambda: _native.Graph(ei, packed.rel_ids, w["N"], packed.num_unique, w["d"]), acc)
    for l in range(w["L"]):
        wts = timed("weight_gen", lambda: model.weight_generators[l](text), acc)
        ln = model.layer_norms[l]
        h = timed("mp_layer", lambda: g.mp_layer(h, wts["W_msg"], wts["W_self"], wts["bias"], ln.weight, ln.bias,
                                                 1e-5, PREC)[0], acc)
    timed("graph_free", lambda: g.__del__(), acc)
    t0 = time.perf_counter()
    out = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, w["N"]))
    torch.cuda.synchronize()
    acc.setdefault("whole_forward", []).append(1e3 * (time.perf_counter() - t0))
for k, v in acc.items():
    v = v[2 * (len(v) // 6):]   # drop the first two iterations
    print(f"{k:14s} {sum(v) / len(v) * (len(v) / 4):8.3f} ms per forward   ({len(v) // 4} calls x {sum(v) / len(v):.3f})")

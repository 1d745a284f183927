"""The literal drop-in call at bench scale: model(node_features, edge_index, List[str]) with a 16M-element Python
list (BASELINE config 3), beside forward_packed on pre-packed strings.  Reports where the host time goes.
    python tools/dropin_call.py [c3]        GPU box only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_grad_enabled(False)
import bench  # noqa: E402
from graph_hypernetwork_forge import _text  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0")
model = bench.build_model(w, dev, "f16")
x, ei, rel, utf8, offsets = bench.make_device_inputs(w, dev)
names = [f"relation_{r:05d}" for r in range(w["R"])]
t0 = time.perf_counter()
texts = [names[r] for r in rel.tolist()]
print(f"building the List[str] of {len(texts)} entries (the caller's cost, not timed below): {time.perf_counter() - t0:.2f} s")


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t) / n, out


ms_packed, a = timed(lambda: model.forward_packed(x, ei, utf8, offsets))
ms_list, b = timed(lambda: model(x, ei, texts))
t = time.perf_counter(); objs, emap = _text.collapse_by_identity(texts); ms_c = 1e3 * (time.perf_counter() - t)
saved, _text._pyhost = _text._pyhost, False
t = time.perf_counter(); _text.collapse_by_identity(texts); ms_np = 1e3 * (time.perf_counter() - t)
_text._pyhost = saved
print(f"forward_packed (strings packed on the device)      : {ms_packed:8.2f} ms")
print(f"model(x, edge_index, List[str]) - the drop-in call  : {ms_list:8.2f} ms   max|diff| {float((a - b).abs().max()):.2e}")
print(f"  of which list -> distinct objects + edge map in C : {ms_c:8.2f} ms   ({len(objs)} distinct objects)")
print(f"  (the numpy formulation it replaces                : {ms_np:8.2f} ms)")
import os  # noqa: E402
buf = np.zeros(len(texts), dtype=np.int32)
for th in (1, 2, 4, 8, 16):
    t = time.perf_counter()
    for _ in range(3):
        _text.collapse_by_identity(texts, out=buf, threads=th)
    print(f"  C pass with {th:2d} thread(s) into a touched buffer       : {1e3 * (time.perf_counter() - t) / 3:8.2f} ms"
          f"   ({os.cpu_count()} host cores)")
print(f"  edge map host -> device: {emap.nbytes / 1e6:.0f} MB")

"""Layout probe for the f16 contraction: one tile, one relation, unit-vector weights (GPU box only).
    python tools/debug_f16.py
With h[u][k] = k + 1 for every node and W_msg = e_{k0} e_{n0}^T, upd[e][n] must be (k0 + 1) at n == n0 and 0 elsewhere.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

from graph_hypernetwork_forge import _native  # noqa: E402

dev = torch.device("cuda:0")
N, E, d = 256, 128, 128
src = torch.arange(E, device=dev).flip(0) + 100
dst = torch.arange(E, device=dev)
ei = torch.stack([src, dst]).long()
rel = torch.zeros(E, dtype=torch.int32, device=dev)
g = _native.Graph(ei, rel, N, 1, d)
h = (torch.arange(d, device=dev, dtype=torch.float32) + 1).repeat(N, 1).contiguous()
ones, zeros = torch.ones(d, device=dev), torch.zeros(d, device=dev)
prec = _native.precision_code(sys.argv[1] if len(sys.argv) > 1 else "f16")


def probe(which, k0, n0):
    Wm = torch.zeros(1, d, d, device=dev)
    Ws = torch.zeros(1, d, d, device=dev)
    (Wm if which == "msg" else Ws)[0, k0, n0] = 1.0
    b = torch.zeros(1, d, device=dev)
    _, upd = g.mp_layer(h, Wm, Ws, b, ones, zeros, 1e-5, prec, want_upd=True)
    torch.cuda.synchronize()
    nz = upd[:E].nonzero()
    rows = sorted(set(nz[:, 0].tolist()))
    cols = sorted(set(nz[:, 1].tolist()))
    vals = sorted(set(upd[:E][upd[:E] != 0].tolist()))
    print(f"{which} k0={k0:3d} n0={n0:3d}: rows {len(rows)} (first {rows[:4]}) cols {cols[:8]} vals {vals[:8]}"
          f"   expect col {n0} val {k0 + 1}")


for which in ("msg", "self"):
    for k0, n0 in ((0, 0), (1, 0), (0, 1), (2, 0), (8, 0), (9, 3), (64, 0), (65, 33), (127, 127), (5, 37)):
        probe(which, k0, n0)

# per-row check: h distinct per node, identity weights
h2 = torch.randn(N, d, device=dev).half().float()
eye = torch.eye(d, device=dev).unsqueeze(0).contiguous()
zero = torch.zeros(1, d, d, device=dev)
b = torch.zeros(1, d, device=dev)
_, upd = g.mp_layer(h2, eye, zero, b, ones, zeros, 1e-5, prec, want_upd=True)
print("identity msg: max|upd - h[src]| =", float((upd[:E] - h2[src]).abs().max()))
_, upd = g.mp_layer(h2, zero, eye, b, ones, zeros, 1e-5, prec, want_upd=True)
print("identity self: max|upd - h[dst]| =", float((upd[:E] - h2[dst]).abs().max()))

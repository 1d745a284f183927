"""Small forwards through every kernel family with awkward geometry (partially filled tiles, units ending mid-tile,
a short last super-block) - the case to put under `compute-sanitizer --tool memcheck` where that tool is available
(it is closed on the round-1 pool; the run without it compares the one-call and the staged forward).  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_grad_enabled(False)   # inference tool: no autograd graph

from graph_hypernetwork_forge import HyperGNN, ToyKnowledgeGraph  # noqa: E402
from oracle import hypergnn_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
kg = ToyKnowledgeGraph(feat_dim=16)
torch.manual_seed(0)
toy = HyperGNN(64, 16, 32, 2).eval().to(dev)
for prec in ("fp32", "tf32"):
    toy.precision = prec
    out = toy(kg.node_features.to(dev), kg.edge_index.to(dev), kg.edge_texts)
    print("toy", prec, tuple(out.shape), float(out.abs().sum()))
# N chosen so that tiles are partially filled, units end mid-tile and the last super-block is short
for d, prec, N, E, R in ((128, "f16", 3001, 20011, 13), (128, "tf32", 3001, 20011, 13), (64, "tf32", 1777, 9001, 7),
                         (48, "fp32", 999, 5003, 5)):
    src, dst, rel, names, feats = O.synthetic_kg(N, E, R, 24, seed=d)
    texts = [names[r] for r in rel]
    torch.manual_seed(1)
    m = HyperGNN(32, 24, d, 2, precision=prec).eval().to(dev)
    ei = torch.from_numpy(np.stack([src, dst])).to(dev)
    x = torch.from_numpy(feats).to(dev)
    out = m(x, ei, texts)
    from graph_hypernetwork_forge import _text
    data, offs = _text.pack_utf8(texts)
    out2 = m.forward_packed(x, ei, torch.from_numpy(data.copy()).to(dev), torch.from_numpy(offs).to(dev))
    torch.cuda.synchronize()
    print(d, prec, tuple(out.shape), float(out.abs().sum()), float((out - out2).abs().max()))
print("done")

"""Generate GRADIENT golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_grad_golden.py

For each case: build the reference `HyperGNN` from a seed (optionally with the log-scales raised so the generated
weights matter), run forward on CPU with autograd, take `loss = (out * loss_weight).sum()` with a seeded random
`loss_weight` (a plain `out.sum()` has zero gradient through LayerNorm), call `loss.backward()` and store the
gradients.  Small tensors are stored whole; for tensors above 64k elements a seeded sample of 4096 entries plus the
L2 norm is stored.  Weights are re-created from the seed by the drop-in constructor (same random stream; the
checksum pins that), so the fixtures stay small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference, synthetic  # noqa: E402

SAMPLE = 4096
WHOLE_BELOW = 1 << 16


def grad_payload(name, g):
    g = g.detach().numpy()
    if g.size <= WHOLE_BELOW:
        return {f"grad/{name}": g}
    rng = np.random.default_rng(sum(map(ord, name)))          # a stable per-tensor seed
    idx = rng.choice(g.size, SAMPLE, replace=False)
    return {f"gradsample_idx/{name}": idx, f"gradsample/{name}": g.reshape(-1)[idx],
            f"gradnorm/{name}": np.array(float(np.linalg.norm(g.astype(np.float64))))}


def save_case(name, ref, ctor, x, edge_index, edge_texts, seed, log_scale):
    torch.manual_seed(seed)
    model = ref.HyperGNN(**ctor)
    model.train()                                   # dropout is 0: train() only documents the intent
    with torch.no_grad():
        for gen in model.weight_generators:
            for p in gen.log_scales.values():
                p.fill_(log_scale)
    checksum = np.array([float(v.double().abs().sum()) for v in model.state_dict().values()])
    x = x.clone().requires_grad_(True)
    out = model(x, edge_index, edge_texts)
    loss_w = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 1000))
    loss = (out * loss_w).sum()
    loss.backward()
    payload = {"ctor": np.array(repr(ctor)), "seed": np.array(seed), "log_scale": np.array(log_scale),
               "node_features": x.detach().numpy(), "edge_index": edge_index.numpy(),
               "edge_texts": np.array(edge_texts, dtype=object), "loss_weight": loss_w.numpy(),
               "loss": np.array(float(loss)), "out": out.detach().numpy(), "param_checksum": checksum}
    payload.update(grad_payload("node_features", x.grad))
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        payload.update(grad_payload(k, p.grad))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, loss = {float(loss):.6f}")


def main():
    ref = load_reference()
    kg = ref.ToyKnowledgeGraph(feat_dim=16)
    # the toy graph of the reference's own training tests (tests/test_hypergnn.py:183-226), hidden 32
    save_case("grad_toy", ref, dict(text_dim=64, node_feat_dim=16, hidden_dim=32, num_layers=2),
              kg.node_features, kg.edge_index, list(kg.edge_texts), seed=0, log_scale=-1.0)
    # hidden_dim not a multiple of 32, multi-edges/self-edges/isolated nodes, empty and non-ASCII strings
    texts = ["", "a", "é€a", "éa", "ëa", "knows", "knows ", "a", "", "\x7f", "日本語", "knows"]
    ei = torch.tensor([[0, 1, 2, 2, 3, 3, 4, 0, 5, 5, 1, 0],
                       [1, 1, 1, 2, 0, 0, 4, 1, 6, 6, 0, 1]], dtype=torch.long)
    x = torch.randn(9, 8, generator=torch.Generator().manual_seed(7))
    save_case("grad_edge_cases_d24", ref, dict(text_dim=16, node_feat_dim=8, hidden_dim=24, num_layers=2),
              x, ei, texts, seed=3, log_scale=-0.5)
    # synthetic, hidden 64 (tf32 path) and hidden 128 (f16 path), three layers
    x, ei, texts = synthetic(300, 2500, 13, 24, seed=11)
    save_case("grad_synth_d64", ref, dict(text_dim=32, node_feat_dim=24, hidden_dim=64, num_layers=3),
              x, ei, texts, seed=11, log_scale=-1.0)
    x, ei, texts = synthetic(400, 3000, 17, 32, seed=12)
    save_case("grad_synth_d128", ref, dict(text_dim=64, node_feat_dim=32, hidden_dim=128, num_layers=2),
              x, ei, texts, seed=12, log_scale=-1.5)


if __name__ == "__main__":
    main()

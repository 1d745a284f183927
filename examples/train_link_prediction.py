"""End-to-end use of the drop-in on one B200: embed a small knowledge graph, score links, train a few steps with a
margin loss, then embed a graph that uses a relation text the model has never seen (zero-shot).

    python examples/train_link_prediction.py

The calls are the reference's (`HyperGNN(...)(node_features, edge_index, edge_texts)`, `score_triple`, a torch
optimiser); the only additions are `.to("cuda")` and `score_edges`, which scores (head, tail) pairs without
materialising the two gathered embedding matrices.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-hypernetwork-forge_b200"))
import torch  # noqa: E402

from graph_hypernetwork_forge import HyperGNN, ToyKnowledgeGraph  # noqa: E402


def main(steps: int = 20, seed: int = 0, verbose: bool = True):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    kg = ToyKnowledgeGraph(feat_dim=16)
    x, ei = kg.node_features.to(dev), kg.edge_index.to(dev)
    model = HyperGNN(text_dim=64, node_feat_dim=16, hidden_dim=32, num_layers=2).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    src, dst = ei
    losses = []
    for step in range(steps):
        model.train()
        opt.zero_grad()
        embs = model(x, ei, kg.edge_texts)
        pos = model.score_edges(embs, src, dst)
        neg = model.score_edges(embs, src, dst[torch.randperm(dst.numel(), device=dev)])
        loss = torch.clamp(1.0 - pos + neg, min=0.0).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        if verbose and (step + 1) % 5 == 0:
            print(f"step {step + 1:3d}  loss {losses[-1]:.4f}")
    model.eval()
    with torch.no_grad():
        texts = list(kg.edge_texts)
        texts[0] = "is a distant cousin of"                    # a relation text never seen in training
        zero_shot = model(x, ei, texts)
    if verbose:
        print("zero-shot embedding norms:", [round(v, 3) for v in zero_shot.norm(dim=1).tolist()])
    return losses, zero_shot


if __name__ == "__main__":
    main()

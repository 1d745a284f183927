"""Differentiable CPU/GPU oracle for the HyperGNN path.  TEST INFRASTRUCTURE ONLY.

A torch restatement (any dtype, normally float64) of the reference algorithm
(danieleschmidt/Graph-Hypernetwork-Forge, `models/hypergnn.py` = "HG", `models/weight_generator.py` = "WG") in the
memory-feasible closed form of `oracle/hypergnn_oracle.py:message_passing`, written with out-of-place torch ops so
that `torch.autograd` yields the GRADIENTS the gradient kernels are compared with.  Only `tests/` may import it;
the product package never does.

Parity status: PINNED.  `tests/golden/grad_*.npz` hold parameter and input gradients of the unmodified reference
(`tests/golden/make_grad_golden.py`, run in the build container); `tests/test_oracle_golden.py` checks this file's
forward and its autograd gradients against them.
"""
from __future__ import annotations

import torch

from .hypergnn_oracle import tokenize


def text_encode(unique_texts, params):
    """HG:73-81: tanh(mean_i Emb[id_i] @ Wp^T + bp), one row per string."""
    emb, w, b = params["text_encoder.char_emb.weight"], params["text_encoder.proj.0.weight"], \
        params["text_encoder.proj.0.bias"]
    rows = []
    for t in unique_texts:
        ids = torch.as_tensor(tokenize(t), dtype=torch.long, device=emb.device)
        rows.append(torch.tanh(emb[ids].mean(dim=0) @ w.T + b))
    return torch.stack(rows) if rows else emb.new_zeros((0, w.shape[0]))


def _mlp(x, params, prefix, dropout=0.0):
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in params if k.startswith(prefix) and k.endswith(".weight")})
    for n, i in enumerate(idx):                                     # WG:97-107: Linear -> ReLU -> Dropout, last Linear bare
        x = x @ params[f"{prefix}{i}.weight"].T + params[f"{prefix}{i}.bias"]
        if n + 1 < len(idx):
            x = torch.relu(x)
            if dropout > 0.0:
                x = torch.nn.functional.dropout(x, dropout)
    return x


def weight_generator(text_emb, params, prefix, d_in, d_out, dropout=0.0):
    """WG:120-143 for a batch [U, T]."""
    out = {}
    for name, shape in (("W_msg", (d_in, d_out)), ("W_self", (d_in, d_out)), ("bias", (d_out,))):
        flat = _mlp(text_emb, params, f"{prefix}generators.{name}.", dropout)
        out[name] = flat.view(text_emb.shape[0], *shape) * params[f"{prefix}log_scales.{name}"].exp()
    return out


def message_passing(h, src, dst, rel, W_msg, W_self, bias):
    """HG:160-230 in closed form: (sum of messages + h_v @ sum of W_self[r_e]) / max(indeg, 1)."""
    N = h.shape[0]
    acc = torch.zeros_like(h)
    for r in torch.unique(rel).tolist():
        idx = (rel == r).nonzero().squeeze(1)
        m = h[src[idx]] @ W_msg[r] + bias[r] + h[dst[idx]] @ W_self[r]
        acc = acc.index_add(0, dst[idx], m)
    cnt = torch.bincount(dst, minlength=N).clamp(min=1).to(h.dtype)
    return acc / cnt.unsqueeze(1)


def hypergnn_forward(params, node_features, edge_index, rel_ids, unique_texts, hidden_dim, num_layers, eps=1e-5,
                     taps=None, dropout=0.0):
    """HG:236-298.  `params`: reference state_dict keys -> tensors (requires_grad where gradients are wanted).
    `dropout` > 0 = training mode with that rate (HG:293-294 and the generator MLPs' Dropout modules, WG:103-104),
    drawing from torch's generator in the reference's call order."""
    h = torch.relu(node_features @ params["input_proj.weight"].T + params["input_proj.bias"])      # HG:261
    te = text_encode(unique_texts, params)
    src, dst = edge_index[0], edge_index[1]
    for l in range(num_layers):
        w = weight_generator(te, params, f"weight_generators.{l}.", hidden_dim, hidden_dim, dropout) \
            if len(unique_texts) else None
        upd = message_passing(h, src, dst, rel_ids, w["W_msg"], w["W_self"], w["bias"]) if w else torch.zeros_like(h)
        if taps is not None:
            taps[f"upd.{l}"] = upd
        x = torch.relu(upd + h)                                                                   # HG:289-291
        if dropout > 0.0:
            x = torch.nn.functional.dropout(x, dropout)                                           # HG:293-294
        h = torch.nn.functional.layer_norm(x, (hidden_dim,), params[f"layer_norms.{l}.weight"],
                                           params[f"layer_norms.{l}.bias"], eps)                  # HG:296
    return h

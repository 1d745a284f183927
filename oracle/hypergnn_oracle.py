"""CPU oracle for the HyperGNN forward path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm
(danieleschmidt/Graph-Hypernetwork-Forge, `graph_hypernetwork_forge/models/hypergnn.py`
= "HG", `.../models/weight_generator.py` = "WG").  It is the checker the CUDA
path is compared with.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product package never does.

Parity status: PINNED.  `tests/golden/*.npz` hold inputs, weights and outputs of
the unmodified reference run in the build container
(`tests/golden/make_golden.py`); `tests/test_oracle_golden.py` checks every
function here against them.

Everything is stated twice where it matters:
  * `message_passing_literal`  follows HG:160-230 line by line (materialises the
    per-edge weight tensors, so small graphs only);
  * `message_passing`          is the memory-feasible closed form
    used at sizes the literal form cannot allocate.
Arithmetic dtype is a parameter (float32 to mirror the reference, float64 to
get a tighter yardstick for tensor-core paths).
"""
from __future__ import annotations

import numpy as np

ASCII_VOCAB = 128  # HG:55


# --------------------------------------------------------------------------
# integer work: dedup, tokenisation, degrees, the relation-grouped edge order
# --------------------------------------------------------------------------
def dedup_texts(edge_texts):
    """HG:264-268.  Unique strings in first-occurrence order + per-edge ids."""
    unique = list(dict.fromkeys(edge_texts))
    index = {t: i for i, t in enumerate(unique)}
    ids = np.fromiter((index[t] for t in edge_texts), dtype=np.int64, count=len(edge_texts))
    return unique, ids


def tokenize(text):
    """HG:66-71.  Code points clamped to 127; the empty string becomes [0]."""
    codes = [min(ord(c), ASCII_VOCAB - 1) for c in text]
    if not codes:
        codes = [0]
    return np.asarray(codes, dtype=np.int64)


def pack_utf8(texts):
    """Host packing used at the C-ABI boundary: UTF-8 bytes + int64 offsets.

    UTF-8 is injective on Python strings (surrogates are passed through), so
    equality of byte strings is equality of the reference's dict keys.
    """
    blobs = [t.encode("utf-8", "surrogatepass") for t in texts]
    offsets = np.zeros(len(blobs) + 1, dtype=np.int64)
    if blobs:
        offsets[1:] = np.cumsum([len(b) for b in blobs])
    data = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy()
    return data, offsets


def tokenize_utf8(data, start, end):
    """Same tokens as `tokenize`, recovered from UTF-8 bytes.

    A code point >= 128 is one lead byte (>= 0xC0) plus continuation bytes
    (0x80..0xBF); it must yield the single token 127.  ASCII bytes map to
    themselves.  This is the rule the CUDA text encoder implements.
    """
    b = data[start:end]
    keep = (b & 0xC0) != 0x80
    tok = np.where(b[keep] < 128, b[keep], 127).astype(np.int64)
    if tok.size == 0:
        tok = np.zeros(1, dtype=np.int64)
    return tok


def dedup_utf8(data, offsets):
    """Dedup over packed byte strings; must equal `dedup_texts` on the originals."""
    seen = {}
    E = len(offsets) - 1
    ids = np.empty(E, dtype=np.int64)
    first = []
    raw = data.tobytes()
    for e in range(E):
        key = raw[offsets[e]:offsets[e + 1]]
        j = seen.get(key)
        if j is None:
            j = len(first)
            seen[key] = j
            first.append(e)
        ids[e] = j
    return ids, np.asarray(first, dtype=np.int64)


def in_degree(dst, num_nodes):
    """HG:208-211: scatter_add of ones by destination (multi-edges count)."""
    return np.bincount(np.asarray(dst, dtype=np.int64), minlength=num_nodes).astype(np.int64)


def edge_order(dst, rel, num_nodes, num_rel, sb_nodes, unit_edges, dst_lo=0, dst_hi=None):
    """The edge order the CUDA graph build must reproduce bit-exactly.

    Edges whose destination lies in [dst_lo, dst_hi) are kept and sorted
    (stably, ties by edge id) by (super-block of dst, relation, local dst).
    Each (super-block, relation) group is cut into work units of at most
    `unit_edges` edges.  Returns (perm, unit_start, unit_count, unit_rel).
    """
    dst = np.asarray(dst, dtype=np.int64)
    rel = np.asarray(rel, dtype=np.int64)
    if dst_hi is None:
        dst_hi = num_nodes
    keep = np.nonzero((dst >= dst_lo) & (dst < dst_hi))[0]
    dl = dst[keep] - dst_lo
    key = ((dl // sb_nodes) * num_rel + rel[keep]) * sb_nodes + (dl % sb_nodes)
    order = np.argsort(key, kind="stable")
    perm = keep[order]
    group = key[order] // sb_nodes
    starts, counts, rels = [], [], []
    if perm.size:
        bounds = np.nonzero(np.diff(group))[0] + 1
        gs = np.concatenate([[0], bounds])
        ge = np.concatenate([bounds, [perm.size]])
        for s, e in zip(gs, ge):
            r = int(group[s] % num_rel)
            for u in range(s, e, unit_edges):
                starts.append(u)
                counts.append(min(unit_edges, e - u))
                rels.append(r)
    return (perm.astype(np.int64), np.asarray(starts, dtype=np.int64),
            np.asarray(counts, dtype=np.int64), np.asarray(rels, dtype=np.int64))


# --------------------------------------------------------------------------
# floating point work
# --------------------------------------------------------------------------
def _linear(x, w, b):
    return x @ w.T + b


def text_encode(unique_texts, char_emb, proj_w, proj_b, dtype=np.float32):
    """HG:73-81: tanh(mean_i Emb[id_i] @ Wp^T + bp), one row per unique string."""
    char_emb = char_emb.astype(dtype)
    proj_w = proj_w.astype(dtype)
    proj_b = proj_b.astype(dtype)
    out = np.empty((len(unique_texts), proj_w.shape[0]), dtype=dtype)
    for i, t in enumerate(unique_texts):
        pooled = char_emb[tokenize(t)].mean(axis=0, dtype=dtype)
        out[i] = np.tanh(_linear(pooled, proj_w, proj_b))
    return out


def _mlp_linears(params, prefix):
    """Linear layers of one nn.Sequential in index order (indices shift when the
    reference inserts Dropout modules, WG:103-104)."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in params
                  if k.startswith(prefix) and k.endswith(".weight")})
    return [(params[f"{prefix}{i}.weight"], params[f"{prefix}{i}.bias"]) for i in idx]


def weight_generator(text_emb, params, prefix, d_in, d_out, dtype=np.float32):
    """WG:120-143.  `params` maps reference state_dict keys to arrays; `prefix`
    is e.g. "weight_generators.0." ("" for a standalone generator)."""
    x = np.asarray(text_emb, dtype=dtype)
    single = x.ndim == 1
    if single:
        x = x[None, :]
    out = {}
    for name, shape in (("W_msg", (d_in, d_out)), ("W_self", (d_in, d_out)), ("bias", (d_out,))):
        layers = _mlp_linears(params, f"{prefix}generators.{name}.")
        a = x
        for li, (w, b) in enumerate(layers):
            a = _linear(a, w.astype(dtype), b.astype(dtype))
            if li + 1 < len(layers):
                a = np.maximum(a, 0)
        scale = np.exp(params[f"{prefix}log_scales.{name}"].astype(dtype))
        w_full = a.reshape((x.shape[0],) + shape) * scale
        out[name] = w_full[0] if single else w_full
    return out


def message_passing_literal(h, src, dst, W_msg_e, W_self_e, bias_e):
    """HG:160-230 as written: per-edge weights [E,d,d] are inputs."""
    N, d = h.shape
    d_out = W_msg_e.shape[-1]
    msg = np.einsum("ei,eio->eo", h[src], W_msg_e) + bias_e
    agg = np.zeros((N, d_out), dtype=h.dtype)
    np.add.at(agg, dst, msg)
    cnt = np.maximum(np.bincount(dst, minlength=N).astype(h.dtype), 1)[:, None]
    agg = agg / cnt
    W_self_agg = np.zeros((N, d, d_out), dtype=h.dtype)
    np.add.at(W_self_agg, dst, W_self_e)
    W_self_agg = W_self_agg / cnt[:, :, None]
    self_out = np.einsum("ni,nio->no", h, W_self_agg)
    return agg + self_out


def message_passing(h, src, dst, rel, W_msg, W_self, bias, num_nodes=None):
    """Closed form of HG:160-230, relation-grouped:

        upd_v = (1/c_v) * sum_{e:(u->v), r_e} (h_u W_msg[r_e] + bias[r_e] + h_v W_self[r_e])

    with c_v = max(in-degree, 1).  No [E,d,d] or [N,d,d] intermediates.
    """
    N = h.shape[0] if num_nodes is None else num_nodes
    d_out = W_msg.shape[-1]
    acc = np.zeros((N, d_out), dtype=h.dtype)
    order = np.argsort(rel, kind="stable")
    rs = rel[order]
    bounds = np.concatenate([[0], np.nonzero(np.diff(rs))[0] + 1, [rs.size]]) if rs.size else [0]
    for s, e in zip(bounds[:-1], bounds[1:]):
        r = int(rs[s])
        idx = order[s:e]
        out = h[src[idx]] @ W_msg[r] + h[dst[idx]] @ W_self[r] + bias[r]
        # destinations repeat inside a relation: reduce sorted runs, then add once
        o2 = np.argsort(dst[idx], kind="stable")
        dsorted = dst[idx][o2]
        first = np.concatenate([[0], np.nonzero(np.diff(dsorted))[0] + 1])
        acc[dsorted[first]] += np.add.reduceat(out[o2], first, axis=0)
    cnt = np.maximum(np.bincount(dst, minlength=N), 1).astype(h.dtype)[:, None]
    return acc / cnt


def layer_norm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm over the last dim: biased variance, affine (HG:152-154)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * weight + bias


def hypergnn_forward(params, node_features, edge_index, edge_texts, hidden_dim, num_layers,
                     dtype=np.float32, literal=False, taps=None):
    """HG:236-298 in eval mode (dropout inactive).  `taps`, when a dict, receives
    edge_rel_ids, in_degree, text_embs and per-layer W_msg/W_self/bias/upd/h."""
    P = {k: np.asarray(v) for k, v in params.items()}
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    if src.shape[0] != len(edge_texts):
        raise ValueError(
            f"edge_index has {src.shape[0]} edges but edge_texts has {len(edge_texts)} entries")
    x = np.asarray(node_features, dtype=dtype)
    N = x.shape[0]
    h = np.maximum(_linear(x, P["input_proj.weight"].astype(dtype), P["input_proj.bias"].astype(dtype)), 0)
    unique, rel = dedup_texts(edge_texts)
    text_embs = text_encode(unique, P["text_encoder.char_emb.weight"],
                            P["text_encoder.proj.0.weight"], P["text_encoder.proj.0.bias"], dtype)
    if taps is not None:
        taps["edge_rel_ids"] = rel
        taps["in_degree"] = in_degree(dst, N)
        taps["text_embs"] = text_embs
        taps["h0"] = h
    for l in range(num_layers):
        w = weight_generator(text_embs, P, f"weight_generators.{l}.", hidden_dim, hidden_dim, dtype)
        if literal:
            upd = message_passing_literal(h, src, dst, w["W_msg"][rel], w["W_self"][rel], w["bias"][rel])
        else:
            upd = message_passing(h, src, dst, rel, w["W_msg"], w["W_self"], w["bias"])
        h = layer_norm(np.maximum(upd + h, 0), P[f"layer_norms.{l}.weight"].astype(dtype),
                       P[f"layer_norms.{l}.bias"].astype(dtype)).astype(dtype)
        if taps is not None:
            taps[f"W_msg.{l}"], taps[f"W_self.{l}"], taps[f"bias.{l}"] = w["W_msg"], w["W_self"], w["bias"]
            taps[f"upd.{l}"] = upd
            taps[f"h.{l}"] = h
    return h


def receptive_fields(src, dst, sample, num_layers, num_nodes):
    """Node sets a sampled forward needs: fields[L] = the sampled nodes, fields[l] = fields[l+1] plus every
    in-neighbour of fields[l+1] (HG:201-219 reads h[src] and h[dst] of every in-edge).  Also returns, per layer,
    the ids of ALL in-edges of fields[l+1] (so in-degrees of those nodes are complete)."""
    order = np.argsort(dst, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=rowptr[1:])
    fields = [None] * (num_layers + 1)
    edges = [None] * num_layers
    fields[num_layers] = np.unique(np.asarray(sample, dtype=np.int64))
    for l in range(num_layers - 1, -1, -1):
        nodes = fields[l + 1]
        ids = np.concatenate([order[rowptr[v]:rowptr[v + 1]] for v in nodes]) if nodes.size else np.zeros(0, np.int64)
        edges[l] = np.sort(ids)
        fields[l] = np.union1d(nodes, src[ids])
    return fields, edges


def sampled_forward(params, node_features, src, dst, rel, unique_texts, sample, hidden_dim, num_layers,
                    dtype=np.float64, rel_chunk=256, taps=None):
    """HG:236-298 restricted to the receptive field of `sample`: exact (every in-edge of every node whose value
    is needed is included), memory-feasible at BASELINE sizes.  Returns (nodes, h_L[nodes]) with nodes = sorted
    unique sample.  `taps` receives per layer `nodes.l`, `upd.l`, `h.l` (rows of fields[l+1]).
    Weights are generated per chunk of the relations that actually occur (c4: 20k relations x 2 x 256^2 doubles
    would not fit otherwise); each chunk goes through `message_passing` (the pinned closed form)."""
    P = {k: np.asarray(v) for k, v in params.items()}
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    rel = np.asarray(rel, dtype=np.int64)
    N = node_features.shape[0]
    fields, edges = receptive_fields(src, dst, sample, num_layers, N)
    text_embs = text_encode(unique_texts, P["text_encoder.char_emb.weight"], P["text_encoder.proj.0.weight"],
                            P["text_encoder.proj.0.bias"], dtype)
    x = np.asarray(node_features[fields[0]], dtype=dtype)
    h = np.maximum(_linear(x, P["input_proj.weight"].astype(dtype), P["input_proj.bias"].astype(dtype)), 0)
    nodes = fields[0]
    for l in range(num_layers):
        ids = edges[l]
        s, t, r = np.searchsorted(nodes, src[ids]), np.searchsorted(nodes, dst[ids]), rel[ids]
        n = nodes.size
        acc = np.zeros((n, hidden_dim), dtype=dtype)
        present = np.unique(r)
        for c0 in range(0, present.size, rel_chunk):
            chunk = present[c0:c0 + rel_chunk]
            w = weight_generator(text_embs[chunk], P, f"weight_generators.{l}.", hidden_dim, hidden_dim, dtype)
            pick = np.nonzero(np.isin(r, chunk))[0]
            part = message_passing(h, s[pick], t[pick], np.searchsorted(chunk, r[pick]), w["W_msg"], w["W_self"],
                                   w["bias"], num_nodes=n)
            acc += part * np.maximum(np.bincount(t[pick], minlength=n), 1).astype(dtype)[:, None]
        upd = acc / np.maximum(np.bincount(t, minlength=n), 1).astype(dtype)[:, None]
        keep = np.searchsorted(nodes, fields[l + 1])           # rows whose in-edges are complete
        h = layer_norm(np.maximum(upd[keep] + h[keep], 0), P[f"layer_norms.{l}.weight"].astype(dtype),
                       P[f"layer_norms.{l}.bias"].astype(dtype)).astype(dtype)
        nodes = fields[l + 1]
        if taps is not None:
            taps[f"nodes.{l}"], taps[f"upd.{l}"], taps[f"h.{l}"] = nodes, upd[keep], h
    return nodes, h


# --------------------------------------------------------------------------
# synthetic workloads shared by tests and bench
# --------------------------------------------------------------------------
def synthetic_kg(num_nodes, num_edges, num_rel, feat_dim, seed=0, skew=False):
    """Uniform (or Zipf-skewed) synthetic KG: src,dst,rel ids, relation names,
    node features.  Names are `relation_%05d`; first-occurrence order differs
    from numeric order, so dedup ranking is exercised."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, num_nodes, num_edges, dtype=np.int64)
    if skew:
        dst = np.minimum((rng.zipf(1.3, num_edges) - 1) % num_nodes, num_nodes - 1).astype(np.int64)
        rel = np.minimum(rng.zipf(1.5, num_edges) - 1, num_rel - 1).astype(np.int64)
    else:
        dst = rng.integers(0, num_nodes, num_edges, dtype=np.int64)
        rel = rng.integers(0, num_rel, num_edges, dtype=np.int64)
    names = [f"relation_{r:05d}" for r in range(num_rel)]
    feats = rng.standard_normal((num_nodes, feat_dim), dtype=np.float32)
    return src, dst, rel, names, feats

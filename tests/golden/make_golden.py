"""Generate golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports `/root/reference/graph_hypernetwork_forge` under the alias
`ghf_reference`, runs it on CPU (eval, no_grad) and stores inputs, weights and
tapped outputs as small .npz fixtures next to this script.  `/root/reference`
does not exist on the GPU box; the fixtures are what travels.

Taps are taken by wrapping the reference's own methods (no reference code is
modified): `text_encoder.forward`, each `weight_generators[l].forward`, each
`_message_passing` call (pre-residual update) and each `layer_norms[l]`.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF_ROOT = "/root/reference/graph_hypernetwork_forge"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location(
        "ghf_reference", os.path.join(REF_ROOT, "__init__.py"),
        submodule_search_locations=[REF_ROOT])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ghf_reference"] = mod
    spec.loader.exec_module(mod)
    return mod


def tapped_forward(model, x, edge_index, edge_texts):
    taps = {}
    te_fwd = model.text_encoder.forward

    def te(texts, device):
        out = te_fwd(texts, device)
        taps["text_embs"] = out.detach().numpy().copy()
        taps["unique_texts"] = list(texts)
        return out
    model.text_encoder.forward = te

    mp = model._message_passing
    layer = {"i": 0}

    def mp_tap(h, ei, rel_weights):
        out = mp(h, ei, rel_weights)
        taps[f"upd.{layer['i']}"] = out.detach().numpy().copy()
        layer["i"] += 1
        return out
    model._message_passing = mp_tap

    hooks = []
    for l, gen in enumerate(model.weight_generators):
        def hook(_m, _inp, out, l=l):
            for k, v in out.items():
                taps[f"{k}.{l}"] = v.detach().numpy().copy()
        hooks.append(gen.register_forward_hook(hook))
    for l, ln in enumerate(model.layer_norms):
        def hook(_m, _inp, out, l=l):
            taps[f"h.{l}"] = out.detach().numpy().copy()
        hooks.append(ln.register_forward_hook(hook))

    with torch.no_grad():
        out = model(x, edge_index, edge_texts)
    for hk in hooks:
        hk.remove()
    model.text_encoder.forward = te_fwd
    model._message_passing = mp
    taps["out"] = out.numpy().copy()
    return taps


def save_case(name, ref, ctor, x, edge_index, edge_texts, seed, store_params=True, log_scale=None):
    torch.manual_seed(seed)
    model = ref.HyperGNN(**ctor).eval()
    if log_scale is not None:
        with torch.no_grad():
            for gen in model.weight_generators:
                for p in gen.log_scales.values():
                    p.fill_(log_scale)
    taps = tapped_forward(model, x, edge_index, edge_texts)
    unique = list(dict.fromkeys(edge_texts))
    idx = {t: i for i, t in enumerate(unique)}
    payload = {
        "ctor": np.array(repr(ctor)),
        "seed": np.array(seed),
        "log_scale": np.array(np.nan if log_scale is None else log_scale),
        "node_features": x.numpy(),
        "edge_index": edge_index.numpy(),
        "edge_texts": np.array(edge_texts, dtype=object),
        "edge_rel_ids": np.array([idx[t] for t in edge_texts], dtype=np.int64),
        "in_degree": np.bincount(edge_index[1].numpy(), minlength=x.shape[0]).astype(np.int64),
    }
    for k, v in taps.items():
        if k == "unique_texts":
            payload[k] = np.array(v, dtype=object)
        elif store_params or not k.startswith(("W_msg", "W_self")):
            payload["tap/" + k] = v
    if store_params:
        for k, v in model.state_dict().items():
            payload["param/" + k] = v.numpy()
    else:
        # weights are re-created from `seed` by the drop-in module's constructor;
        # a checksum pins that the two constructors draw the same stream
        payload["param_checksum"] = np.array(
            [float(v.double().abs().sum()) for v in model.state_dict().values()])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    print(f"{name}: {os.path.getsize(path)/1e6:.2f} MB, out|sum| = {np.abs(taps['out']).sum():.4f}")


def synthetic(N, E, R, F, seed, texts=None):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, N, (2, E), generator=g)
    rel = torch.randint(0, R, (E,), generator=g)
    x = torch.randn(N, F, generator=g)
    names = texts or [f"relation_{r:05d}" for r in range(R)]
    return x, ei, [names[r] for r in rel.tolist()]


def main():
    ref = load_reference()

    # c1: BASELINE config 1, verbatim (ToyKnowledgeGraph, text_dim 64, feat 16, hidden 32, 2 layers)
    kg = ref.ToyKnowledgeGraph(feat_dim=16)
    save_case("toy_c1", ref, dict(text_dim=64, node_feat_dim=16, hidden_dim=32, num_layers=2),
              kg.node_features, kg.edge_index, list(kg.edge_texts), seed=0)

    # edge cases: empty string, non-ASCII code points (clamped to 127), strings that
    # differ only above 127 (distinct relations, identical embeddings), multi-edges,
    # self-edges, isolated nodes, hidden_dim not a multiple of 32
    texts = ["", "a", "é€a", "éa", "ëa", "knows", "knows ", "a", "", "\x7f", "日本語", "knows"]
    ei = torch.tensor([[0, 1, 2, 2, 3, 3, 4, 0, 5, 5, 1, 0],
                       [1, 1, 1, 2, 0, 0, 4, 1, 6, 6, 0, 1]], dtype=torch.long)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(9, 8, generator=g)   # nodes 7, 8 isolated; 3 has no in-edges
    save_case("edge_cases", ref, dict(text_dim=32, node_feat_dim=8, hidden_dim=16, num_layers=2),
              x, ei, texts, seed=1)
    save_case("edge_cases_d24", ref, dict(text_dim=12, node_feat_dim=8, hidden_dim=24, num_layers=1,
                                          char_emb_dim=8),
              x, ei, texts, seed=2)

    # small synthetic, three layers; and the same with log_scales := 0 so that the
    # generated weights are O(1) and `upd` is not drowned by the residual
    x, ei, et = synthetic(200, 1500, 23, 24, seed=3)
    save_case("synth_small", ref, dict(text_dim=16, node_feat_dim=24, hidden_dim=32, num_layers=3),
              x, ei, et, seed=3)
    save_case("synth_small_scale1", ref, dict(text_dim=16, node_feat_dim=24, hidden_dim=32, num_layers=3),
              x, ei, et, seed=3, log_scale=0.0)

    # hidden 64 / 128 (the tensor-core shapes): weights are too large to store, so only
    # inputs + outputs + a parameter checksum are kept (weights re-drawn from the seed)
    x, ei, et = synthetic(300, 4000, 40, 32, seed=4)
    save_case("synth_d64", ref, dict(text_dim=32, node_feat_dim=32, hidden_dim=64, num_layers=2),
              x, ei, et, seed=4, store_params=False)
    x, ei, et = synthetic(256, 3000, 37, 48, seed=5)
    save_case("synth_d128", ref, dict(text_dim=64, node_feat_dim=48, hidden_dim=128, num_layers=2),
              x, ei, et, seed=5, store_params=False)
    save_case("synth_d128_scale1", ref, dict(text_dim=64, node_feat_dim=48, hidden_dim=128, num_layers=2),
              x, ei, et, seed=5, store_params=False, log_scale=0.0)

    # standalone WeightGenerator (WG:120-143): non-square, 1-D input, num_hidden=0
    out = {}
    for tag, kw in (("nonsquare", dict(text_dim=16, d_in=8, d_out=24, hidden_dim=64)),
                    ("depth0", dict(text_dim=16, d_in=4, d_out=4, num_hidden=0)),
                    ("default", dict(text_dim=32, d_in=16, d_out=16))):
        torch.manual_seed(11)
        gen = ref.WeightGenerator(**kw).eval()
        g = torch.Generator().manual_seed(12)
        emb = torch.randn(5, kw["text_dim"], generator=g)
        with torch.no_grad():
            wb = gen(emb)
            w1 = gen(emb[0])
        out[f"{tag}/ctor"] = np.array(repr(kw))
        out[f"{tag}/emb"] = emb.numpy()
        for k, v in gen.state_dict().items():
            out[f"{tag}/param/{k}"] = v.numpy()
        for k in wb:
            out[f"{tag}/batched/{k}"] = wb[k].numpy()
            out[f"{tag}/single/{k}"] = w1[k].numpy()
    path = os.path.join(HERE, "weight_generator.npz")
    np.savez_compressed(path, **out)
    print(f"weight_generator: {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()

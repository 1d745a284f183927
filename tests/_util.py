"""Test helpers: golden fixtures, model reconstruction, tolerances."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORWARD_CASES = ["toy_c1", "edge_cases", "edge_cases_d24", "synth_small", "synth_small_scale1", "synth_d64",
                 "synth_d128", "synth_d128_scale1"]

# Tolerances (stated once, used everywhere):
#   fp32 path: rtol 1e-5 + atol 1e-6 on every tap (north_star: "about 1e-5 for an fp32 path").
#   tf32 path: the contraction rounds/truncates operands to 10 mantissa bits, so the update `upd`
#   carries ~1e-3 relative error measured against max|upd|; everything downstream inherits it scaled
#   by |upd|/|h| (1e-3 at init scale, O(1) with log_scales = 0).
FP32_RTOL, FP32_ATOL = 1e-5, 1e-6
TF32_UPD_REL = 3e-3          # max|upd - ref| <= TF32_UPD_REL * max|ref|
TF32_H_ATOL_INIT = 2e-5      # final h at init scale (|upd| ~ 1e-3 |h|)
TF32_H_ATOL_SCALE1 = 2e-2    # final h with log_scales = 0 (|upd| ~ |h|; LayerNorm output is O(1))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    case = {"name": name, "ctor": ast.literal_eval(str(z["ctor"])), "seed": int(z["seed"]),
            "log_scale": None if np.isnan(float(z["log_scale"])) else float(z["log_scale"]),
            "node_features": z["node_features"], "edge_index": z["edge_index"],
            "edge_texts": [str(t) for t in z["edge_texts"]], "edge_rel_ids": z["edge_rel_ids"],
            "in_degree": z["in_degree"], "unique_texts": [str(t) for t in z["unique_texts"]],
            "taps": {k[4:]: z[k] for k in z.files if k.startswith("tap/")},
            "params": {k[6:]: z[k] for k in z.files if k.startswith("param/")} or None,
            "param_checksum": z["param_checksum"] if "param_checksum" in z.files else None}
    return case


def build_model(case, device="cpu", precision=None):
    """Drop-in model carrying the reference's weights for this case."""
    from graph_hypernetwork_forge import HyperGNN
    torch.manual_seed(case["seed"])
    model = HyperGNN(**case["ctor"], precision=precision).eval()
    if case["params"] is not None:
        model.load_state_dict({k: torch.from_numpy(v) for k, v in case["params"].items()})
    if case["log_scale"] is not None:
        with torch.no_grad():
            for gen in model.weight_generators:
                for p in gen.log_scales.values():
                    p.fill_(case["log_scale"])
    return model.to(device)


def model_params_numpy(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def assert_close(got, want, rtol, atol, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    bound = atol + rtol * np.abs(want)
    worst = float((err - bound).max()) if err.size else 0.0
    assert worst <= 0, f"{what}: max|err|={err.max():.3e}, max|ref|={np.abs(want).max():.3e}, exceeds by {worst:.3e}"


def assert_rel_to_max(got, want, rel, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got - want).max())
    assert err <= rel * scale, f"{what}: max|err|={err:.3e} > {rel:g} * max|ref|={scale:.3e}"


# ---- gradient fixtures (tests/golden/make_grad_golden.py) ----
GRAD_CASES = ["grad_toy", "grad_edge_cases_d24", "grad_synth_d64", "grad_synth_d128"]


def load_grad_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    case = {"name": name, "ctor": ast.literal_eval(str(z["ctor"])), "seed": int(z["seed"]),
            "log_scale": float(z["log_scale"]), "params": None, "param_checksum": z["param_checksum"],
            "node_features": z["node_features"], "edge_index": z["edge_index"],
            "edge_texts": [str(t) for t in z["edge_texts"]], "loss_weight": z["loss_weight"],
            "loss": float(z["loss"]), "out": z["out"],
            "grads": {k[5:]: z[k] for k in z.files if k.startswith("grad/")},
            "grad_samples": {k[11:]: (z["gradsample_idx/" + k[11:]], z[k], float(z["gradnorm/" + k[11:]]))
                             for k in z.files if k.startswith("gradsample/")}}
    return case


def check_grads(case, grads, rel, what=""):
    """`grads`: name -> array-like (full gradient tensors) against the reference's, each relative to the largest
    reference entry of that tensor (tensors stored as a sample: the sampled entries and the L2 norm)."""
    names = set(case["grads"]) | set(case["grad_samples"])
    missing = names - set(grads)
    assert not missing, f"{what}: no gradient for {sorted(missing)}"
    for k, want in case["grads"].items():
        assert_rel_to_max(np.asarray(grads[k]).reshape(want.shape), want, rel, f"{what} grad {k}")
    for k, (idx, want, norm) in case["grad_samples"].items():
        got = np.asarray(grads[k], dtype=np.float64).reshape(-1)
        assert_rel_to_max(got[idx], want, rel, f"{what} grad {k} (sample)")
        assert abs(float(np.linalg.norm(got)) - norm) <= 2 * rel * norm, f"{what} grad {k}: norm"

"""Test helpers: golden fixtures, model reconstruction, tolerances."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORWARD_CASES = ["toy_c1", "edge_cases", "edge_cases_d24", "synth_small", "synth_small_scale1", "synth_d64",
                 "synth_d128", "synth_d128_scale1"]

# Tolerances (stated once, used everywhere; DESIGN.md section 2 carries the same table):
#   fp32 path: rtol 1e-5 + atol 1e-6 on text embeddings and generated weights (north_star: "about 1e-5 for an fp32
#   path"); rtol 1e-4 / atol 2e-5 on `upd`, `h` (summation order of the per-destination atomics).
#   Tensor-core paths round the operands of the per-edge contraction to an 11-bit significand (fp16 transport with
#   power-of-two scales, or TF32) and accumulate in fp32.  The bounds below are ~3x the LARGEST error measured over
#   the whole GPU suite in round 2 (every comparison is recorded, see RECORDS / gpurun_out/parity_errors.json;
#   the summary is committed as profiles/r02_parity_errors.txt):
#     quantity                         measured max (f16 / tf32)     bound (f16 / tf32)
#     upd.l, relative to max|upd.l|    2.96e-4 / 6.44e-4             1e-3 / 2e-3
#     h.l, out with O(1) weights       1.56e-4 / 2.53e-4 (absolute)  5e-4 / 8e-4
#     h.l, out at init scale (1e-2)    1.9e-6  / 2.9e-6  (absolute)  1e-5
#   (round 1 held these to 3e-3 / 2e-2 / 2e-5.)  At in-degree 6.4 a dropped edge moves upd by ~1/6 of a message,
#   i.e. >= 1e-2 of max|upd| - an order of magnitude above the bound.
FP32_RTOL, FP32_ATOL = 1e-5, 1e-6
TC_UPD_REL = {"f16": 1e-3, "tf32": 2e-3}         # max|upd - ref| <= TC_UPD_REL * max|ref|, every layer
TC_H_ATOL_INIT = 1e-5                            # h.l / out at init scale (|upd| ~ 1e-3 |h|)
TC_H_ATOL_SCALE1 = {"f16": 5e-4, "tf32": 8e-4}   # h.l / out with O(1) generated weights (LayerNorm output is O(1))
RERUN_ATOL = 1.5e-4                              # run-to-run / edge-order differences of the fp32 atomics (measured 4.2e-5)


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


# Every comparison is recorded (measured error next to the bound it was held to); conftest.py prints the table at
# the end of the run and writes it to gpurun_out/parity_errors.json, so the tolerances above can be audited
# against what the kernels actually deliver.
RECORDS = []


def _record(what, err, ref_max, bound, kind):
    RECORDS.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "what": what, "kind": kind,
                    "max_err": float(err), "max_ref": float(ref_max), "bound": float(bound),
                    "used": float(err / bound) if bound > 0 else (0.0 if err == 0 else float("inf"))})


def assert_close(got, want, rtol, atol, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    bound = atol + rtol * np.abs(want)
    worst = float((err - bound).max()) if err.size else 0.0
    if err.size:
        i = int(np.argmax(err - bound))
        _record(what, err.flat[i], np.abs(want).max(), bound.flat[i], f"rtol={rtol:g} atol={atol:g}")
    assert worst <= 0, f"{what}: max|err|={err.max():.3e}, max|ref|={np.abs(want).max():.3e}, exceeds by {worst:.3e}"


def assert_rel_to_max(got, want, rel, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got - want).max()) if got.size else 0.0
    _record(what, err, scale, rel * scale, f"rel-to-max={rel:g}")
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

"""CPU-side checks: drop-in surface, host packing logic, the C-ABI library loads and exports every
symbol include/ghf_b200.h declares, and the product path refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import hypergnn_oracle as O
from _util import build_model, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    from graph_hypernetwork_forge import _native
    header = open(os.path.join(ROOT, "include", "ghf_b200.h")).read()
    declared = set(re.findall(r"\b(ghf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTED_SYMBOLS)
    L = ctypes.CDLL(_native.lib_path())
    for name in declared:
        assert hasattr(L, name), name
    assert _native.lib().ghf_abi_version() == _native.ABI_VERSION == 6


def test_state_dict_and_init_stream_match_reference():
    for name in ("toy_c1", "edge_cases", "synth_small"):
        c = load_case(name)
        from graph_hypernetwork_forge import HyperGNN
        torch.manual_seed(c["seed"])
        sd = HyperGNN(**c["ctor"]).state_dict()
        assert list(sd.keys()) == list(c["params"].keys())      # same names, same order
        for k, v in c["params"].items():
            assert np.array_equal(sd[k].numpy(), v), k          # same random stream


def test_constructor_surface_and_errors():
    from graph_hypernetwork_forge import HyperGNN, WeightGenerator
    from graph_hypernetwork_forge.models import TextEncoder
    m = HyperGNN(text_dim=32, node_feat_dim=16, hidden_dim=16, num_layers=3)
    assert len(m.weight_generators) == 3 and len(m.layer_norms) == 3 and m.num_layers == 3
    assert (m.text_dim, m.node_feat_dim, m.hidden_dim, m.dropout) == (32, 16, 16, 0.0)
    assert m.num_parameters() > 0
    assert m.weight_generators[0].generators["W_msg"][0].out_features == 64   # max(64, 2*text_dim)
    with pytest.raises(ValueError):
        HyperGNN(32, 16, 16, num_layers=0)
    for bad in (dict(text_dim=0, d_in=4, d_out=4), dict(text_dim=4, d_in=-1, d_out=4), dict(text_dim=4, d_in=4, d_out=0)):
        with pytest.raises(ValueError):
            WeightGenerator(**bad)
    g = WeightGenerator(32, 16, 16, hidden_dim=256)
    assert (g.text_dim, g.d_in, g.d_out, g.init_scale) == (32, 16, 16, 0.01)
    assert g.generators["W_msg"][0].out_features == 256
    assert sum("log_scale" in n for n, _ in g.named_parameters()) == 3
    assert list(WeightGenerator(8, 4, 4, num_hidden=0).generators["bias"].state_dict()) == ["0.weight", "0.bias"]
    assert list(WeightGenerator(8, 4, 4, dropout=0.1).generators["bias"].state_dict())[-2:] == ["6.weight", "6.bias"]
    assert TextEncoder.ASCII_VOCAB == 128
    assert TextEncoder(8)._tokenize("é€a", "cpu").tolist() == [127, 127, 97]
    assert TextEncoder(8)._tokenize("", "cpu").tolist() == [0]


def test_no_cpu_fallback():
    c = load_case("toy_c1")
    m = build_model(c)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.from_numpy(c["node_features"]), torch.from_numpy(c["edge_index"]), c["edge_texts"])
    with pytest.raises(ValueError):   # count mismatch is checked before anything touches a device
        m(torch.from_numpy(c["node_features"]), torch.from_numpy(c["edge_index"]), c["edge_texts"][:-1])
    with pytest.raises(RuntimeError):
        m.weight_generators[0](torch.randn(64))
    with pytest.raises(RuntimeError):
        m.text_encoder(["knows"], torch.device("cpu"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graph-hypernetwork-forge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle.edge_order()", ""), os.path.join(dirpath, f)


def test_text_packing_matches_oracle():
    from graph_hypernetwork_forge import _text
    texts = ["", "a", "é€a", "knows", "日本語", "a", "\U0001F600", "knows "]
    d1, o1 = _text.pack_utf8(texts)
    d2, o2 = O.pack_utf8(texts)
    assert np.array_equal(d1, d2) and np.array_equal(o1, o2)
    d1, o1 = _text.pack_utf8(["abc", "", "de"])
    assert bytes(d1) == b"abcde" and o1.tolist() == [0, 3, 3, 5]
    assert _text.pack_utf8([])[1].tolist() == [0]


def test_identity_collapse_composes_to_first_occurrence_ids():
    """collapse_by_identity + content dedup of the collapsed list == dict.fromkeys ids."""
    from graph_hypernetwork_forge import _text
    names = [f"relation_{i:05d}" for i in range(50)]
    rng = np.random.default_rng(1)
    texts = [names[i] for i in rng.integers(0, 50, 10000)]
    texts[5] = "relation_00007x"[:-1]                      # equal content, distinct object
    texts[9] = "".join(["relation_", "00007"])
    data, off, edge_map = _text.pack_texts(texts)
    assert edge_map is not None and off.size - 1 < 60
    ids_packed, _ = O.dedup_utf8(np.asarray(data), off)    # what the device kernel computes
    want_unique, want = O.dedup_texts(texts)
    assert np.array_equal(ids_packed[edge_map], want)
    assert len(set(ids_packed.tolist())) == len(want_unique)


def test_analytic_bound_of_generated_weights_holds_and_is_not_too_loose():
    """The generator -> operand-image fusion scales a relation's fp16 image by a power of two chosen from
    |W[r]| <= exp(log_scale) * (|z_r|_1 * max|W3| + max|b3|) without looking at the generated values
    (image_scale_kernel, mp_f16_ss.cu).  Checked here on the numpy oracle's generator: the bound holds for every
    relation and is loose by far less than the 2^14 of exponent headroom fp16 leaves below its 11-bit significand."""
    import numpy as np
    from oracle import hypergnn_oracle as O
    rng = np.random.default_rng(0)
    T, H, d, R = 64, 128, 64, 300
    text = np.tanh(rng.standard_normal((R, T))).astype(np.float32)
    prefix = "g."
    params = {}
    for kind, n_out in (("W_msg", d * d), ("W_self", d * d), ("bias", d)):
        dims = [T, H, H, n_out]
        for li, idx in enumerate((0, 2, 4)):
            fan_in = dims[li]
            lim = 1.0 / np.sqrt(fan_in)
            params[f"{prefix}generators.{kind}.{idx}.weight"] = rng.uniform(-lim, lim, (dims[li + 1], fan_in)).astype(np.float32)
            params[f"{prefix}generators.{kind}.{idx}.bias"] = rng.uniform(-lim, lim, dims[li + 1]).astype(np.float32)
        params[f"{prefix}log_scales.{kind}"] = np.array([rng.uniform(-3, 0.5)], dtype=np.float32)
    w = O.weight_generator(text, params, prefix, d, d, dtype=np.float64)
    worst_loose = 0.0
    for kind in ("W_msg", "W_self"):
        z = text.astype(np.float64)
        for idx in (0, 2):                                   # the hidden layers: Linear + ReLU
            z = np.maximum(z @ params[f"{prefix}generators.{kind}.{idx}.weight"].T.astype(np.float64)
                           + params[f"{prefix}generators.{kind}.{idx}.bias"], 0)
        W3, b3 = params[f"{prefix}generators.{kind}.4.weight"], params[f"{prefix}generators.{kind}.4.bias"]
        alpha = float(np.exp(params[f"{prefix}log_scales.{kind}"][0]))
        bound = alpha * (np.abs(z).sum(axis=1) * np.abs(W3).max() + np.abs(b3).max())       # [R]
        actual = np.abs(w[kind]).reshape(R, -1).max(axis=1)
        assert np.all(actual <= bound * (1 + 1e-6)), kind
        worst_loose = max(worst_loose, float((bound / actual).max()))
    assert worst_loose < 2 ** 8, f"bound loose by {worst_loose:.1f}x"


@pytest.mark.grad
def test_fusion_eligibility_and_autograd_routing_are_host_logic():
    """`_can_fuse_generator` (when the generator may write operand images) and `autograd.wants_grad` need no GPU."""
    import torch
    from graph_hypernetwork_forge import HyperGNN, _native, autograd
    m256 = HyperGNN(64, 16, 256, 2)          # generator width max(64, 2 * 64) = 128
    m64_narrow = HyperGNN(16, 16, 64, 2)     # generator width 64: the tcgen05 Linear needs 128
    m128 = HyperGNN(64, 16, 128, 2)
    t = torch.zeros(1)
    assert m256._can_fuse_generator(_native.PREC_F16, 200, t)
    assert not m256._can_fuse_generator(_native.PREC_F16, 20, t)            # too few relations to fill the kernel
    assert not m256._can_fuse_generator(_native.PREC_FP32, 200, t)
    assert not m64_narrow._can_fuse_generator(_native.PREC_F16, 5000, t)
    assert not m128._can_fuse_generator(_native.PREC_F16, 5000, t)          # hidden 128 keeps weights in TMEM instead
    assert HyperGNN(64, 16, 64, 1)._can_fuse_generator(_native.PREC_F16, 600, t)
    x = torch.zeros(2, requires_grad=True)
    assert autograd.wants_grad(x) and not autograd.wants_grad(x.detach(), None)
    with torch.no_grad():
        assert not autograd.wants_grad(x)
    assert m128._precision_code() == _native.PREC_F16 and HyperGNN(8, 8, 32, 1)._precision_code() == _native.PREC_TF32
    assert HyperGNN(8, 8, 24, 1)._precision_code() == _native.PREC_FP32


def test_list_collapse_in_c_equals_the_numpy_formulation():
    """csrc/pyhost.c (one C pass over the Python list, keyed on object identity) against the numpy formulation it
    replaces: same distinct objects in first-occurrence order, same edge map; tuples and huge vocabularies fall back."""
    from graph_hypernetwork_forge import _text
    if _text._pyhost_lib() is None:
        pytest.skip("lib/libghf_pyhost.so not built")
    names = [f"relation_{r:05d}" for r in range(300)]
    rng = np.random.default_rng(0)
    texts = [names[i] for i in rng.integers(0, 300, 50_000)]
    texts[17] = "relation_00005x"[:-1]                   # equal content, different object
    objs_c, map_c = _text.collapse_by_identity(texts)
    saved, _text._pyhost = _text._pyhost, False
    try:
        objs_np, map_np = _text.collapse_by_identity(texts)
    finally:
        _text._pyhost = saved
    assert all(a is b for a, b in zip(objs_c, objs_np)) and len(objs_c) == len(objs_np) == 301
    assert np.array_equal(map_c, map_np) and map_c.dtype == np.int32
    objs_t, map_t = _text.collapse_by_identity(tuple(texts))      # not a list: the numpy path
    assert np.array_equal(map_t, map_c)
    many = [str(i) for i in range(5000)]                 # every element its own object: the table grows
    objs_m, map_m = _text.collapse_by_identity(many)
    assert objs_m == many and np.array_equal(map_m, np.arange(5000))


@pytest.mark.parametrize("threads", [1, 2, 3, 7, 16])
def test_threaded_list_collapse_equals_the_single_pass(threads):
    """The multi-threaded walk (chunk-local ranks, merged in chunk order, rewritten to global ranks) gives exactly
    the single-threaded result: objects that first appear in a late chunk, chunks with disjoint vocabularies,
    every element its own object, an output buffer supplied by the caller."""
    from graph_hypernetwork_forge import _text
    if _text._pyhost_lib() is None:
        pytest.skip("lib/libghf_pyhost.so not built")
    rng = np.random.default_rng(threads)
    names = [f"rel-{r}" for r in range(97)]
    n = 40_001
    ids = rng.integers(0, 60, n)
    ids[n // 2:] = rng.integers(40, 97, n - n // 2)      # names 60..96 first appear in the second half
    ids[n - 5] = 96
    texts = [names[i] for i in ids]
    objs_1, map_1 = _text.collapse_by_identity(texts, threads=1)
    out = np.full(n, -1, dtype=np.int32)
    objs_t, map_t = _text.collapse_by_identity(texts, out=out, threads=threads)
    assert map_t is out and np.array_equal(map_t, map_1)
    assert len(objs_t) == len(objs_1) and all(a is b for a, b in zip(objs_t, objs_1))
    want = list(dict.fromkeys(texts))                    # the reference's own dedup (HG:264-265)
    assert all(a is b for a, b in zip(objs_t, want)) and len(want) == len(objs_t)
    assert all(texts[i] is objs_t[map_t[i]] for i in range(0, n, 997))
    many = [str(i) for i in range(3000)]                 # every element its own object
    objs_m, map_m = _text.collapse_by_identity(many, threads=threads)
    assert objs_m == many and np.array_equal(map_m, np.arange(3000))
    assert _text.collapse_by_identity([], threads=threads)[0] == []
    with pytest.raises(ValueError):
        _text.collapse_by_identity(texts, out=np.zeros(3, dtype=np.int32))


def test_presync_wrapper_runs_the_hook_exactly_once_on_the_host_side():
    """`_call_with_presync`: when the native entry point never reaches its wait (here: a stand-in call that does not
    touch the device) the hook still runs, once, after the call; nothing stays registered in the library."""
    from graph_hypernetwork_forge import _native
    ran = []
    assert _native._call_with_presync(lambda: "result", lambda: ran.append(1)) == "result"
    assert ran == [1]
    assert _native._call_with_presync(lambda: 7, None) == 7          # no hook: a plain call

    def failing_call():
        raise RuntimeError("native failure")
    with pytest.raises(RuntimeError, match="native failure"):
        _native._call_with_presync(failing_call, lambda: ran.append(2))
    assert ran == [1]                                                # the entry failed: its hook is not run afterwards
    # an edge-less selection returns before any device call: loadable and callable without a GPU
    n = ctypes.c_int64(-1)
    rc = _native.lib().ghf_select_edges(None, 0, 0, 10, None, ctypes.byref(n), None)
    assert rc == 0 and n.value == 0


def test_dropout_stream_restatement_is_pinned_to_philox_known_answers():
    """tools/dropout_stream_probe.py restates torch's dropout counters with a numpy Philox4x32-10; the generator is
    held to the published known-answer vectors of Philox4x32-10 (Random123 kat_vectors) here, on the CPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dropout_probe", os.path.join(ROOT, "tools", "dropout_stream_probe.py"))
    probe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(probe)
    u64 = lambda v: np.array([v], dtype=np.uint64)   # noqa: E731
    kat = [((0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffffffffffff, 0xffffffffffffffff, 0xffffffffffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd))]
    for (seed, ctr, sub), want in kat:
        got = probe.philox(seed, u64(ctr), u64(sub))[:, 0]
        assert tuple(int(x) for x in got) == want
    # the mapping of elements to (thread, step, component) used by the kernel (mp_fuse.cuh: dropout_mult4)
    keep, adv = probe.predicted_keep(4096, 0.25, 1234, 8, sms=148, max_threads_per_sm=2048)
    assert keep.shape == (4096,) and adv == 4 and 0.70 < keep.mean() < 0.80
    from graph_hypernetwork_forge import _native
    assert _native.DropoutState.supported(4096) and not _native.DropoutState.supported(4097)
    assert not _native.DropoutState.supported(0)

"""GPU parity at the sizes BASELINE.json names (`-m gpu`, B200).

The literal oracle cannot run at these sizes, so each case is checked through
  * the SAMPLED-RECEPTIVE-FIELD oracle (`oracle.sampled_forward`): for a few dozen sampled nodes the whole
    forward - every layer's pre-residual update and LayerNorm output, and the final embeddings - is recomputed
    in float64 on the host from the raw edge list, the raw features and the model parameters (every in-edge of
    every node whose value is needed, L hops back).  Nothing of the GPU's own intermediate results enters it.
    Both entries are held to it: `forward_prepared` (with taps, layer by layer) and `forward_packed`
    (`ghf_hypergnn_forward_device`, the call bench.py times);
  * integer work bit-exact at full size: relation ids against a host ranking of first occurrences, in-degrees
    against a host bincount;
  * size-independent properties: the result does not depend on the order of the edge list (the reference sums
    per destination, HG:207-219); repeated runs agree up to the order of atomic additions (bit-identical in the
    deterministic mode).
c2 (FB15k-237 shape) is small enough for the full float64 oracle as well.
"""
import os

import numpy as np
import pytest
import torch

from oracle import hypergnn_oracle as O
from _util import (RERUN_ATOL, TC_H_ATOL_SCALE1, TC_UPD_REL, assert_close, assert_rel_to_max, model_params_numpy)

F16_UPD, F16_H = TC_UPD_REL["f16"], TC_H_ATOL_SCALE1["f16"]

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAME_LEN = 14


def synthetic_on_device(N, E, R, F, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    ei = torch.randint(0, N, (2, E), generator=g, device=DEV, dtype=torch.int64)
    rel = torch.randint(0, R, (E,), generator=g, device=DEV, dtype=torch.int64)
    x = torch.randn(N, F, generator=g, device=DEV, dtype=torch.float32)
    names = np.frombuffer("".join(f"relation_{r:05d}" for r in range(R)).encode(), dtype=np.uint8).reshape(R, NAME_LEN)
    utf8 = torch.from_numpy(names.copy()).to(DEV)[rel].reshape(-1).contiguous()
    offsets = torch.arange(E + 1, device=DEV, dtype=torch.int64) * NAME_LEN
    return x, ei, rel, utf8, offsets


def build(T, F, d, L, precision, log_scale=-1.0, seed=0):
    from graph_hypernetwork_forge import HyperGNN
    torch.manual_seed(seed)
    m = HyperGNN(T, F, d, L, precision=precision).eval()
    with torch.no_grad():                      # O(1) generated weights: the update must matter (SURVEY hard part 4)
        for gen in m.weight_generators:
            for p in gen.log_scales.values():
                p.fill_(log_scale)
    return m.to(DEV)


def host_relation_ranking(rel_np):
    """HG:264-268 on the host: relation value -> id in first-occurrence order, and the distinct texts in that order."""
    values, first = np.unique(rel_np, return_index=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(int(values.max()) + 1, dtype=np.int64)
    rank[values[order]] = np.arange(order.size)
    return rank[rel_np], [f"relation_{int(v):05d}" for v in values[order]]


def assert_same_on_device(got, want, atol, what):
    """max |got - want| over ALL rows, computed on the device (the tensors are gigabytes), recorded like the others."""
    from _util import _record
    err = float((got - want).abs().max())
    _record(what, err, float(want.abs().max()), atol, f"atol={atol:g} (all rows, on device)")
    assert err <= atol, f"{what}: max|diff|={err:.3e} > {atol:g}"


class SampledOracle:
    """float64 recomputation of the forward on the receptive field of `n_samples` nodes (oracle.sampled_forward)."""

    def __init__(self, model, x, ei, rel, n_samples, seed=1):
        self.L, self.d = model.num_layers, model.hidden_dim
        ei_np = ei.cpu().numpy()
        self.rel_ids, self.unique = host_relation_ranking(rel.cpu().numpy())
        self.in_degree = np.bincount(ei_np[1], minlength=x.size(0))
        sample = np.random.default_rng(seed).choice(x.size(0), n_samples, replace=False)
        self.taps = {}
        self.nodes, self.out = O.sampled_forward(model_params_numpy(model), x.cpu().numpy(), ei_np[0], ei_np[1],
                                                 self.rel_ids, self.unique, sample, self.d, self.L, taps=self.taps)

    def check_integer_taps(self, taps):
        assert np.array_equal(taps["edge_rel_ids"].cpu().numpy().astype(np.int64), self.rel_ids), "relation ids"
        assert np.array_equal(taps["in_degree"].cpu().numpy().astype(np.int64), self.in_degree), "in-degree"

    def check_layers(self, taps, upd_rel, h_atol, what):
        for l in range(self.L):
            nodes = torch.from_numpy(self.taps[f"nodes.{l}"]).to(DEV)
            assert_rel_to_max(taps[f"upd.{l}"][nodes].cpu().numpy(), self.taps[f"upd.{l}"], upd_rel,
                              f"upd.{l} {what} ({nodes.numel()} nodes)")
            assert_close(taps[f"h.{l}"][nodes].cpu().numpy(), self.taps[f"h.{l}"], 0.0, h_atol,
                         f"h.{l} {what} ({nodes.numel()} nodes)")

    def check_out(self, out, h_atol, what):
        got = out[torch.from_numpy(self.nodes).to(DEV)].cpu().numpy()
        assert_close(got, self.out, 0.0, h_atol, f"out {what} ({self.nodes.size} nodes)")


@pytest.mark.parametrize("precision", ["f16", "fp32"])
def test_c2_fb15k237_shape_full_oracle(precision):
    """BASELINE config 2 at full size: 14,541 nodes, 272,115 edges, 237 relations, hidden 128, 2 layers: every tap of
    every layer against the full float64 oracle; the one-call entry against the same."""
    from graph_hypernetwork_forge import _text
    N, E, R, d, L, T, F = 14_541, 272_115, 237, 128, 2, 64, 128
    src, dst, rel, names, feats = O.synthetic_kg(N, E, R, F, seed=2)
    texts = [names[r] for r in rel]
    model = build(T, F, d, L, precision)
    params = model_params_numpy(model)
    ref_taps = {}
    ref = O.hypergnn_forward(params, feats, np.stack([src, dst]), texts, d, L, dtype=np.float64, taps=ref_taps)
    taps = {}
    ei = torch.from_numpy(np.stack([src, dst])).to(DEV)
    x = torch.from_numpy(feats).to(DEV)
    out = model.forward_prepared(x, model.prepare(ei, texts, N), taps=taps)
    data, offs = _text.pack_utf8(texts)
    out_packed = model.forward_packed(x, ei, torch.from_numpy(data.copy()).to(DEV), torch.from_numpy(offs).to(DEV))
    assert np.array_equal(taps["edge_rel_ids"].cpu().numpy().astype(np.int64), ref_taps["edge_rel_ids"])
    assert np.array_equal(taps["in_degree"].cpu().numpy().astype(np.int64), ref_taps["in_degree"])
    for l in range(L):
        upd, want = taps[f"upd.{l}"].cpu().numpy(), ref_taps[f"upd.{l}"]
        if precision == "fp32":
            assert_close(upd, want, 1e-4, 2e-5 * float(np.abs(want).max()), f"upd.{l}")
            assert_close(taps[f"h.{l}"].cpu().numpy(), ref_taps[f"h.{l}"], 1e-4, 5e-5, f"h.{l}")
        else:
            assert_rel_to_max(upd, want, F16_UPD, f"upd.{l} f16")
            assert_close(taps[f"h.{l}"].cpu().numpy(), ref_taps[f"h.{l}"], 0.0, F16_H, f"h.{l} f16")
    for name, o in (("out", out), ("out (forward_packed)", out_packed)):
        if precision == "fp32":
            assert_close(o.cpu().numpy(), ref, 1e-4, 5e-5, name)
        else:
            assert_close(o.cpu().numpy(), ref, 0.0, F16_H, name + " f16")


def test_c3_wikikg2_shape_full_size():
    """BASELINE config 3 at full size (2.5M nodes, 16M edges, 535 relations, hidden 128, 3 layers), f16 engine - the
    configuration and the entry point bench.py reports."""
    N, E, R, d, L, T, F = 2_500_000, 16_000_000, 535, 128, 3, 64, 128
    x, ei, rel, utf8, offsets = synthetic_on_device(N, E, R, F)
    model = build(T, F, d, L, "f16")
    oracle = SampledOracle(model, x, ei, rel, 48)
    # (1) the benched entry: one native call
    out = model.forward_packed(x, ei, utf8, offsets)
    assert out.shape == (N, d) and bool(torch.isfinite(out).all())
    oracle.check_out(out, F16_H, "forward_packed")
    # (2) the staged entry, layer by layer
    taps = {}
    out_staged = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, N), taps=taps)
    oracle.check_integer_taps(taps)
    oracle.check_layers(taps, F16_UPD, F16_H, "forward_prepared")
    oracle.check_out(out_staged, F16_H, "forward_prepared")
    del taps
    assert_same_on_device(out_staged, out, RERUN_ATOL, "forward_packed vs forward_prepared f16 (all rows)")
    del out_staged
    # (3) repeated run: only the order of the atomic additions differs
    out2 = model.forward_packed(x, ei, utf8, offsets)
    assert_same_on_device(out2, out, RERUN_ATOL, "repeated run f16 (all rows)")
    del out2
    # (4) the edge list in another order (strings permuted alike) gives the same embeddings
    perm = torch.randperm(E, device=DEV)
    ei_p = ei[:, perm].contiguous()
    utf8_p = utf8.view(E, NAME_LEN)[perm].reshape(-1).contiguous()
    out3 = model.forward_packed(x, ei_p, utf8_p, offsets)
    assert_same_on_device(out3, out, RERUN_ATOL, "permuted edge list f16 (all rows)")


def test_c3_deterministic_mode_is_bit_identical():
    """GHF_DETERMINISTIC=1: per-destination sums are accumulated in fixed point (integer additions commute), so
    repeated runs and permuted edge lists give bit-identical embeddings; still within the f16 tolerance."""
    N, E, R, d, L, T, F = 2_500_000, 16_000_000, 535, 128, 3, 64, 128
    x, ei, rel, utf8, offsets = synthetic_on_device(N, E, R, F)
    model = build(T, F, d, L, "f16")
    oracle = SampledOracle(model, x, ei, rel, 24, seed=5)
    os.environ["GHF_DETERMINISTIC"] = "1"
    try:
        out = model.forward_packed(x, ei, utf8, offsets)
        out2 = model.forward_packed(x, ei, utf8, offsets)
        perm = torch.randperm(E, device=DEV)
        out3 = model.forward_packed(x, ei[:, perm].contiguous(),
                                    utf8.view(E, NAME_LEN)[perm].reshape(-1).contiguous(), offsets)
    finally:
        del os.environ["GHF_DETERMINISTIC"]
    oracle.check_out(out, F16_H, "deterministic forward_packed")
    assert torch.equal(out, out2), "repeated deterministic runs differ"
    assert torch.equal(out, out3), "deterministic run on a permuted edge list differs"


@pytest.mark.parametrize("precision", ["tf32", "f16"])
def test_c5_large_shape_scaled_hidden64(precision):
    """BASELINE config 5's shape (hidden 64, 1k relations, in-degree 10) at 1/100 scale on one GPU: the tf32 engine
    and the f16 engine with streamed weights, both entries, every layer."""
    N, E, R, d, L, T, F = 500_000, 5_000_000, 1000, 64, 2, 64, 64
    x, ei, rel, utf8, offsets = synthetic_on_device(N, E, R, F, seed=3)
    model = build(T, F, d, L, precision)
    oracle = SampledOracle(model, x, ei, rel, 48)
    taps = {}
    out = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, N), taps=taps)
    oracle.check_integer_taps(taps)
    oracle.check_layers(taps, TC_UPD_REL[precision], TC_H_ATOL_SCALE1[precision], f"forward_prepared {precision}")
    oracle.check_out(out, TC_H_ATOL_SCALE1[precision], f"forward_prepared {precision}")
    oracle.check_out(model.forward_packed(x, ei, utf8, offsets), TC_H_ATOL_SCALE1[precision], f"forward_packed {precision}")


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_c4_zero_shot_shape_scaled_hidden256(precision):
    """BASELINE config 4's shape (hidden 256, 1 relation text per 100 edges) at 1/10 scale: the fp32 path and the
    f16 path with streamed weights (mp_f16_ss_kernel)."""
    N, E, R, d, L, T, F = 10_000, 200_000, 2_000, 256, 2, 64, 256
    x, ei, rel, utf8, offsets = synthetic_on_device(N, E, R, F, seed=4)
    model = build(T, F, d, L, precision)
    oracle = SampledOracle(model, x, ei, rel, 16)
    taps = {}
    out = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, N), taps=taps)
    oracle.check_integer_taps(taps)
    if precision == "fp32":
        oracle.check_layers(taps, 1e-4, 1e-4, "forward_prepared fp32")
        oracle.check_out(out, 1e-4, "forward_prepared fp32")
    else:
        oracle.check_layers(taps, F16_UPD, F16_H, "forward_prepared f16")
        oracle.check_out(out, F16_H, "forward_prepared f16")
        oracle.check_out(model.forward_packed(x, ei, utf8, offsets), F16_H, "forward_packed f16 (generator-written images)")


def test_c4_zero_shot_shape_full_size():
    """BASELINE config 4 at FULL size: 100k nodes, 2M edges, 20k distinct relation texts, hidden 256, 2 layers, f16
    engine.  The one-call entry takes the generator -> operand-image path (no fp32 weights exist), the staged entry
    with taps generates fp32 weights (10.5 GB per layer) and packs them; both against the float64 oracle on the
    receptive field of 6 sampled nodes (~2.5k distinct relations, generated chunk by chunk on the host)."""
    N, E, R, d, L, T, F = 100_000, 2_000_000, 20_000, 256, 2, 64, 256
    x, ei, rel, utf8, offsets = synthetic_on_device(N, E, R, F, seed=4)
    model = build(T, F, d, L, "f16")
    oracle = SampledOracle(model, x, ei, rel, 6)
    out = model.forward_packed(x, ei, utf8, offsets)
    assert bool(torch.isfinite(out).all())
    oracle.check_out(out, F16_H, "forward_packed c4 (generator-written images)")
    staged = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, N))      # the same pieces from Python
    oracle.check_out(staged, F16_H, "forward_prepared c4 (generator-written images)")
    del staged
    taps = {}
    out_t = model.forward_prepared(x, model.prepare_packed(ei, utf8, offsets, N), taps=taps)   # fp32 weights + packing
    oracle.check_integer_taps(taps)
    oracle.check_layers(taps, F16_UPD, F16_H, "forward_prepared c4 (packed fp32 weights)")
    oracle.check_out(out_t, F16_H, "forward_prepared c4 (packed fp32 weights)")


@pytest.mark.grad
def test_c3_wikikg2_shape_full_size_gradients_sampled():
    """BASELINE config 3 at full size, one layer of the training step: the gradient kernels against float64
    recomputations from the raw edge list on samples -
      * dL/dW_msg[r], dL/dW_self[r], dL/dbias[r] of three relations (every edge of those relations),
      * dL/dh at 256 sampled nodes (every out-edge and in-edge of those nodes),
    given g_acc = dL/d acc.  Size-independent property: sum_r dL/dbias[r] = sum_v indeg_v * g_acc_v."""
    from graph_hypernetwork_forge import _native
    N, E, R, d = 2_500_000, 16_000_000, 535, 128
    gen = torch.Generator(device=DEV).manual_seed(21)
    ei = torch.randint(0, N, (2, E), generator=gen, device=DEV, dtype=torch.int64)
    rel = torch.randint(0, R, (E,), generator=gen, device=DEV, dtype=torch.int32)
    h = torch.randn(N, d, generator=gen, device=DEV)
    g_acc = torch.randn(N, d, generator=gen, device=DEV) * 1e-3
    W_msg = torch.randn(R, d, d, generator=gen, device=DEV) * 0.05
    W_self = torch.randn(R, d, d, generator=gen, device=DEV) * 0.05
    graph = _native.Graph(ei, rel, N, R, d)
    src, dst = ei[0], ei[1]

    gm, gs, gb = graph.weight_grad(h, g_acc, _native.PREC_F16)
    for r in (0, 77, R - 1):
        idx = (rel == r).nonzero().squeeze(1)
        hs, hd, ga = h[src[idx]].double(), h[dst[idx]].double(), g_acc[dst[idx]].double()
        assert_rel_to_max(gm[r].cpu().numpy(), (hs.T @ ga).cpu().numpy(), 3e-3, f"dL/dW_msg[{r}] ({idx.numel()} edges)")
        assert_rel_to_max(gs[r].cpu().numpy(), (hd.T @ ga).cpu().numpy(), 3e-3, f"dL/dW_self[{r}]")
        assert_rel_to_max(gb[r].cpu().numpy(), ga.sum(0).cpu().numpy(), 3e-3, f"dL/dbias[{r}]")
    indeg = graph.in_degree().double()
    want_total = (indeg.unsqueeze(1) * g_acc.double()).sum(0)
    assert_rel_to_max(gb.double().sum(0).cpu().numpy(), want_total.cpu().numpy(), 3e-3, "sum_r dL/dbias[r]")

    # dL/dh through the two contractions (reversed graph for the messages, the graph itself for the self-loop)
    zero_w, zero_b = torch.zeros_like(W_msg), torch.zeros(R, d, device=DEV)
    g_h = graph.reversed().contract(g_acc, W_msg.transpose(1, 2).contiguous(), zero_w, zero_b, _native.PREC_F16)
    graph.contract(g_acc, zero_w, W_self.transpose(1, 2).contiguous(), zero_b, _native.PREC_F16, out=g_h,
                   accumulate=True)
    # the same through the engine's own transposition and half-skipping (what autograd.py uses at hidden 128)
    g_h2 = graph.reversed().contract(g_acc, W_msg, None, None, _native.PREC_F16, transposed=True)
    graph.contract(g_acc, None, W_self, None, _native.PREC_F16, out=g_h2, accumulate=True, transposed=True)
    assert float((g_h - g_h2).abs().max()) <= 1e-5 * float(g_h.abs().max()), "transposed/NULL-half contraction"
    sample = torch.unique(torch.randint(0, N, (256,), generator=gen, device=DEV))
    want = torch.zeros(sample.numel(), d, dtype=torch.float64, device=DEV)
    out_e = torch.isin(src, sample).nonzero().squeeze(1)          # messages sent by the sampled nodes
    in_e = torch.isin(dst, sample).nonzero().squeeze(1)           # self-loop terms of the sampled nodes
    for edges, at, W in ((out_e, src, W_msg), (in_e, dst, W_self)):
        for lo in range(0, edges.numel(), 2048):
            e = edges[lo:lo + 2048]
            contrib = torch.bmm(g_acc[dst[e]].double().unsqueeze(1), W[rel[e].long()].double().transpose(1, 2)).squeeze(1)
            want.index_add_(0, torch.searchsorted(sample, at[e]), contrib)
    assert_rel_to_max(g_h[sample].cpu().numpy(), want.cpu().numpy(), 3e-3, "dL/dh at sampled nodes")

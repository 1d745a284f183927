"""Gradients of the B200 path (`-m gpu`): `loss.backward()` through `HyperGNN.forward` against
  * the gradients of the UNMODIFIED reference (tests/golden/grad_*.npz, made by tests/golden/make_grad_golden.py), and
  * the float64 autograd of the differentiable oracle (oracle/hypergnn_torch.py, itself pinned to those fixtures)
    on graphs too large for a fixture,
plus the reference's own five training tests (tests/test_hypergnn.py:183-226, tests/test_weight_generator.py:86-106).

Tolerances, each relative to the largest entry of the reference gradient tensor: fp32 path 2e-4 (float32 atomics
and a different summation order than ATen); tensor-core paths (tf32, f16: 11-bit operands in two chained
contractions per layer) 2e-2.
"""
import numpy as np
import pytest
import torch

from _util import GRAD_CASES, build_model, check_grads, load_grad_case, assert_rel_to_max

pytestmark = [pytest.mark.gpu, pytest.mark.grad]
DEV = torch.device("cuda:0")
FP32_GRAD_REL, TC_GRAD_REL = 2e-4, 2e-2


def run_case(gc, precision):
    model = build_model(gc, DEV, precision).train()
    x = torch.tensor(gc["node_features"], device=DEV, requires_grad=True)
    out = model(x, torch.tensor(gc["edge_index"], device=DEV), gc["edge_texts"])
    loss = (out * torch.tensor(gc["loss_weight"], device=DEV)).sum()
    loss.backward()
    grads = {k: p.grad.cpu().numpy() for k, p in model.named_parameters()}
    grads["node_features"] = x.grad.cpu().numpy()
    return out.detach().cpu().numpy(), float(loss.detach()), grads


@pytest.mark.parametrize("name", GRAD_CASES)
def test_fp32_gradients_match_reference(name):
    gc = load_grad_case(name)
    out, loss, grads = run_case(gc, "fp32")
    assert np.abs(out - gc["out"]).max() <= 5e-5
    check_grads(gc, grads, FP32_GRAD_REL, f"{name} fp32")


@pytest.mark.parametrize("name,precision", [("grad_toy", "tf32"), ("grad_synth_d64", "tf32"),
                                            ("grad_synth_d128", "tf32"), ("grad_synth_d128", "f16")])
def test_tensor_core_gradients_match_reference(name, precision):
    gc = load_grad_case(name)
    out, loss, grads = run_case(gc, precision)
    check_grads(gc, grads, TC_GRAD_REL, f"{name} {precision}")


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_GRAD_REL), ("f16", TC_GRAD_REL)])
def test_medium_graph_against_float64_oracle(precision, tol):
    """20k nodes, 200k edges, 37 relations, hidden 128, 2 layers: many units per relation and several edges per
    destination, against float64 autograd of the torch oracle on the same device."""
    from graph_hypernetwork_forge import HyperGNN
    from oracle import hypergnn_torch as OT
    N, E, R, d, L, T, F = 20_000, 200_000, 37, 128, 2, 64, 48
    g = torch.Generator(device=DEV).manual_seed(5)
    ei = torch.randint(0, N, (2, E), generator=g, device=DEV)
    rel = torch.randint(0, R, (E,), generator=g, device=DEV)
    x = torch.randn(N, F, generator=g, device=DEV)
    loss_w = torch.randn(N, d, generator=g, device=DEV)
    names = [f"relation_{r:05d}" for r in range(R)]
    torch.manual_seed(5)
    model = HyperGNN(T, F, d, L, precision=precision)
    with torch.no_grad():
        for gen in model.weight_generators:
            for p in gen.log_scales.values():
                p.fill_(-1.5)
    model = model.to(DEV).train()
    prepared = model.prepare_ids(ei, rel, names, N)
    xg = x.clone().requires_grad_(True)
    out = model.forward_prepared(xg, prepared)
    (out * loss_w).sum().backward()

    params = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    ref = OT.hypergnn_forward(params, x64, ei, rel, names, d, L)
    (ref * loss_w.double()).sum().backward()
    assert_rel_to_max(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), 5e-3 if precision == "f16" else 1e-4, "out")
    assert_rel_to_max(xg.grad.cpu().numpy(), x64.grad.cpu().numpy(), tol, "grad node_features")
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        assert_rel_to_max(p.grad.cpu().numpy(), params[k].grad.cpu().numpy(), tol, f"grad {k}")


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_wide_hidden_dim_trains(dropout):
    """hidden_dim 288 (> 256: the epilogue gradient stages rows in shared memory instead of registers), fp32 engine,
    against float64 autograd of the torch oracle - with and without training-mode dropout (same seed, same stream)."""
    from graph_hypernetwork_forge import HyperGNN
    from oracle import hypergnn_torch as OT
    N, E, R, d, L, T, F = 700, 6000, 5, 288, 2, 16, 20
    g = torch.Generator(device=DEV).manual_seed(288)
    ei = torch.randint(0, N, (2, E), generator=g, device=DEV)
    rel = torch.randint(0, R, (E,), generator=g, device=DEV)
    x = torch.randn(N, F, generator=g, device=DEV)
    loss_w = torch.randn(N, d, generator=g, device=DEV)
    names = [f"wide_{r}" for r in range(R)]
    torch.manual_seed(288)
    model = HyperGNN(T, F, d, L, dropout=dropout, precision="fp32")
    with torch.no_grad():
        for gen in model.weight_generators:
            for q in gen.log_scales.values():
                q.fill_(-2.0)
    model = model.to(DEV).train()
    prepared = model.prepare_ids(ei, rel, names, N)
    xg = x.clone().requires_grad_(True)
    torch.manual_seed(4321)
    out = model.forward_prepared(xg, prepared)
    (out * loss_w).sum().backward()
    if dropout:
        # dropout: float32 oracle on the device (F.dropout has no float64 stream of its own to compare masks with)
        params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        xr, lw, tol = x.clone().requires_grad_(True), loss_w, 1e-3
    else:
        params = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
        xr, lw, tol = x.double().requires_grad_(True), loss_w.double(), FP32_GRAD_REL
    torch.manual_seed(4321)
    ref = OT.hypergnn_forward(params, xr, ei, rel, names, d, L, dropout=dropout)
    (ref * lw).sum().backward()
    assert_rel_to_max(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), 1e-4, "out, hidden 288")
    assert_rel_to_max(xg.grad.cpu().numpy(), xr.grad.cpu().numpy(), tol, "grad node_features, hidden 288")
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        assert_rel_to_max(p.grad.cpu().numpy(), params[k].grad.cpu().numpy(), tol, f"grad {k}, hidden 288")


# ---- the reference's training tests, on CUDA tensors (tests/test_hypergnn.py:183-226) ----
@pytest.fixture
def toy_kg():
    from graph_hypernetwork_forge import ToyKnowledgeGraph
    kg = ToyKnowledgeGraph(feat_dim=16)
    kg.node_features, kg.edge_index = kg.node_features.to(DEV), kg.edge_index.to(DEV)
    return kg


@pytest.fixture
def small_model():
    from graph_hypernetwork_forge import HyperGNN
    return HyperGNN(text_dim=32, node_feat_dim=16, hidden_dim=16, num_layers=2, dropout=0.0).to(DEV)


class TestTraining:
    def test_backward_no_error(self, small_model, toy_kg):
        out = small_model(toy_kg.node_features, toy_kg.edge_index, toy_kg.edge_texts)
        out.sum().backward()

    def test_parameters_update(self, small_model, toy_kg):
        opt = torch.optim.SGD(small_model.parameters(), lr=0.1)
        before = {n: p.clone().detach() for n, p in small_model.named_parameters()}
        opt.zero_grad()
        out = small_model(toy_kg.node_features, toy_kg.edge_index, toy_kg.edge_texts)
        out.sum().backward()
        opt.step()
        changed = sum(not torch.allclose(before[n], p.detach()) for n, p in small_model.named_parameters())
        assert changed > 0, "No parameters changed after an optimiser step"

    def test_loss_decreases(self, toy_kg):
        from graph_hypernetwork_forge import HyperGNN
        model = HyperGNN(text_dim=32, node_feat_dim=16, hidden_dim=16, num_layers=2).to(DEV)
        opt = torch.optim.Adam(model.parameters(), lr=1e-2)
        losses = []
        src, dst = toy_kg.edge_index
        for _ in range(15):
            opt.zero_grad()
            embs = model(toy_kg.node_features, toy_kg.edge_index, toy_kg.edge_texts)
            pos = model.score_triple(embs[src], embs[dst])
            perm = torch.randperm(dst.size(0), device=DEV)
            neg = model.score_triple(embs[src], embs[dst[perm]])
            loss = torch.clamp(1.0 - pos + neg, min=0.0).mean()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert losses[-1] <= losses[0] * 2, "Loss does not appear to decrease at all"

    def test_every_parameter_gets_a_gradient(self, small_model, toy_kg):
        out = small_model(toy_kg.node_features, toy_kg.edge_index, toy_kg.edge_texts)
        (out * torch.randn_like(out)).sum().backward()
        for n, p in small_model.named_parameters():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
            assert float(p.grad.abs().max()) > 0, f"{n}: gradient is identically zero"

    def test_no_grad_forward_is_unchanged(self, small_model, toy_kg):
        assert torch.is_grad_enabled()
        out = small_model(toy_kg.node_features, toy_kg.edge_index, toy_kg.edge_texts)
        with torch.no_grad():
            out2 = small_model(toy_kg.node_features, toy_kg.edge_index, toy_kg.edge_texts)
        assert out.requires_grad and not out2.requires_grad
        assert torch.allclose(out, out2, atol=1e-5)


class TestWeightGeneratorGradients:
    """tests/test_weight_generator.py:86-106."""

    @pytest.fixture
    def weight_gen(self):
        from graph_hypernetwork_forge import WeightGenerator
        return WeightGenerator(text_dim=32, d_in=16, d_out=16, hidden_dim=64).to(DEV)

    def test_gradients_flow(self, weight_gen):
        emb = torch.randn(32, device=DEV, requires_grad=True)
        out = weight_gen(emb)
        (out["W_msg"].sum() + out["W_self"].sum() + out["bias"].sum()).backward()
        assert emb.grad is not None and emb.grad.shape == emb.shape

    def test_scales_appear_in_optimizer(self, weight_gen):
        opt = torch.optim.Adam(weight_gen.parameters(), lr=1e-3)
        out = weight_gen(torch.randn(32, device=DEV))
        sum(v.sum() for v in out.values()).backward()
        opt.step()

    def test_generator_gradients_match_torch(self, weight_gen):
        """LinearFn against torch's own autograd on the same Sequential stacks."""
        x = torch.randn(7, 32, device=DEV)
        out = weight_gen(x)
        wts = [torch.randn_like(v) for v in out.values()]
        sum((v * w).sum() for v, w in zip(out.values(), wts)).backward()
        got = {n: p.grad.clone() for n, p in weight_gen.named_parameters()}
        weight_gen.zero_grad()
        ref = [weight_gen.generators[k](x).view(out[k].shape) * weight_gen.log_scales[k].exp() for k in out]
        sum((v * w).sum() for v, w in zip(ref, wts)).backward()
        for n, p in weight_gen.named_parameters():
            assert_rel_to_max(got[n].cpu().numpy(), p.grad.cpu().numpy(), 1e-4, f"grad {n}")


@pytest.mark.parametrize("M,K,N,relu,scaled", [
    (1000, 64, 128, True, False),       # generator hidden Linear
    (535, 128, 16384, False, True),     # generator head: long contraction for dL/dx (split + atomics), log_scale
    (37, 24, 50, True, True),           # nothing divisible by 4: scalar loaders
    (3, 7, 5, False, False),
    (200_000, 128, 128, True, False),   # input projection: the sum over rows split across CTAs
    (4096, 48, 32, True, False),
])
def test_native_linear_backward_matches_float64(M, K, N, relu, scaled):
    """ghf_linear_backward (dL/dx, dL/dW, dL/db, dL/dlog_scale of y = exp(s) act(x W^T + b)) against float64
    autograd of the same expression; 2e-5 of each gradient's largest entry (fp32 sums in another order)."""
    from graph_hypernetwork_forge import _native
    g = torch.Generator(device=DEV).manual_seed(M + 3 * K + 7 * N)
    x = torch.randn(M, K, generator=g, device=DEV)
    w = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    b = torch.randn(N, generator=g, device=DEV)
    ls = torch.tensor([-0.7], device=DEV) if scaled else None
    g_y = torch.randn(M, N, generator=g, device=DEV)
    y = _native.linear(x, w, b, relu=relu, log_scale=ls)
    g_x, g_w, g_b, g_ls = _native.linear_backward(x, w, ls, y, g_y, relu, need_ls=scaled)

    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, b))
    ls64 = ls.double().requires_grad_(True) if scaled else None
    z = x64 @ w64.t() + b64
    if relu:
        # the mask must be the kernel's own (y > 0): a pre-activation within rounding of zero may fall either side
        z = z * (y > 0).double()
    y64 = z * ls64.exp() if scaled else z
    (y64 * g_y.double()).sum().backward()
    assert_rel_to_max(g_x.cpu().numpy(), x64.grad.cpu().numpy(), 2e-5, "g_x")
    assert_rel_to_max(g_w.cpu().numpy(), w64.grad.cpu().numpy(), 2e-5, "g_w")
    assert_rel_to_max(g_b.cpu().numpy(), b64.grad.cpu().numpy(), 2e-5, "g_b")
    if scaled:
        # the kernel sums g_y * y with the fp32 y of the forward
        want = (g_y.double() * y.double()).sum().reshape(1)
        assert_rel_to_max(g_ls.cpu().numpy(), want.cpu().numpy(), 2e-5, "g_log_scale")
        assert abs(float(ls64.grad) - float(g_ls)) <= 1e-4 * max(1.0, abs(float(ls64.grad)))
    # subsets of outputs
    only_w = _native.linear_backward(x, w, ls, y, g_y, relu, need_x=False, need_b=False)
    assert only_w[0] is None and only_w[2] is None
    assert_rel_to_max(only_w[1].cpu().numpy(), w64.grad.cpu().numpy(), 2e-5, "g_w alone")
    only_b = _native.linear_backward(x, w, ls, y, g_y, relu, need_x=False, need_w=False)
    assert_rel_to_max(only_b[2].cpu().numpy(), b64.grad.cpu().numpy(), 2e-5, "g_b alone")


def test_dropout_in_training_mode_matches_torch_stream():
    """Training mode with dropout > 0 (HG:293-294, WG:103-104): the drop-in draws its masks with F.dropout in the
    reference's call order, so with the same CUDA seed it reproduces the oracle's forward and gradients."""
    from graph_hypernetwork_forge import HyperGNN
    from oracle import hypergnn_torch as OT
    N, E, R, d, L, T, F, p = 500, 4000, 9, 32, 2, 16, 12, 0.25
    g = torch.Generator(device=DEV).manual_seed(9)
    ei = torch.randint(0, N, (2, E), generator=g, device=DEV)
    rel = torch.randint(0, R, (E,), generator=g, device=DEV)
    x = torch.randn(N, F, generator=g, device=DEV)
    loss_w = torch.randn(N, d, generator=g, device=DEV)
    names = [f"r{r}" for r in range(R)]
    torch.manual_seed(9)
    model = HyperGNN(T, F, d, L, dropout=p, precision="fp32")
    with torch.no_grad():
        for gen in model.weight_generators:
            for q in gen.log_scales.values():
                q.fill_(-1.0)
    model = model.to(DEV).train()
    prepared = model.prepare_ids(ei, rel, names, N)
    torch.manual_seed(1234)
    out = model.forward_prepared(x, prepared)
    (out * loss_w).sum().backward()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    torch.manual_seed(1234)
    ref = OT.hypergnn_forward(params, x, ei, rel, names, d, L, dropout=p)
    (ref * loss_w).sum().backward()
    assert_rel_to_max(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), 1e-4, "out with dropout")
    for k, q in model.named_parameters():
        assert_rel_to_max(q.grad.cpu().numpy(), params[k].grad.cpu().numpy(), 5e-4, f"grad {k} with dropout")
    model.eval()                                     # eval mode: no dropout, the fused path
    with torch.no_grad():
        a, b = model.forward_prepared(x, prepared), model.forward_prepared(x, prepared)
    assert torch.equal(a, b) or float((a - b).abs().max()) < 1e-5


@pytest.mark.parametrize("d,precision,N", [(32, "fp32", 500), (64, "tf32", 700), (128, "f16", 3000), (24, "fp32", 333),
                                           (128, "fp32", 2500)])
def test_native_dropout_follows_torch_generator(d, precision, N, monkeypatch):
    """Dropout inside the native row epilogue (ghf_mp_layer_dropout) against the split path that calls F.dropout
    itself: same seed -> the same elements dropped (outputs and every gradient agree to rounding), and the CUDA
    generator ends at the same offset, so whatever draws next sees the reference's stream.  Covers the three vector
    kernels (hidden 32 / 64 / 128), the generic one (hidden 24) and more elements than one step of torch's launch."""
    from graph_hypernetwork_forge import HyperGNN
    E, R, L, T, F, p = 8 * N, 7, 2, 16, 12, 0.3
    g = torch.Generator(device=DEV).manual_seed(d + N)
    ei = torch.randint(0, N, (2, E), generator=g, device=DEV)
    rel = torch.randint(0, R, (E,), generator=g, device=DEV)
    x = torch.randn(N, F, generator=g, device=DEV)
    loss_w = torch.randn(N, d, generator=g, device=DEV)
    names = [f"r{r}" for r in range(R)]
    torch.manual_seed(3)
    model = HyperGNN(T, F, d, L, dropout=p, precision=precision)
    with torch.no_grad():
        for gen in model.weight_generators:
            for q in gen.log_scales.values():
                q.fill_(-1.0)
    model = model.to(DEV).train()
    prepared = model.prepare_ids(ei, rel, names, N)
    gen = torch.cuda.default_generators[0]

    def run(torch_dropout):
        if torch_dropout:
            monkeypatch.setenv("GHF_TORCH_DROPOUT", "1")
        else:
            monkeypatch.delenv("GHF_TORCH_DROPOUT", raising=False)
        model.zero_grad(set_to_none=True)
        torch.manual_seed(77)
        torch.rand(8, device=DEV)                       # the stream does not start at offset 0
        out = model.forward_prepared(x, prepared)
        (out * loss_w).sum().backward()
        return out.detach().clone(), {k: q.grad.clone() for k, q in model.named_parameters()}, gen.get_offset()

    out_n, grads_n, off_n = run(False)
    out_t, grads_t, off_t = run(True)
    assert off_n == off_t, f"generator offsets differ: native {off_n}, torch {off_t}"
    tol = 2e-5 if precision == "fp32" else 2e-3     # tensor-core paths: the two variants chain different fp16 shadows
    assert_rel_to_max(out_n.cpu().numpy(), out_t.cpu().numpy(), tol, "out, native dropout vs F.dropout")
    for k in grads_t:
        assert_rel_to_max(grads_n[k].cpu().numpy(), grads_t[k].cpu().numpy(), 5e-4 if precision == "fp32" else 2e-2,
                          f"grad {k}, native dropout vs F.dropout")
    # one flipped mask element would move a whole LayerNorm row by O(1); also compare against no dropout at all
    model.eval()
    with torch.no_grad():
        out_e = model.forward_prepared(x, prepared)
    assert float((out_e - out_n).abs().max()) > 0.1


def test_score_edges_matches_score_triple_and_its_gradient():
    """`score_edges(embs, heads, tails)` = `score_triple(embs[heads], embs[tails])` (HG:304-318), forward and
    gradient, with repeated ids; out-of-range ids raise."""
    from graph_hypernetwork_forge import HyperGNN
    model = HyperGNN(16, 8, 32, 1).to(DEV)
    for d in (32, 50, 128):
        g = torch.Generator(device=DEV).manual_seed(d)
        embs = torch.randn(300, d, generator=g, device=DEV, requires_grad=True)
        heads = torch.randint(0, 300, (2000,), generator=g, device=DEV)
        tails = torch.randint(0, 300, (2000,), generator=g, device=DEV)
        w = torch.randn(2000, generator=g, device=DEV)
        got = model.score_edges(embs, heads, tails)
        (got * w).sum().backward()
        g_got = embs.grad.clone()
        embs.grad = None
        want = model.score_triple(embs[heads], embs[tails])
        (want * w).sum().backward()
        assert_rel_to_max(got.detach().cpu().numpy(), want.detach().cpu().numpy(), 1e-5, f"scores d={d}")
        assert_rel_to_max(g_got.cpu().numpy(), embs.grad.cpu().numpy(), 1e-5, f"score gradient d={d}")
    with pytest.raises(RuntimeError):
        model.score_edges(embs.detach(), torch.tensor([0, 300], device=DEV), torch.tensor([1, 2], device=DEV))


def test_reference_shaped_message_passing_is_differentiable():
    """`_message_passing(h, edge_index, per-edge weights)` (HG:160-230) against the literal torch formula, forward
    and gradients w.r.t. h and the per-edge tensors."""
    from graph_hypernetwork_forge import HyperGNN
    N, E, d = 40, 300, 24
    g = torch.Generator(device=DEV).manual_seed(3)
    model = HyperGNN(16, 8, d, 1).to(DEV)
    ei = torch.randint(0, N, (2, E), generator=g, device=DEV)
    leaves = [torch.randn(N, d, generator=g, device=DEV), torch.randn(E, d, d, generator=g, device=DEV) * 0.1,
              torch.randn(E, d, d, generator=g, device=DEV) * 0.1, torch.randn(E, d, generator=g, device=DEV)]
    w = torch.randn(N, d, generator=g, device=DEV)
    grads = []
    for native in (True, False):
        h, Wm, Ws, b = [t.clone().requires_grad_(True) for t in leaves]
        if native:
            out = model._message_passing(h, ei, {"W_msg": Wm, "W_self": Ws, "bias": b})
        else:                                                      # HG:201-228, literally
            src, dst = ei
            msg = torch.bmm(h[src].unsqueeze(1), Wm).squeeze(1) + b
            cnt = torch.zeros(N, 1, device=DEV).index_add_(0, dst, torch.ones(E, 1, device=DEV)).clamp(min=1)
            agg = torch.zeros(N, d, device=DEV).index_add(0, dst, msg) / cnt
            S = torch.zeros(N, d, d, device=DEV).index_add(0, dst, Ws) / cnt.unsqueeze(-1)
            out = agg + torch.bmm(h.unsqueeze(1), S).squeeze(1)
        (out * w).sum().backward()
        grads.append([out.detach()] + [t.grad for t in (h, Wm, Ws, b)])
    for name, a, b_ in zip(("out", "g_h", "g_W_msg", "g_W_self", "g_bias"), *grads):
        assert_rel_to_max(a.cpu().numpy(), b_.cpu().numpy(), 2e-4, name)


def test_backward_without_edges_and_with_frozen_parameters():
    """No edges at all: the layers reduce to LN(relu(h)) and only the projection / LayerNorm parameters receive a
    gradient; and with every parameter frozen the gradient w.r.t. the node features still flows."""
    from graph_hypernetwork_forge import HyperGNN
    torch.manual_seed(4)
    model = HyperGNN(16, 8, 32, 2, precision="fp32").to(DEV)
    x = torch.randn(6, 8, device=DEV, requires_grad=True)
    out = model(x, torch.zeros(2, 0, dtype=torch.long, device=DEV), [])
    w = torch.randn_like(out)
    (out * w).sum().backward()
    ref_h = torch.relu(x.detach() @ model.input_proj.weight.T + model.input_proj.bias)
    for ln in model.layer_norms:
        ref_h = torch.nn.functional.layer_norm(torch.relu(ref_h), (32,), ln.weight, ln.bias, ln.eps)
    assert_rel_to_max(out.detach().cpu().numpy(), ref_h.detach().cpu().numpy(), 1e-5, "no-edge forward")
    assert model.input_proj.weight.grad is not None and float(model.input_proj.weight.grad.abs().max()) > 0
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())

    kg_ei = torch.tensor([[0, 1, 2, 3, 4], [1, 2, 3, 4, 5]], device=DEV)
    model.requires_grad_(False)
    model.zero_grad(set_to_none=True)
    x2 = torch.randn(6, 8, device=DEV, requires_grad=True)
    out2 = model(x2, kg_ei, ["a", "b", "a", "c", "b"])
    (out2 * w).sum().backward()
    assert x2.grad is not None and float(x2.grad.abs().max()) > 0
    assert all(p.grad is None for p in model.parameters())


def test_example_training_script_runs():
    """examples/train_link_prediction.py: the reference's demo flow (train with a margin loss, zero-shot relation)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                        "train_link_prediction.py")
    spec = importlib.util.spec_from_file_location("train_link_prediction", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    losses, zero_shot = mod.main(steps=15, verbose=False)
    assert all(l == l for l in losses) and losses[-1] <= losses[0] * 2
    assert zero_shot.shape == (8, 32) and bool(torch.isfinite(zero_shot).all())

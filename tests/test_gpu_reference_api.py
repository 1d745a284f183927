"""The reference's own forward-only tests (tests/test_hypergnn.py, tests/test_weight_generator.py of
danieleschmidt/Graph-Hypernetwork-Forge), restated for CUDA tensors: same names, same assertions.
The five tests that call backward() are outside the forward-only path."""
import pytest
import torch

from graph_hypernetwork_forge import HyperGNN, ToyKnowledgeGraph, WeightGenerator
from graph_hypernetwork_forge.models.hypergnn import TextEncoder

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def simple_kg():
    edge_index = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]], dtype=torch.long, device=DEV)
    return torch.randn(5, 8, device=DEV), edge_index, ["knows", "knows", "works with", "knows"]


@pytest.fixture
def small_model():
    return HyperGNN(text_dim=32, node_feat_dim=16, hidden_dim=16, num_layers=2, dropout=0.0).to(DEV)


@pytest.fixture
def weight_gen():
    return WeightGenerator(text_dim=32, d_in=16, d_out=16, hidden_dim=64).to(DEV)


class TestWeightGeneratorShapes:
    def test_unbatched(self, weight_gen):
        w = weight_gen(torch.randn(32, device=DEV))
        assert w["W_msg"].shape == (16, 16) and w["W_self"].shape == (16, 16) and w["bias"].shape == (16,)

    def test_batched(self, weight_gen):
        w = weight_gen(torch.randn(5, 32, device=DEV))
        assert w["W_msg"].shape == (5, 16, 16) and w["W_self"].shape == (5, 16, 16) and w["bias"].shape == (5, 16)
        assert list(w.keys()) == ["W_msg", "W_self", "bias"]

    def test_batch_of_one_stays_batched(self, weight_gen):
        assert weight_gen(torch.randn(1, 32, device=DEV))["W_msg"].shape == (1, 16, 16)

    def test_non_square(self):
        g = WeightGenerator(text_dim=16, d_in=8, d_out=24).to(DEV)
        w = g(torch.randn(16, device=DEV))
        assert w["W_msg"].shape == (8, 24) and w["bias"].shape == (24,)
        w = WeightGenerator(text_dim=16, d_in=4, d_out=8).to(DEV)(torch.randn(3, 16, device=DEV))
        assert w["W_msg"].shape == (3, 4, 8) and w["bias"].shape == (3, 8)


class TestWeightGeneratorBehaviour:
    def test_deterministic(self, weight_gen):
        e = torch.randn(32, device=DEV)
        assert torch.allclose(weight_gen(e)["W_msg"], weight_gen(e)["W_msg"])

    def test_different_inputs_differ(self, weight_gen):
        a, b = weight_gen(torch.randn(32, device=DEV)), weight_gen(torch.randn(32, device=DEV))
        assert not torch.allclose(a["W_msg"], b["W_msg"])

    def test_num_hidden_zero(self):
        g = WeightGenerator(text_dim=8, d_in=4, d_out=4, num_hidden=0).to(DEV)
        assert g(torch.randn(8, device=DEV))["W_msg"].shape == (4, 4)

    def test_init_scale_keeps_weights_small(self):
        g = WeightGenerator(text_dim=16, d_in=8, d_out=8, init_scale=1e-4).to(DEV)
        assert g(torch.randn(16, device=DEV))["W_msg"].abs().max().item() < 1.0


class TestTextEncoder:
    def test_single_string_shape(self):
        assert TextEncoder(text_dim=32, char_emb_dim=16).to(DEV).encode_one("hello world", DEV).shape == (32,)

    def test_batch_shape(self):
        enc = TextEncoder(text_dim=32, char_emb_dim=16).to(DEV)
        assert enc(["knows", "works at", "is parent of"], DEV).shape == (3, 32)

    def test_empty_string_safe(self):
        out = TextEncoder(text_dim=32).to(DEV).encode_one("", DEV)
        assert out.shape == (32,) and torch.isfinite(out).all()

    def test_different_strings_different_outputs(self):
        enc = TextEncoder(text_dim=32).to(DEV)
        assert not torch.allclose(enc.encode_one("knows", DEV), enc.encode_one("works at", DEV))


class TestHyperGNNForward:
    def test_output_shape(self, small_model, toy_kg):
        out = small_model(toy_kg.node_features.to(DEV), toy_kg.edge_index.to(DEV), toy_kg.edge_texts)
        assert out.shape == (toy_kg.num_nodes, small_model.hidden_dim)

    def test_no_nan_inf(self, small_model, toy_kg):
        out = small_model(toy_kg.node_features.to(DEV), toy_kg.edge_index.to(DEV), toy_kg.edge_texts)
        assert not torch.isnan(out).any() and not torch.isinf(out).any()

    def test_two_nodes_one_edge(self):
        m = HyperGNN(text_dim=16, node_feat_dim=4, hidden_dim=8, num_layers=1).to(DEV)
        out = m(torch.randn(2, 4, device=DEV), torch.tensor([[0], [1]], device=DEV), ["knows"])
        assert out.shape == (2, 8)

    def test_single_layer(self):
        x, ei, et = simple_kg()
        assert HyperGNN(text_dim=16, node_feat_dim=8, hidden_dim=16, num_layers=1).to(DEV)(x, ei, et).shape == (5, 16)

    def test_mismatched_text_count_raises(self, small_model, toy_kg):
        with pytest.raises(ValueError):
            small_model(toy_kg.node_features.to(DEV), toy_kg.edge_index.to(DEV), toy_kg.edge_texts[:-1])


class TestZeroShot:
    def test_unseen_relation(self, small_model, toy_kg):
        ei = torch.cat([toy_kg.edge_index, torch.tensor([[1], [2]])], dim=1).to(DEV)
        out = small_model(toy_kg.node_features.to(DEV), ei, toy_kg.edge_texts + ["is colleague of"])
        assert out.shape == (8, 16) and torch.isfinite(out).all()

    def test_all_unseen(self, small_model):
        x, ei, _ = simple_kg()
        out = small_model.__class__(32, 8, 16, 2).to(DEV)(x, ei, ["α", "β rel", "completely new", "zzz"])
        assert torch.isfinite(out).all()

    def test_single_char_relations(self):
        x, ei, _ = simple_kg()
        out = HyperGNN(16, 8, 16, 2).to(DEV)(x, ei, ["a", "b", "a", "b"])
        assert out.shape == (5, 16) and torch.isfinite(out).all()


class TestScoreTriple:
    def test_shapes_and_self_score(self, small_model):
        a, b = torch.randn(16, device=DEV), torch.randn(16, device=DEV)
        assert small_model.score_triple(a, b).dim() == 0
        assert small_model.score_triple(torch.randn(4, 16, device=DEV), torch.randn(4, 16, device=DEV)).shape == (4,)
        assert small_model.score_triple(a, a).item() > 0

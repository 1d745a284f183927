"""HyperGNN forward on B200 - drop-in for the reference `models/hypergnn.py`.

Same public surface as the reference (`TextEncoder` hypergnn.py:39-81, `HyperGNN`
hypergnn.py:88-322): constructor arguments, attributes, parameter names, the
three `ValueError`s, `score_triple`, `num_parameters`.  What differs is where the
work happens.  `HyperGNN.forward` is

    pack strings (host) -> device dedup -> fused char-bag text encoder
    -> graph build (in-degree, relation-grouped edge order; once per call)
    -> per layer: generator MLPs -> message passing + self-loop + residual + ReLU + LayerNorm

with every stage a hand-written sm_100a kernel behind the C ABI in
`include/ghf_b200.h`.  Tensors must live on a CUDA device; there is no CPU or
eager-PyTorch fallback.  With gradients enabled the same kernels run under
`autograd.py` and `loss.backward()` reaches every parameter.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _native, _text, autograd
from .weight_generator import WeightGenerator


_SIDE_STREAMS = {}   # device -> side stream of forward_prepared (the generators run beside the input projection)


def _to_device(arr: np.ndarray, device: torch.device) -> torch.Tensor:
    arr = np.ascontiguousarray(arr)
    t = torch.from_numpy(arr if arr.flags.writeable else arr.copy())
    if t.numel() > (1 << 16):
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


class PackedTexts:
    """Relation strings on the device: bytes, offsets, per-edge relation ids, distinct-string index."""

    def __init__(self, texts: Optional[List[str]], device: torch.device, utf8: Optional[torch.Tensor] = None,
                 offsets: Optional[torch.Tensor] = None, subset: Optional[torch.Tensor] = None, before_sync=None):
        """`subset` (packed path only): ids of the strings to consider (a rank's own edges); `rel_ids` is then
        indexed like `subset`.  `before_sync`: called once while dedup waits for its host round trip."""
        self.subset = subset
        if texts is None:   # already packed on the device (UTF-8 bytes + int64 offsets[E+1])
            if utf8.dtype != torch.uint8 or offsets.dtype != torch.int64:
                raise RuntimeError("packed texts must be uint8 bytes and int64 offsets")
            self.num_edges = offsets.numel() - 1
            self.utf8, self.offsets, edge_map = utf8.contiguous(), offsets.contiguous(), None
        else:
            # the edge -> distinct-object map is written straight into pinned memory: one copy to the device, no
            # staging copy on the host (64 MB at 16M edges)
            pinned = None
            if len(texts) >= (1 << 16) and torch.device(device).type == "cuda":
                pinned = torch.empty(len(texts), dtype=torch.int32, pin_memory=True)
            data, offs, edge_map = _text.pack_texts(texts, None if pinned is None else pinned.numpy())
            if edge_map is not None and pinned is not None:
                edge_map = pinned                                     # (a tensor: _to_device below takes it as is)
            self.num_edges = len(texts)
            self.utf8 = _to_device(data, device) if data.size else torch.zeros(1, dtype=torch.uint8, device=device)
            self.offsets = _to_device(offs, device)
        ids, first = _native.dedup_texts(self.utf8, self.offsets, subset, before_sync=before_sync)   # packed strings
        self.first = first                                          # packed-string index of each distinct text
        self.num_unique = int(first.numel())
        if edge_map is None:
            self.rel_ids = ids
        else:                                                       # edges -> packed strings -> relation ids
            if isinstance(edge_map, torch.Tensor):
                emap = edge_map.to(device, non_blocking=True)
            else:
                emap = _to_device(edge_map, device)
            self.rel_ids = ids[emap.long()].contiguous()


class RelationIds(PackedTexts):
    """The ids-in form of the relation input: per-edge relation ids plus the distinct relation texts, for callers
    that already hold a relation vocabulary (no per-edge strings, no dedup pass).  `rel_ids[e]` indexes
    `unique_texts`; ids need not be in first-occurrence order - they only name rows of the generated weights."""

    def __init__(self, rel_ids: torch.Tensor, unique_texts: List[str], device: torch.device):
        if rel_ids.dim() != 1 or rel_ids.dtype not in (torch.int32, torch.int64):
            raise RuntimeError("rel_ids must be a 1-D int32/int64 tensor")
        data, offs = _text.pack_utf8(unique_texts)
        self.subset = None
        self.num_edges = rel_ids.numel()
        self.utf8 = _to_device(data, device) if data.size else torch.zeros(4, dtype=torch.uint8, device=device)
        self.offsets = _to_device(offs, device)
        self.num_unique = len(unique_texts)
        self.first = torch.arange(self.num_unique, dtype=torch.int64, device=device)   # text u is packed string u
        self.rel_ids = rel_ids.to(device=device, dtype=torch.int32).contiguous()


class TextEncoder(nn.Module):
    """Character-bag relation encoder: ``tanh(mean(Emb[min(ord(c),127)]) @ W^T + b)``."""

    ASCII_VOCAB = 128

    def __init__(self, text_dim: int, char_emb_dim: int = 32) -> None:
        super().__init__()
        self.text_dim = text_dim
        self.char_emb = nn.Embedding(self.ASCII_VOCAB, char_emb_dim)
        self.proj = nn.Sequential(nn.Linear(char_emb_dim, text_dim), nn.Tanh())

    def _tokenize(self, text: str, device: torch.device) -> torch.Tensor:
        """Token ids as the reference defines them (kept for API parity; the kernel tokenises UTF-8 itself)."""
        codes = [min(ord(ch), self.ASCII_VOCAB - 1) for ch in text] or [0]
        return torch.tensor(codes, dtype=torch.long, device=device)

    def _encode(self, utf8, offsets, index, num) -> torch.Tensor:
        params = (self.char_emb.weight, self.proj[0].weight, self.proj[0].bias)
        if autograd.wants_grad(*params):
            return autograd.TextEncodeFn.apply(*params, utf8, offsets, index, num)
        return _native.text_encode(utf8, offsets, index, num, *params)

    def encode_packed(self, packed: PackedTexts, distinct_only: bool = True) -> torch.Tensor:
        index = packed.first if distinct_only else None
        num = packed.num_unique if distinct_only else packed.offsets.numel() - 1
        return self._encode(packed.utf8, packed.offsets, index, num)

    def forward(self, texts: List[str], device: torch.device) -> torch.Tensor:
        """``[len(texts), text_dim]`` - one row per string, in list order (no dedup)."""
        device = torch.device(device)
        own = _native.require_cuda(self.char_emb.weight)
        if device.type != "cuda" or (device.index is not None and device.index != own.index):
            raise RuntimeError(f"TextEncoder lives on {own}, asked for {device} (no CPU fallback)")
        device = own
        data, offsets = _text.pack_utf8(texts)
        utf8 = _to_device(data, device) if data.size else torch.zeros(1, dtype=torch.uint8, device=device)
        return self._encode(utf8, _to_device(offsets, device), None, len(texts))

    def encode_one(self, text: str, device: torch.device) -> torch.Tensor:
        return self.forward([text], device)[0]


class PreparedGraph:
    """Everything `HyperGNN.forward` derives from (edge_index, edge_texts) alone; reusable across calls."""

    def __init__(self, packed: PackedTexts, graph: "_native.Graph"):
        self.packed = packed
        self.graph = graph


class HyperGNN(nn.Module):
    """Hypernetwork-conditioned relational GNN (forward pass), B200-native.

    ``HyperGNN(text_dim, node_feat_dim, hidden_dim, num_layers=2, dropout=0.0, char_emb_dim=32)``
    as in the reference.  Extension (keyword-only): ``precision`` - ``"f16"`` gathers an fp16 shadow copy
    of the node features and runs the per-edge contraction on tcgen05 kind::f16 with fp32 accumulation
    (hidden_dim 128, and 64 / 256 with streamed weights; same 11-bit operand significand as TF32), ``"tf32"`` runs it on tcgen05 kind::tf32
    (hidden_dim 32/64/128), ``"fp32"`` on CUDA cores; ``None``/"auto" picks f16, then tf32, then fp32 as
    the shape allows (env ``GHF_PRECISION`` fills in when no precision is given).  NOTE: the default is therefore
    NOT the reference's fp32 arithmetic at hidden 32/64/128/256 - tensor-core engines round the operands of the
    per-edge contraction to an 11-bit significand (measured error and tolerances: tests/_util.py); pass
    ``precision="fp32"`` for fp32 end to end.
    """

    def __init__(self, text_dim: int, node_feat_dim: int, hidden_dim: int, num_layers: int = 2,
                 dropout: float = 0.0, char_emb_dim: int = 32, *, precision: Optional[str] = None) -> None:
        super().__init__()
        if num_layers < 1:
            raise ValueError("num_layers must be at least 1")
        self.text_dim, self.node_feat_dim, self.hidden_dim = text_dim, node_feat_dim, hidden_dim
        self.num_layers, self.dropout = num_layers, dropout
        self.precision = precision

        self.text_encoder = TextEncoder(text_dim=text_dim, char_emb_dim=char_emb_dim)
        self.input_proj = nn.Linear(node_feat_dim, hidden_dim)
        self.weight_generators = nn.ModuleList(
            WeightGenerator(text_dim=text_dim, d_in=hidden_dim, d_out=hidden_dim,
                            hidden_dim=max(64, 2 * text_dim), num_hidden=2, dropout=dropout)
            for _ in range(num_layers))
        self.layer_norms = nn.ModuleList(nn.LayerNorm(hidden_dim) for _ in range(num_layers))

    # ------------------------------------------------------------------
    def _precision_code(self) -> int:
        # an explicit constructor argument wins; the environment variable only fills in for precision=None/"auto"
        name = self.precision if self.precision not in (None, "auto") else (os.environ.get("GHF_PRECISION") or "auto")
        if name == "auto":
            name = "f16" if self.hidden_dim in (64, 128, 256) else "tf32" if self.hidden_dim == 32 else "fp32"
        return _native.precision_code(name)

    def prepare(self, edge_index: torch.Tensor, edge_texts: List[str], num_nodes: int,
                dst_range=None) -> PreparedGraph:
        """Dedup the relation strings and build the graph tables for (edge_index, edge_texts)."""
        if edge_index.size(1) != len(edge_texts):
            raise ValueError(
                f"edge_index has {edge_index.size(1)} edges but edge_texts has {len(edge_texts)} entries")
        device = _native.require_cuda(edge_index, self.input_proj.weight)
        return self._prepare(edge_index, PackedTexts(edge_texts, device), num_nodes, dst_range)

    def prepare_packed(self, edge_index: torch.Tensor, utf8: torch.Tensor, offsets: torch.Tensor, num_nodes: int,
                       dst_range=None) -> PreparedGraph:
        """`prepare` for relation strings that are already packed on the device (UTF-8 + offsets[E+1]);
        the way in for edge lists too large for a Python list of str."""
        if edge_index.size(1) != offsets.numel() - 1:
            raise ValueError(
                f"edge_index has {edge_index.size(1)} edges but edge_texts has {offsets.numel() - 1} entries")
        device = _native.require_cuda(edge_index, utf8, offsets, self.input_proj.weight)
        subset = None
        if dst_range is not None and tuple(dst_range) != (0, num_nodes):
            # a rank's share: select its edges once; dedup and graph build then touch only those
            subset = _native.select_edges(edge_index, dst_range[0], dst_range[1])
        return self._prepare(edge_index, PackedTexts(None, device, utf8, offsets, subset), num_nodes, dst_range)

    def prepare_ids(self, edge_index: torch.Tensor, rel_ids: torch.Tensor, unique_texts: List[str], num_nodes: int,
                    dst_range=None) -> PreparedGraph:
        """`prepare` for callers that hold a relation vocabulary: `rel_ids[e]` indexes `unique_texts` (the ids-in
        API of SURVEY 8f; the way in for edge lists of 1e8+ edges, where per-edge strings are impractical)."""
        if edge_index.size(1) != rel_ids.numel():
            raise ValueError(
                f"edge_index has {edge_index.size(1)} edges but rel_ids has {rel_ids.numel()} entries")
        device = _native.require_cuda(edge_index, self.input_proj.weight)
        if rel_ids.numel() and (int(rel_ids.min()) < 0 or int(rel_ids.max()) >= len(unique_texts)):
            raise ValueError("rel_ids must index unique_texts")
        return self._prepare(edge_index, RelationIds(rel_ids, unique_texts, device), num_nodes, dst_range)

    def _prepare(self, edge_index, packed, num_nodes, dst_range) -> PreparedGraph:
        lo, hi = (0, num_nodes) if dst_range is None else dst_range
        graph = _native.Graph(edge_index, packed.rel_ids, num_nodes, max(packed.num_unique, 1),
                              self.hidden_dim, dst_lo=lo, dst_hi=hi,
                              sb_nodes=int(os.environ.get("GHF_SB_NODES", "0")),
                              unit_edges=int(os.environ.get("GHF_UNIT_EDGES", "0")), edge_ids=packed.subset)
        return PreparedGraph(packed, graph)

    def _message_passing(self, h: torch.Tensor, edge_index: torch.Tensor, rel_weights: dict) -> torch.Tensor:
        """Reference-shaped entry (hypergnn.py:160-230): per-EDGE weights ``[E,d,d]``/``[E,d]`` in, the
        pre-residual update ``agg + self_out`` out.  Every edge is its own relation here, so this is the
        slow way in; `forward` passes per-RELATION weights straight to the kernel instead."""
        device = _native.require_cuda(h, edge_index)
        E = edge_index.size(1)
        rel = torch.arange(E, dtype=torch.int32, device=device)
        g = _native.Graph(edge_index, rel, h.size(0), max(E, 1), h.size(1))
        d = h.size(1)
        ones, zeros = torch.ones(d, device=device), torch.zeros(d, device=device)
        W_msg, W_self, bias = rel_weights["W_msg"], rel_weights["W_self"], rel_weights["bias"]
        if E == 0:
            W_msg = W_self = torch.zeros(1, d, d, device=device)
            bias = torch.zeros(1, d, device=device)
        if autograd.wants_grad(h, W_msg, W_self, bias):     # the reference's entry is differentiable; so is this one
            return autograd.mp_update(g, h, W_msg, W_self, bias, _native.PREC_FP32)
        _, upd = g.mp_layer(h, W_msg, W_self, bias, ones, zeros, 1e-5, _native.PREC_FP32, want_upd=True)
        return upd

    def forward_prepared(self, node_features: torch.Tensor, prepared: PreparedGraph,
                         taps: Optional[dict] = None) -> torch.Tensor:
        _native.require_cuda(node_features, self.input_proj.weight)
        graph, packed = prepared.graph, prepared.packed
        if node_features.size(0) != graph.num_nodes:
            raise RuntimeError(f"graph was prepared for {graph.num_nodes} nodes, got {node_features.size(0)}")
        if graph.dst_lo != 0 or graph.dst_hi != graph.num_nodes:
            raise RuntimeError("forward_prepared needs a full-range graph; see distributed.ShardedHyperGNN")
        prec = self._precision_code()
        dropping = self.training and self.dropout > 0.0
        if dropping or self._wants_grad(node_features):
            return self._forward_autograd(node_features, prepared, prec, taps, dropping)
        with torch.no_grad():
            h16 = None   # fp16 shadow of h, chained from layer to layer on the f16 path
            chain = prec == _native.PREC_F16 and self.hidden_dim in (64, 128)   # shadows chained layer to layer
            # The text encoder and the generators of every layer depend on the relation texts only: on a side stream
            # they run beside the input projection instead of between it and the first layer (as in the one-call
            # native forward).  Forked from and joined back into the current stream, so CUDA-graph capture sees one graph.
            dev = node_features.device
            main = torch.cuda.current_stream(dev)
            side = None
            if taps is None and not os.environ.get("GHF_NO_SIDE_GENERATORS"):
                side = _SIDE_STREAMS.get(dev)
                if side is None:
                    side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
                side.wait_stream(main)
            with torch.cuda.stream(side if side is not None else main):
                text_embs = self.text_encoder.encode_packed(packed)
                fuse = taps is None and self._can_fuse_generator(prec, packed.num_unique, text_embs)
                all_w = None if fuse else self._generate_all(text_embs, packed.num_unique)
            if chain:
                h, h16 = _native.linear(node_features, self.input_proj.weight, self.input_proj.bias, relu=True,
                                        want_f16=True)
            else:
                h = _native.linear(node_features, self.input_proj.weight, self.input_proj.bias, relu=True)
            if side is not None:
                main.wait_stream(side)
                text_embs.record_stream(main)                # allocated on the side stream, consumed on this one
                for w in all_w or ():
                    for t in w.values():
                        t.record_stream(main)
            if taps is not None:
                taps["edge_rel_ids"], taps["text_embs"], taps["h0"] = packed.rel_ids, text_embs, h
                taps["in_degree"] = graph.export()["indeg"]
            for l in range(self.num_layers):
                ln = self.layer_norms[l]
                out16 = None
                if chain and l + 1 < self.num_layers:
                    out16 = _native.Shadow(torch.empty((graph.num_local, self.hidden_dim), dtype=torch.float16,
                                                       device=h.device))
                if fuse:    # hidden 64 / 256: the generator writes the contraction's fp16 operand images itself
                    gen = self.weight_generators[l]
                    (zm, lm), (zs, lsf) = gen.hidden("W_msg", text_embs), gen.hidden("W_self", text_embs)
                    images = _native.weight_images(zm, zs, lm.weight, lm.bias, gen.log_scales["W_msg"], lsf.weight,
                                                   lsf.bias, gen.log_scales["W_self"], self.hidden_dim)
                    bias = gen._run_mlp("bias", text_embs).view(packed.num_unique, self.hidden_dim)
                    h = graph.mp_layer_images(h, images, bias, ln.weight, ln.bias, ln.eps, h16=h16, out16=out16)
                    h16 = out16
                    continue
                w = all_w[l]
                h, upd = graph.mp_layer(h, w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, ln.eps, prec,
                                        want_upd=taps is not None, h16=h16, out16=out16)
                h16 = out16
                if taps is not None:
                    taps[f"W_msg.{l}"], taps[f"W_self.{l}"], taps[f"bias.{l}"] = w["W_msg"], w["W_self"], w["bias"]
                    taps[f"upd.{l}"], taps[f"h.{l}"] = upd, h
        return h

    def _can_fuse_generator(self, prec: int, num_unique: int, text_embs: torch.Tensor) -> bool:
        """Generator -> operand-image fusion (`ghf_weight_images_f16`): f16 engine at hidden 64 / 256, generator MLPs
        with hidden layers of width 128, no dropout in flight, enough relations to fill the tcgen05 Linear."""
        d = self.hidden_dim
        if prec != _native.PREC_F16 or d not in (64, 256) or os.environ.get("GHF_NO_FUSED_GENERATOR"):
            return False
        if self.training and self.dropout > 0.0:
            return False
        lin = [m for m in self.weight_generators[0].generators["W_msg"] if isinstance(m, nn.Linear)]
        return len(lin) >= 2 and lin[-1].in_features == 128 and num_unique >= 64 and num_unique * d * d >= (1 << 21)

    def _forward_autograd(self, node_features, prepared: PreparedGraph, prec: int, taps, dropping=False) -> torch.Tensor:
        """The same forward with the autograd graph recorded (`autograd.py`): every stage is the forward kernel
        wrapped in a Function whose backward is the gradient kernel, so `loss.backward()` and an optimiser step
        work as they do on the reference (tests/test_hypergnn.py:183-226, demo.py:79-101).

        `dropping` (training mode with dropout > 0, HG:293-294): the dropout sits between the ReLU and the
        LayerNorm, i.e. inside the row epilogue, and that is where it happens (`ghf_mp_layer_dropout`): the kernel
        regenerates the mask `F.dropout` would draw from torch's CUDA generator at this point of the reference's call
        order (Philox counters restated in tools/dropout_stream_probe.py), so the drop-in follows the reference's
        random stream without a mask tensor; the generator is advanced as `F.dropout` would.  Only when the tensor
        size is outside torch's vectorised kernel (N d % 4 != 0) is the layer split - contraction and mean native
        (`autograd.MPUpdateFn`), residual + ReLU + `F.dropout` + LayerNorm as torch ops."""
        graph, packed = prepared.graph, prepared.packed
        N, d = graph.num_nodes, self.hidden_dim
        # dropout inside the native row epilogue, on torch's own Philox stream, whenever torch's vectorised dropout
        # kernel would handle the [N, d] tensor; otherwise the layer is split and F.dropout itself runs
        native_drop = dropping and _native.DropoutState.supported(N * d) and not os.environ.get("GHF_TORCH_DROPOUT")
        split = dropping and not native_drop
        chain = prec == _native.PREC_F16 and d in (64, 128) and not split   # fp16 shadows chained layer to layer
        made = [] if chain else None
        fast = prec != _native.PREC_FP32             # tensor-core precision mode: the backward GEMMs may use TF32
        for gen in self.weight_generators:
            gen.grad_tf32 = fast
        h = autograd.linear(node_features, self.input_proj.weight, self.input_proj.bias, relu=True, shadow=made,
                            tf32=fast)
        h16 = made[0] if chain else None
        text_embs = self.text_encoder.encode_packed(packed)
        if taps is not None:
            taps["edge_rel_ids"], taps["text_embs"], taps["h0"] = packed.rel_ids, text_embs, h
        for l in range(self.num_layers):
            w = self._generate(l, text_embs, packed.num_unique)
            ln = self.layer_norms[l]
            out16 = None
            if chain and l + 1 < self.num_layers:
                out16 = _native.Shadow(torch.empty((graph.num_local, d), dtype=torch.float16, device=h.device))
            if split:
                upd = autograd.mp_update(graph, h, w["W_msg"], w["W_self"], w["bias"], prec, h16=h16)
                x = nn.functional.dropout(torch.relu(upd + h), p=self.dropout)
                h = nn.functional.layer_norm(x, (d,), ln.weight, ln.bias, ln.eps)
            else:
                drop = _native.DropoutState.draw(self.dropout, N * d, h.device) if native_drop else None
                h = autograd.mp_layer(graph, h, w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, ln.eps,
                                      prec, h16=h16, out16=out16, dropout=drop)
            h16 = out16
            if taps is not None:
                taps[f"W_msg.{l}"], taps[f"W_self.{l}"], taps[f"bias.{l}"] = w["W_msg"], w["W_self"], w["bias"]
                taps[f"h.{l}"] = h
        return h

    def capture_prepared(self, node_features: torch.Tensor, prepared: PreparedGraph):
        """CUDA-graph the layers of `forward_prepared` for one prepared graph: -> (replay, static_input, static_output).

        Small graphs are launch-bound (BASELINE config 2: ~50 kernels of a few microseconds each); after the graph
        tables are built nothing in the forward needs the host (the fp16 scales are chosen on the device), so the
        whole sequence is captured once and replayed with one launch.  Copy new features into `static_input`
        (same shape), call `replay()`, read `static_output`.  Parameters are read at replay time, in place (do not
        reallocate them, e.g. by `model.to(...)`, between capture and replay).  `replay` keeps `prepared` alive: the
        captured kernels hold raw pointers into its graph tables and workspace."""
        dev = _native.require_cuda(node_features, self.input_proj.weight)
        static_in = node_features.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm up outside capture: lazy kernel configuration
            for _ in range(2):
                self.forward_prepared(static_in, prepared)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = self.forward_prepared(static_in, prepared)

        def replay(_keep=(prepared, graph, static_in, static_out)):
            _keep[1].replay()
        return replay, static_in, static_out

    def _generate(self, layer: int, text_embs: torch.Tensor, num_unique: int) -> dict:
        d = self.hidden_dim
        if num_unique == 0:  # no edges: one all-zero relation keeps shapes valid
            dev = text_embs.device
            return {"W_msg": torch.zeros(1, d, d, device=dev), "W_self": torch.zeros(1, d, d, device=dev),
                    "bias": torch.zeros(1, d, device=dev)}
        return self.weight_generators[layer](text_embs)

    def _wants_grad(self, *inputs) -> bool:
        """autograd.wants_grad over the inputs and every parameter - without walking the module tree (75 parameters,
        ~0.1 ms) when gradients are disabled anyway, which is every inference call."""
        if not torch.is_grad_enabled():
            return False
        return autograd.wants_grad(*inputs, *self.parameters())

    def _generate_all(self, text_embs: torch.Tensor, num_unique: int) -> List[dict]:
        """The generated weights of EVERY layer in one native call (they depend on the text embeddings only):
        `ghf_weight_generators` batches the hidden Linears of all 3 L MLPs into one launch per depth level.
        Inference only; falls back to layer-by-layer generation when autograd or dropout is in play."""
        gens = list(self.weight_generators)
        if num_unique == 0 or (self.training and self.dropout > 0.0) or \
                self._wants_grad(text_embs) or len(gens) > 16:
            return [self._generate(l, text_embs, num_unique) for l in range(self.num_layers)]
        mlps = [[[(m.weight, m.bias) for m in gen.generators[kind] if isinstance(m, nn.Linear)]
                 for kind in ("W_msg", "W_self", "bias")] for gen in gens]
        scales = [[gen.log_scales[k] for k in ("W_msg", "W_self", "bias")] for gen in gens]
        return _native.weight_generators(text_embs, mlps, scales, self.hidden_dim, self.hidden_dim)

    def forward_packed(self, node_features: torch.Tensor, edge_index: torch.Tensor, utf8: torch.Tensor,
                       offsets: torch.Tensor) -> torch.Tensor:
        """The whole forward for relation strings already packed on the device (UTF-8 bytes + int64 offsets[E+1]) in
        ONE native call (`ghf_hypergnn_forward_device`): dedup, text encoder, projection, graph build and layers are
        enqueued from C++, so the host language costs one call instead of ~60.  Same result as
        ``forward_prepared(x, prepare_packed(...))``."""
        if edge_index.size(1) != offsets.numel() - 1:
            raise ValueError(
                f"edge_index has {edge_index.size(1)} edges but edge_texts has {offsets.numel() - 1} entries")
        _native.require_cuda(node_features, edge_index, utf8, offsets, self.input_proj.weight)
        if (self.training and self.dropout > 0.0) or self._wants_grad(node_features):
            # the one-call native forward records no autograd graph and has no dropout: go stage by stage
            return self.forward_prepared(node_features,
                                         self.prepare_packed(edge_index, utf8, offsets, node_features.size(0)))
        gen = self.weight_generators[0].generators["W_msg"]
        linears = [m for m in gen if isinstance(m, nn.Linear)]
        desc = _native.ModelDesc(self.text_dim, self.node_feat_dim, self.hidden_dim, self.num_layers,
                                 self.text_encoder.char_emb.embedding_dim,
                                 linears[0].out_features if len(linears) > 1 else 0, len(linears) - 1,
                                 self._precision_code(), float(self.layer_norms[0].eps))
        with torch.no_grad():
            return _native.forward_device(desc, self.flat_parameters(), node_features, edge_index, utf8, offsets)

    def forward(self, node_features: torch.Tensor, edge_index: torch.Tensor, edge_texts: List[str]) -> torch.Tensor:
        """``[N, node_feat_dim]``, ``[2, E]`` int64, E strings -> ``[N, hidden_dim]`` (hypergnn.py:236-298)."""
        prepared = self.prepare(edge_index, edge_texts, node_features.size(0))
        return self.forward_prepared(node_features, prepared)

    # ------------------------------------------------------------------
    def score_triple(self, head_emb: torch.Tensor, tail_emb: torch.Tensor) -> torch.Tensor:
        """Dot-product link score (hypergnn.py:304-318); plain tensor arithmetic, not a kernel target."""
        return (head_emb * tail_emb).sum(dim=-1)

    def score_edges(self, embs: torch.Tensor, heads: torch.Tensor, tails: torch.Tensor) -> torch.Tensor:
        """``score_triple(embs[heads], embs[tails])`` in one kernel: the two ``[B, hidden]`` gathers the training
        loop of the reference materialises (demo.py:90-94) are read in place.  Differentiable w.r.t. ``embs``."""
        _native.require_cuda(embs, heads, tails)
        if autograd.wants_grad(embs):
            return autograd.ScorePairsFn.apply(embs, heads, tails)
        return _native.score_pairs(embs, heads, tails)

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    # flat parameter list in the order ghf_hypergnn_forward_host expects
    def flat_parameters(self) -> List[torch.Tensor]:
        te = self.text_encoder
        out = [te.char_emb.weight, te.proj[0].weight, te.proj[0].bias, self.input_proj.weight, self.input_proj.bias]
        for gen, ln in zip(self.weight_generators, self.layer_norms):
            for kind in ("W_msg", "W_self", "bias"):
                for m in gen.generators[kind]:
                    if isinstance(m, nn.Linear):
                        out += [m.weight, m.bias]
            out += [gen.log_scales[k] for k in ("W_msg", "W_self", "bias")]
            out += [ln.weight, ln.bias]
        return [p.detach().contiguous() for p in out]

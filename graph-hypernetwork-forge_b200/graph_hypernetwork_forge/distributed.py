"""Multi-GPU forward: 1-D destination-range partition + one all-gather of h per layer.

The reference is single-process; this is the scaling design BASELINE.json's
north_star prescribes.  One process per GPU.  Rank k owns destination rows [lo_k, hi_k): it keeps
every edge whose destination falls in that range (sources are arbitrary, so each rank holds all
of h), runs the message-passing layer for its rows, writes them into its slice of the next h and
all-gathers the slices (NCCL over NVLink; gloo in the CPU tests of the host logic).  Generated
relation weights are replicated: every rank runs the text encoder and the generators itself.

`plan_partition` is pure host logic (unit-tested on CPU with gloo, world_size 2).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def plan_partition(num_nodes: int, world_size: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Equal destination ranges, padded so every rank contributes the same number of rows to the
    all-gather.  -> (rows_per_rank, [(lo, hi)] * world_size); hi - lo may be < rows_per_rank (even 0)
    on trailing ranks."""
    if world_size < 1 or num_nodes < 0:
        raise ValueError("world_size must be >= 1 and num_nodes >= 0")
    rows = -(-num_nodes // world_size) if num_nodes else 0
    ranges = []
    for r in range(world_size):
        lo = min(r * rows, num_nodes)
        ranges.append((lo, min(lo + rows, num_nodes)))
    return rows, ranges


def gather_rows(buf: torch.Tensor, rows: int, rank: int, group, async_op: bool = False):
    """In-place all-gather: rank r has filled buf[r*rows:(r+1)*rows]; afterwards every rank has all of buf.
    async_op: returns the work handle; kernels enqueued before `.wait()` overlap the transfer."""
    return dist.all_gather_into_tensor(buf, buf[rank * rows:(rank + 1) * rows], group=group, async_op=async_op)


class ShardedForward:
    """HyperGNN forward over a destination-partitioned graph (one instance per rank)."""

    def __init__(self, model, num_nodes: int, group=None):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.num_nodes = num_nodes
        self.rows, ranges = plan_partition(num_nodes, self.world)
        self.lo, self.hi = ranges[self.rank]
        self.num_kept = 0
        self._bufs: Optional[List[torch.Tensor]] = None

    def _buffers(self, device, d):
        if self._bufs is None:
            shape = (self.rows * self.world, d)
            self._bufs = [torch.zeros(shape, dtype=torch.float32, device=device) for _ in range(2)]
        return self._bufs

    def _buffers16(self, device, d):
        if getattr(self, "_bufs16", None) is None:
            from . import _native
            shape = (self.rows * self.world, d)
            self._bufs16 = [_native.Shadow(torch.zeros(shape, dtype=torch.float16, device=device)) for _ in range(2)]
        return self._bufs16

    def forward_packed(self, node_features, edge_index, utf8, offsets) -> torch.Tensor:
        # the projection of this rank's rows and the all-gather of their fp16 shadow do not depend on the graph:
        # they are enqueued first, and the rows travel over NVLink while edge selection, dedup and graph build run
        started = self._start_h0(node_features)
        prepared = self.model.prepare_packed(edge_index, utf8, offsets, self.num_nodes, dst_range=(self.lo, self.hi))
        return self.forward_prepared(node_features, prepared, _started=started)

    def forward(self, node_features, edge_index, edge_texts) -> torch.Tensor:
        prepared = self.model.prepare(edge_index, edge_texts, self.num_nodes, dst_range=(self.lo, self.hi))
        return self.forward_prepared(node_features, prepared)

    def forward_prepared(self, node_features, prepared, _started=None) -> torch.Tensor:
        """-> [num_nodes, hidden] on every rank (the last layer's slices are gathered as well)."""
        from . import _native
        m = self.model
        if m.training and m.dropout > 0.0:
            raise NotImplementedError("the multi-GPU path is inference-only (no dropout, no gradients)")
        graph, packed = prepared.graph, prepared.packed
        self.num_kept = graph.num_kept
        N, d = self.num_nodes, m.hidden_dim
        prec = m._precision_code()
        cur, nxt = self._buffers(node_features.device, d)
        if prec == _native.PREC_F16 and d == 128:
            return self._forward_f16(graph, packed, _started or self._start_h0(node_features))
        with torch.no_grad():
            # every rank projects all nodes (h is needed in full as the gather source)
            cur[:N] = _native.linear(node_features, m.input_proj.weight, m.input_proj.bias, relu=True)
            text_embs = m.text_encoder.encode_packed(packed)
            for l in range(m.num_layers):
                w = m._generate(l, text_embs, packed.num_unique)
                ln = m.layer_norms[l]
                out = nxt[self.lo:self.hi] if self.hi > self.lo else None
                if out is not None:
                    graph.mp_layer(cur[:N], w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, ln.eps, prec,
                                   out=out)
                gather_rows(nxt, self.rows, self.rank, self.group)
                cur, nxt = nxt, cur
        self._bufs = [cur, nxt]
        return cur[:N]

    def _start_h0(self, node_features):
        """PREC_F16, layer-0 input: project this rank's rows, agree on one scale, convert, start the all-gather of
        the fp16 rows.  -> (cur, nxt, cur16, nxt16, pending work) or None when the path is not the f16 one."""
        from . import _native
        m = self.model
        N, d, lo, hi = self.num_nodes, m.hidden_dim, self.lo, self.hi
        if m._precision_code() != _native.PREC_F16 or d != 128:      # fp16 shadows are chained at hidden 128 only
            return None
        cur, nxt = self._buffers(node_features.device, d)
        cur16, nxt16 = self._buffers16(node_features.device, d)
        with torch.no_grad():
            cur16.scale.zero_()
            if hi > lo:
                cur[lo:hi] = _native.linear(node_features[lo:hi], m.input_proj.weight, m.input_proj.bias, relu=True)
                _native.absmax(cur[lo:hi], cur16)
            # one scale for the whole shadow: the ranks agree on max |h0| first (4 bytes, stream-ordered)
            dist.all_reduce(cur16.scale[1:2], op=dist.ReduceOp.MAX, group=self.group)
            if hi > lo:
                _native.to_f16(cur[lo:hi], cur16.rows(lo, hi), have_amax=True)
            else:
                _native.to_f16(cur[:0], cur16.rows(0, 0), have_amax=True)   # still writes the scale
            # every all-gather is asynchronous: kernels enqueued before the wait (graph preparation here, the
            # generator of the next layer later) run while the rows travel over NVLink
            pending = gather_rows(cur16.data, self.rows, self.rank, self.group, async_op=True)
        return cur, nxt, cur16, nxt16, pending

    def _forward_f16(self, graph, packed, started) -> torch.Tensor:
        """PREC_F16: a rank needs fp32 h only for its own rows (residual), and the fp16 shadow of h in full (gather
        source).  So every rank projects only its own rows, and what travels between layers is the fp16 copy -
        half the all-gather bytes; the fp32 rows are gathered once, after the last layer, for the return value."""
        from . import _native
        m = self.model
        N, lo, hi = self.num_nodes, self.lo, self.hi
        cur, nxt, cur16, nxt16, pending = started
        with torch.no_grad():
            text_embs = m.text_encoder.encode_packed(packed)
            w = m._generate(0, text_embs, packed.num_unique)
            for l in range(m.num_layers):
                ln = m.layer_norms[l]
                last = l + 1 == m.num_layers
                pending.wait()
                if hi > lo:
                    graph.mp_layer(cur[:N], w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, ln.eps,
                                   _native.PREC_F16, out=nxt[lo:hi], h16=cur16.rows(0, N),
                                   out16=None if last else nxt16.rows(lo, hi))
                pending = gather_rows(nxt if last else nxt16.data, self.rows, self.rank, self.group, async_op=True)
                if not last:
                    w = m._generate(l + 1, text_embs, packed.num_unique)
                cur, nxt, cur16, nxt16 = nxt, cur, nxt16, cur16
            pending.wait()
        self._bufs, self._bufs16 = [cur, nxt], [cur16, nxt16]
        return cur[:N]

"""Multi-GPU forward: 1-D destination-range partition, one exchange of h per layer.

The reference is single-process; this is the scaling design BASELINE.json's north_star prescribes
(SURVEY 8e).  One process per GPU.  Rank k owns destination rows [lo_k, hi_k): it keeps every edge whose destination
falls in that range (sources are arbitrary, so each rank needs all of h as gather source), runs the layers for its
rows and makes the new rows visible to every rank.  Generated relation weights are replicated: every rank runs the
text encoder and the generators itself.

What travels, and how:

  * On the f16 engine (hidden 64 / 128) the kernels gather from the fp16 SHADOW of h and read fp32 h only at a rank's
    own rows (residual).  So a rank keeps fp32 rows for its own range only, and only fp16 rows travel - half the bytes.
  * transport "p2p" (default on CUDA): the shadow lives in torch symmetric memory, [N, d] fp16 on every rank, two
    buffers (read layer l / write layer l + 1).  A rank's epilogue kernel writes the new rows into its own copy;
    a side stream then pushes them into every peer's copy with plain device-to-device copies over NVLink (copy
    engines - the persistent contraction kernel owns every SM, a collective kernel could not run beside it), chunk
    by chunk while the next chunk of rows is still being computed, and a device-side barrier closes the layer.
    No staging buffer, no equal-size constraint: ranges may be balanced by EDGES (`plan_partition_by_edges`).
  * transport "collective" (gloo on CPU, or NCCL when symmetric memory is unavailable): an in-place all-gather of
    padded equal slices, or one broadcast per rank when the ranges are uneven.
  * The result stays sharded - `forward*` return the rank's own rows [hi - lo, d] ("the last layer's
    output can stay sharded"); `gather_output=True` all-gathers the fp32 rows for callers that want the reference's
    full [N, d] return value on every rank.

`plan_partition`, `plan_partition_by_edges` and the collective transport are host logic, unit-tested on CPU with gloo
(world_size 2, tests/test_distributed_cpu.py); the p2p transport needs GPUs (tests/test_gpu_multi.py, bench.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def plan_partition(num_nodes: int, world_size: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Equal destination ranges, padded so every rank contributes the same number of rows to an all-gather.
    -> (rows_per_rank, [(lo, hi)] * world_size); hi - lo may be < rows_per_rank (even 0) on trailing ranks."""
    if world_size < 1 or num_nodes < 0:
        raise ValueError("world_size must be >= 1 and num_nodes >= 0")
    rows = -(-num_nodes // world_size) if num_nodes else 0
    ranges = []
    for r in range(world_size):
        lo = min(r * rows, num_nodes)
        ranges.append((lo, min(lo + rows, num_nodes)))
    return rows, ranges


def plan_partition_by_edges(rowptr, world_size: int, node_weight: float = 1.0) -> List[Tuple[int, int]]:
    """Contiguous destination ranges balanced by WORK instead of by node count (power-law in-degree makes
    equal-node ranges unbalanced).  `rowptr` [N + 1] is the exclusive scan of the in-degrees of all nodes (the
    dst-CSR row pointer).  Work of a range = its edges + node_weight * its nodes (the per-node epilogue is not free);
    boundary k is the first node where the cumulative work reaches k / world_size of the total.  Deterministic:
    every rank computes the same ranges from the same rowptr."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    rp = torch.as_tensor(rowptr).to(torch.float64).cpu()
    n = rp.numel() - 1
    if n < 0:
        raise ValueError("rowptr must have at least one entry")
    work = rp + node_weight * torch.arange(n + 1, dtype=torch.float64)
    targets = work[-1] * torch.arange(1, world_size, dtype=torch.float64) / world_size
    cuts = torch.searchsorted(work, targets, right=False).clamp_(0, n).tolist()
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):                      # monotone even with empty ranges
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_rows(buf: torch.Tensor, rows: int, rank: int, group, async_op: bool = False):
    """In-place all-gather: rank r has filled buf[r*rows:(r+1)*rows]; afterwards every rank has all of buf.
    async_op: returns the work handle; kernels enqueued before `.wait()` overlap the transfer."""
    return dist.all_gather_into_tensor(buf, buf[rank * rows:(rank + 1) * rows], group=group, async_op=async_op)


def exchange_rows(buf: torch.Tensor, ranges: Sequence[Tuple[int, int]], rank: int, group, async_op: bool = False):
    """Every rank has filled buf[lo_r:hi_r] of its own range; afterwards every rank has all rows.  Equal padded
    ranges (plan_partition) go through one in-place all-gather; uneven ranges (plan_partition_by_edges) through one
    broadcast per rank.  -> list of work handles (empty when not async)."""
    world = len(ranges)
    rows, equal_ranges = plan_partition(ranges[-1][1], world)
    if [tuple(r) for r in ranges] == equal_ranges and rows > 0 and buf.shape[0] >= rows * world:
        w = gather_rows(buf[:rows * world], rows, rank, group, async_op)
        return [w] if async_op else []
    works = []
    for r, (lo, hi) in enumerate(ranges):
        if hi > lo:
            src = dist.get_global_rank(group, r) if group is not None else r
            w = dist.broadcast(buf[lo:hi], src=src, group=group, async_op=async_op)
            if async_op:
                works.append(w)
    return works


class _SymmetricRows:
    """[N, d] fp16 table replicated on every rank in torch symmetric memory (two buffers).  A rank writes its own
    rows locally and pushes them into the peers' copies with device-to-device copies on a side stream."""

    def __init__(self, num_rows: int, d: int, group, device):
        import torch.distributed._symmetric_memory as symm
        from . import _native
        self._native = _native
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        name = group.group_name if hasattr(group, "group_name") else dist.group.WORLD.group_name
        try:                                            # older torch needs the group enabled explicitly
            if hasattr(symm, "is_symm_mem_enabled_for_group") and not symm.is_symm_mem_enabled_for_group(name):
                symm.enable_symm_mem_for_group(name)
        except Exception:  # noqa: BLE001
            pass
        self.bufs, self.handles, self.peers = [], [], []
        for _ in range(2):
            t = symm.empty((max(num_rows, 1), d), dtype=torch.float16, device=device)
            hdl = symm.rendezvous(t, name)
            self.bufs.append(t)
            self.handles.append(hdl)
            self.peers.append([t if r == self.rank else hdl.get_buffer(r, tuple(t.shape), torch.float16)
                               for r in range(self.world)])
        self.comm = torch.cuda.Stream(device=device)
        self.bytes_pushed = 0
        # base pointers of every rank's tables, on the device (the epilogue kernel stores through them)
        self.table_ptrs = [torch.tensor([t.data_ptr() for t in self.peers[b]], dtype=torch.int64, device=device)
                           for b in range(2)]

    def push(self, b: int, lo: int, hi: int) -> None:
        """Rows [lo, hi) of buffer b, already written locally by work enqueued on the CURRENT stream, go to every
        peer (enqueued on the side stream, after that work)."""
        if hi <= lo or self.world == 1:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.comm.wait_event(ev)
        src = self.bufs[b][lo:hi]
        with torch.cuda.stream(self.comm):
            for k in range(1, self.world):              # start with the next rank: the pushes of all ranks interleave
                r = (self.rank + k) % self.world
                self._native.copy_async(self.peers[b][r][lo:hi], src)
        self.bytes_pushed += (self.world - 1) * src.numel() * 2

    def close_layer(self, b: int) -> None:
        """All pushes of every rank into buffer b have landed everywhere before anything enqueued on the CURRENT stream
        after this call runs (device-side barrier on the side stream; the host does not wait)."""
        if self.world == 1:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))   # peers must not be signalled before our own compute
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            self.handles[b].barrier(channel=0)
            done = torch.cuda.Event()
            done.record(self.comm)
        torch.cuda.current_stream(self.device).wait_event(done)


class _ShardedLayerFn(torch.autograd.Function):
    """One message-passing layer of a rank (f16 engine, hidden 128) with its gradients.

    Forward: `ghf_mp_layer_f16` on the rank's graph - fp32 rows `h_local` of the own range (residual), the fp16 table
    `h16` of ALL rows (gather source; a non-differentiable carrier: every rank holds the same values).  Backward, given
    dL/d out for the own rows: the epilogue gradient and the self-loop term are local; the weight gradients are this
    rank's share (summed over ranks at the end, `ShardedForward.allreduce_gradients`); the message term
    `g_acc_v W_msg[r]^T` lands on the SOURCE u of every edge (u -> v) with v in the own range - any row of the graph -
    so it is contracted over the rank's reversed edges into a partial [N, d] and summed across ranks, each owner
    keeping its rows (reduce-scatter; the one collective of a layer's backward)."""

    @staticmethod
    def forward(ctx, h_local, W_msg, W_self, bias, ln_w, ln_b, sf, graph, rev_graph, eps, h16, out16):
        from . import _native
        out, upd = graph.mp_layer(h_local, W_msg, W_self, bias, ln_w, ln_b, eps, _native.PREC_F16, want_upd=True,
                                  h16=h16, out16=out16, h_row0=sf.lo)
        ctx.sf, ctx.graph, ctx.rev_graph, ctx.eps, ctx.h16 = sf, graph, rev_graph, eps, h16
        ctx.save_for_backward(h_local, W_msg, W_self, ln_w, upd)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        from . import _native
        h_local, W_msg, W_self, ln_w, upd = ctx.saved_tensors
        sf, graph, rev, prec = ctx.sf, ctx.graph, ctx.rev_graph, _native.PREC_F16
        N, d, lo, hi = sf.num_nodes, graph.hidden_dim, sf.lo, sf.hi
        dev = g_out.device
        # fp16 table of g_acc over ALL rows; only the own rows are ever read (destinations of the own edges, sources
        # of the reversed ones), so the rest stays unwritten
        g16 = _native.Shadow(torch.empty((N, d), dtype=torch.float16, device=dev))
        g_pre, g_acc, g_ln_w, g_ln_b, _ = graph.epilogue_backward(g_out.contiguous(), upd, h_local, ln_w, ctx.eps,
                                                                  h_row0=lo, shadow_out=g16.rows(lo, hi))
        needs = ctx.needs_input_grad
        g_wm = g_ws = g_b = None
        if any(needs[1:4]):
            g_wm, g_ws, g_b = graph.weight_grad(None, g_acc, prec, h16=ctx.h16, g16=g16.rows(lo, hi))
        g_h = None
        if needs[0]:
            # self-loop term: stays at the own rows, added to the residual's share
            graph.contract(None, None, W_self, None, prec, x16=g16, out=g_pre, accumulate=True, transposed=True)
            # message term: partial sums for every row of the graph from the own (reversed) edges ...
            partial = rev.contract(None, W_msg, None, None, prec, x16=g16, transposed=True)
            # ... summed over ranks; each owner keeps its rows
            g_h = g_pre
            if sf.world > 1:
                equal = all(b - a == sf.rows for a, b in sf.ranges) and sf.rows * sf.world == N
                if equal and dev.type == "cuda":
                    mine = torch.empty((hi - lo, d), dtype=partial.dtype, device=dev)
                    dist.reduce_scatter_tensor(mine, partial, group=sf.group)
                    g_h += mine
                else:
                    dist.all_reduce(partial, group=sf.group)
                    g_h += partial[lo:hi]
            else:
                g_h += partial[lo:hi]
        return g_h, g_wm, g_ws, g_b, g_ln_w, g_ln_b, None, None, None, None, None, None


class ShardedForward:
    """HyperGNN forward over a destination-partitioned graph (one instance per rank).

    ranges: optional [(lo, hi)] per rank (e.g. from `plan_partition_by_edges`); default equal node counts.
    transport: "p2p" | "collective" | None (p2p on CUDA when symmetric memory works, else collective).
    chunks: the rank's super-blocks are processed in this many pieces (`ghf_mp_layer_f16_range`) so that the rows of a
    finished piece travel while the next one is computed (p2p transport, f16 engine, push="copy"); 0 = one piece per
    ~8 super-blocks, at most 4.
    push (p2p transport): "kernel" - the layer kernel's row epilogue stores each new fp16 row straight into the tables
    of the peers that READ it (the sources of their edges; one all-to-all of byte masks per graph): about half the
    bytes of an all-gather at in-degree 6, and the NVLink stores of one super-block overlap the contraction of the
    next ones; "copy" - whole row ranges go to every peer with device-to-device copies on a side stream (copy
    engines), overlapping the next chunk's compute; "auto" (default) picks per graph from a two-line cost model
    (`_pick_push`)."""

    # cost model of `_pick_push`: seconds per edge and layer of the contraction at hidden 128, NVLink bytes per second
    _T_EDGE_128, _NVLINK_BPS = 1.9e-10, 7.0e11

    def __init__(self, model, num_nodes: int, group=None, ranges=None, transport: Optional[str] = None,
                 chunks: int = 0, push: str = "auto"):
        self.model = model
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.num_nodes = num_nodes
        self.rows, equal = plan_partition(num_nodes, self.world)
        self.ranges = [tuple(map(int, r)) for r in (ranges if ranges is not None else equal)]
        if len(self.ranges) != self.world or self.ranges[0][0] != 0 or self.ranges[-1][1] != num_nodes or \
                any(a[1] != b[0] for a, b in zip(self.ranges, self.ranges[1:])):
            raise ValueError("ranges must tile [0, num_nodes) in rank order")
        self.lo, self.hi = self.ranges[self.rank]
        self.transport = transport
        self.chunks = max(0, int(chunks))
        if push not in ("kernel", "copy", "auto"):
            raise ValueError("push must be 'kernel', 'copy' or 'auto'")
        self.push = push
        self.push_used = None
        self.rows_needed_by_peers = None
        self.num_kept = 0
        self.profile = False          # True: CUDA events on the main stream between the stages (read with `stage_ms`)
        self._marks: List[Tuple[str, torch.cuda.Event]] = []
        self._bufs: Optional[List[torch.Tensor]] = None
        self._sym: Optional[_SymmetricRows] = None

    def _mark(self, label: str) -> None:
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((label, ev))

    def stage_ms(self) -> dict:
        """Main-stream time per stage of the profiled forwards since the last call (synchronises): `prep` (projection
        of the own rows, edge selection, dedup, graph build, text encoder, first generator), `compute` (layer kernels
        + generators), `wait` (the main stream waiting for rows to arrive = exposed exchange time)."""
        out = {"prep": 0.0, "compute": 0.0, "wait": 0.0}
        if self._marks:
            self._marks[-1][1].synchronize()
        for (_, a), (label, b) in zip(self._marks, self._marks[1:]):
            if label != "start":
                out[label] += a.elapsed_time(b)
        self._marks = []
        return out

    # ------------------------------------------------------------------ helpers
    def _f16_engine(self) -> bool:
        from . import _native
        m = self.model
        return m._precision_code() == _native.PREC_F16 and m.hidden_dim in (64, 128)

    def _symmetric(self, device) -> Optional[_SymmetricRows]:
        if self.transport == "collective" or device.type != "cuda":
            return None
        if self._sym is None:
            try:
                self._sym = _SymmetricRows(self.num_nodes, self.model.hidden_dim, self.group, device)
            except Exception as e:  # noqa: BLE001
                if self.transport == "p2p":
                    raise RuntimeError(f"p2p transport unavailable: {e!r}") from e
                self.transport = "collective"
                return None
            self.transport = "p2p"
        return self._sym

    def _phase_chunks(self, graph) -> List[Tuple[int, int]]:
        """The graph's super-blocks (a Graph, or their number) in `chunks` contiguous pieces (a piece is one launch
        of the layer kernels)."""
        n = graph if isinstance(graph, int) else graph.num_phases
        c = max(1, min(self.chunks, n)) if self.chunks else max(1, min(4, n // 8))
        cuts = [round(i * n / c) for i in range(c + 1)]
        return [(a, b) for a, b in zip(cuts, cuts[1:]) if b > a]

    def _local_rows(self, node_features: torch.Tensor) -> torch.Tensor:
        """Features of this rank's own rows: accepts all [N, F] rows or just the [hi - lo, F] local ones."""
        if node_features.shape[0] == self.num_nodes:
            return node_features[self.lo:self.hi]
        if node_features.shape[0] == self.hi - self.lo:
            return node_features
        raise RuntimeError(f"node_features must have {self.num_nodes} or {self.hi - self.lo} rows")

    def _gather_output(self, local: torch.Tensor) -> torch.Tensor:
        """[hi - lo, d] on each rank -> [N, d] on every rank (fp32 rows over the collective)."""
        d = local.shape[1]
        full = torch.empty((max(self.rows * self.world, self.num_nodes), d), dtype=local.dtype, device=local.device)
        full[self.lo:self.hi] = local
        exchange_rows(full, self.ranges, self.rank, self.group)
        return full[:self.num_nodes]

    # ------------------------------------------------------------------ entries
    def forward_packed(self, node_features, edge_index, utf8, offsets, gather_output: bool = False) -> torch.Tensor:
        """Full edge list + packed relation strings on every rank; the rank selects its own edges."""
        from . import _native
        from .models.hypergnn import PackedTexts
        m = self.model
        self._mark("start")
        device = _native.require_cuda(edge_index, utf8, offsets, m.input_proj.weight)
        if m._wants_grad(node_features):                 # training: the autograd path (no overlap tricks)
            subset = None
            if (self.lo, self.hi) != (0, self.num_nodes):
                subset = _native.select_edges(edge_index, self.lo, self.hi)
            return self._run_autograd(node_features, edge_index, PackedTexts(None, device, utf8, offsets, subset),
                                      gather_output)
        box = {}

        def start_h0():                                  # the rows travel while edge selection and graph build run
            box["started"] = self._start_h0(node_features)
        subset = None
        if (self.lo, self.hi) != (0, self.num_nodes):
            # the selection kernels go first; the projection of the own rows, the scale agreement and the push of
            # the fp16 rows are enqueued while the host would otherwise just wait for the number of selected edges
            subset = _native.select_edges(edge_index, self.lo, self.hi, before_sync=start_h0)
        else:
            start_h0()

        def start_peer_mask():
            # who reads which of this rank's rows needs only the selected edges: marked and exchanged while dedup
            # waits for its count, instead of between graph build and the first layer
            st = box["started"]
            if st is not None and self._sym is not None and self.world > 1:
                kept = int(subset.numel()) if subset is not None else int(edge_index.shape[1])
                sb = max(1024, (48 << 20) // (8 * m.hidden_dim))          # graph build's default super-block
                self._decide_push(st, kept, -(-max(self.hi - self.lo, 1) // sb), edge_index, subset)
        packed = PackedTexts(None, device, utf8, offsets, subset, before_sync=start_peer_mask)
        return self._run(node_features, edge_index, packed, box["started"], gather_output)

    def forward(self, node_features, edge_index, edge_texts, gather_output: bool = False) -> torch.Tensor:
        """The reference's call shape (List[str]); every rank holds the full edge list."""
        from . import _native
        from .models.hypergnn import PackedTexts
        if edge_index.size(1) != len(edge_texts):
            raise ValueError(f"edge_index has {edge_index.size(1)} edges but edge_texts has {len(edge_texts)} entries")
        self._mark("start")
        device = _native.require_cuda(edge_index, self.model.input_proj.weight)
        if self.model._wants_grad(node_features):
            return self._run_autograd(node_features, edge_index, PackedTexts(edge_texts, device), gather_output)
        started = self._start_h0(node_features)
        packed = PackedTexts(edge_texts, device)
        return self._run(node_features, edge_index, packed, started, gather_output)

    def forward_ids(self, node_features, edge_index, rel_ids, unique_texts, gather_output: bool = False) -> torch.Tensor:
        """ids-in entry for PRE-SHARDED edge lists: `edge_index` holds only edges whose destination
        lies in this rank's range (edges outside it are ignored), `rel_ids[e]` indexes `unique_texts`, the same
        vocabulary on every rank.  The way in for graphs whose global edge list fits no single GPU (BASELINE config 5)."""
        from . import _native
        from .models.hypergnn import RelationIds
        if edge_index.size(1) != rel_ids.numel():
            raise ValueError(f"edge_index has {edge_index.size(1)} edges but rel_ids has {rel_ids.numel()} entries")
        self._mark("start")
        device = _native.require_cuda(edge_index, self.model.input_proj.weight)
        if self.model._wants_grad(node_features):
            return self._run_autograd(node_features, edge_index, RelationIds(rel_ids, unique_texts, device),
                                      gather_output)
        started = self._start_h0(node_features)
        packed = RelationIds(rel_ids, unique_texts, device)
        return self._run(node_features, edge_index, packed, started, gather_output)

    def forward_prepared(self, node_features, prepared, _started=None, gather_output: bool = True) -> torch.Tensor:
        """Round-1 entry kept for callers that built the rank's graph themselves (`model.prepare*(dst_range=...)`):
        full [N, d] result on every rank by default."""
        graph, packed = prepared.graph, prepared.packed
        if (graph.dst_lo, graph.dst_hi) != (self.lo, self.hi):
            raise RuntimeError("the prepared graph covers another destination range")
        started = _started if _started is not None else self._start_h0(node_features)
        return self._layers(node_features, graph, packed, started, gather_output)

    # ------------------------------------------------------------------ the forward
    def _run(self, node_features, edge_index, packed, started, gather_output):
        import os
        from . import _native
        m = self.model
        if m.training and m.dropout > 0.0:
            raise NotImplementedError("the multi-GPU path has no dropout")
        hook = None
        if started is not None:
            # text encoder + the generators of every layer need only the dedup result.  They are enqueued (on the side
            # stream, ~0.3 ms of host time) from inside graph build, after its kernels are in the queue and before it
            # waits for the table sizes: they run beside the sort instead of after it, and cost no time on the host.
            dedup_done = torch.cuda.Event()
            dedup_done.record(torch.cuda.current_stream(node_features.device))

            def hook():
                started["weights"] = self._start_generators(packed, node_features.device, after=dedup_done)
        graph = _native.Graph(edge_index, packed.rel_ids, self.num_nodes, max(packed.num_unique, 1), m.hidden_dim,
                              dst_lo=self.lo, dst_hi=self.hi, sb_nodes=int(os.environ.get("GHF_SB_NODES", "0")),
                              unit_edges=int(os.environ.get("GHF_UNIT_EDGES", "0")), edge_ids=packed.subset,
                              before_sync=hook)
        if started is not None and self._sym is not None and self.world > 1 and "push" not in started:
            self._decide_push(started, graph.num_kept, graph.num_phases, edge_index, packed.subset)
        return self._layers(node_features, graph, packed, started, gather_output)

    def _run_autograd(self, node_features, edge_index, packed, gather_output):
        """The sharded forward with the autograd graph recorded (f16 engine, hidden 128): `loss.backward()` on a loss
        over the rank's own rows leaves this rank's SHARE of every parameter gradient in `.grad`;
        `allreduce_gradients()` sums the shares.  Layers exchange fp16 rows over the collective transport (training
        keeps one fp16 table per layer for the backward instead of the two ping-pong tables of inference)."""
        from . import _native, autograd
        m = self.model
        N, d, lo, hi = self.num_nodes, m.hidden_dim, self.lo, self.hi
        if m.training and m.dropout > 0.0:
            raise NotImplementedError("the multi-GPU path has no dropout")
        if m._precision_code() != _native.PREC_F16 or d != 128:
            raise NotImplementedError("gradients on the multi-GPU path need the f16 engine at hidden_dim 128")
        if gather_output:
            raise NotImplementedError("gradients on the multi-GPU path: the loss is taken over the rank's own rows")
        device = node_features.device
        graph = _native.Graph(edge_index, packed.rel_ids, N, max(packed.num_unique, 1), d, dst_lo=lo, dst_hi=hi,
                              edge_ids=packed.subset)
        # the rank's edges reversed (destination = the original source, any row of the graph): the graph the message
        # term of dL/dh is contracted over
        if packed.subset is not None:                    # forward_packed: the selected edges, relation ids alike
            ei_own, rel_own = edge_index[:, packed.subset.long()], packed.rel_ids
        elif (lo, hi) != (0, N):                         # forward / forward_ids: one relation id per edge of the list
            own = ((edge_index[1] >= lo) & (edge_index[1] < hi)).nonzero().squeeze(1)
            ei_own, rel_own = edge_index[:, own], packed.rel_ids[own].contiguous()
        else:
            ei_own, rel_own = edge_index, packed.rel_ids
        rev = _native.Graph(ei_own.flip(0).contiguous(), rel_own, N, max(packed.num_unique, 1), d)
        x_local = self._local_rows(node_features)
        h = autograd.linear(x_local, m.input_proj.weight, m.input_proj.bias, relu=True)
        text_embs = m.text_encoder.encode_packed(packed)
        rows_pad = max(self.rows * self.world, N, 1)
        for l in range(m.num_layers):
            # fp16 table of h_l: own rows converted with a scale all ranks agree on, then exchanged
            table = torch.zeros((rows_pad, d), dtype=torch.float16, device=device)
            h16 = _native.Shadow(table[:N], torch.zeros(2, dtype=torch.float32, device=device))
            with torch.no_grad():
                if hi > lo:
                    _native.absmax(h.detach(), h16)
                dist.all_reduce(h16.scale[1:2], op=dist.ReduceOp.MAX, group=self.group)
                _native.to_f16(h.detach(), h16.rows(lo, hi), have_amax=True)
                exchange_rows(table, self.ranges, self.rank, self.group)
            w = m._generate(l, text_embs, packed.num_unique)
            ln = m.layer_norms[l]
            h = _ShardedLayerFn.apply(h, w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, self, graph, rev,
                                      ln.eps, h16, None)
        self.num_kept = graph.num_kept
        return h

    def allreduce_gradients(self) -> None:
        """Sum every parameter's `.grad` over the ranks (each rank's backward leaves its share; parameters a rank's
        graph never touched contribute zeros)."""
        for p in self.model.parameters():
            if p.requires_grad:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                dist.all_reduce(p.grad, group=self.group)

    def _decide_push(self, started, num_kept: int, num_phases: int, edge_index, subset) -> None:
        """Pick the push mode of this forward (once) and, for "kernel", exchange the read masks."""
        self.push_used = self._pick_push(num_kept, num_phases) if self.push == "auto" else self.push
        started["push"] = self.push_used
        if self.push_used == "kernel":
            started["peer_mask"] = self._peer_mask(edge_index, subset)

    def _pick_push(self, num_kept: int, num_phases: int) -> str:
        """Per layer: "kernel" costs compute + f X, "copy" costs max(compute, X) + min(compute, X) / chunks, with
        X = time to receive everybody's rows, f = share of rows a peer really reads ~ 1 - exp(-edges per rank / N)
        (uniform sources), compute ~ edges per rank x t_edge.  c3 on 8 GPUs: 0.38 + 0.55 x 0.80 < 0.80 + 0.38 -> kernel;
        c5 on 8 GPUs: 5.9 + 0.71 x 8.0 > 8.0 + 5.9 / 4 -> copy."""
        import math
        d, n = self.model.hidden_dim, max(self.num_nodes, 1)
        compute = num_kept * self._T_EDGE_128 * d / 128.0
        x = (self.world - 1) / self.world * n * d * 2 / self._NVLINK_BPS
        f = 1.0 - math.exp(-num_kept / n)
        chunks = len(self._phase_chunks(num_phases))
        return "kernel" if compute + f * x < max(compute, x) + min(compute, x) / chunks else "copy"

    def _peer_mask(self, edge_index, subset) -> torch.Tensor:
        """uint8 [world, local rows]: entry [q, r] says rank q gathers this rank's row r (it is the source of one of
        q's edges).  Every rank marks the sources of its own edges and one all-to-all hands each range to its owner."""
        from . import _native
        need = _native.mark_rows(edge_index[0], self.num_nodes, subset)
        sizes = [hi - lo for lo, hi in self.ranges]
        n_local = self.hi - self.lo
        got = torch.empty(self.world * n_local, dtype=torch.uint8, device=need.device)
        dist.all_to_all_single(got, need, output_split_sizes=[n_local] * self.world, input_split_sizes=sizes,
                               group=self.group)
        return got.view(self.world, n_local)

    def _layers(self, node_features, graph, packed, started, gather_output):
        self.num_kept = graph.num_kept
        if started is not None:
            local = self._layers_f16(graph, packed, started)
        else:
            local = self._layers_generic(node_features, graph, packed)
        return self._gather_output(local) if gather_output else local

    def _layers_generic(self, node_features, graph, packed) -> torch.Tensor:
        """fp32 / tf32 engines (and hidden sizes without an f16 engine): fp32 rows of h travel, over the collective."""
        from . import _native
        m = self.model
        N, d = self.num_nodes, m.hidden_dim
        prec = m._precision_code()
        if self._bufs is None:
            shape = (max(self.rows * self.world, N), d)
            self._bufs = [torch.zeros(shape, dtype=torch.float32, device=node_features.device) for _ in range(2)]
        cur, nxt = self._bufs
        with torch.no_grad():
            if node_features.shape[0] == N:               # every rank projects all nodes: nothing to exchange
                cur[:N] = _native.linear(node_features, m.input_proj.weight, m.input_proj.bias, relu=True)
            else:
                if self.hi > self.lo:
                    cur[self.lo:self.hi] = _native.linear(self._local_rows(node_features), m.input_proj.weight,
                                                          m.input_proj.bias, relu=True)
                exchange_rows(cur, self.ranges, self.rank, self.group)
            text_embs = m.text_encoder.encode_packed(packed)
            out = None
            all_w = m._generate_all(text_embs, packed.num_unique)
            for l in range(m.num_layers):
                w = all_w[l]
                ln = m.layer_norms[l]
                last = l + 1 == m.num_layers
                if self.hi > self.lo:
                    out = nxt[self.lo:self.hi]
                    graph.mp_layer(cur[:N], w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, ln.eps, prec, out=out)
                if not last:
                    exchange_rows(nxt, self.ranges, self.rank, self.group)
                cur, nxt = nxt, cur
        self._bufs = [cur, nxt]
        d_out = cur[self.lo:self.hi] if self.hi > self.lo else cur[:0]
        return d_out.clone()                              # the buffers are reused by the next call

    def _start_h0(self, node_features):
        """f16 engine, layer-0 input: project this rank's rows, agree on one scale, convert, start the exchange of the
        fp16 rows.  -> state for `_layers_f16`, or None when the path is not the f16 one."""
        from . import _native
        m = self.model
        if not self._f16_engine():
            return None
        device = _native.require_cuda(node_features, m.input_proj.weight)
        N, d, lo, hi = self.num_nodes, m.hidden_dim, self.lo, self.hi
        sym = self._symmetric(device)
        if sym is not None:
            tables = sym.bufs
        else:
            if getattr(self, "_bufs16", None) is None:
                shape = (max(self.rows * self.world, N, 1), d)
                self._bufs16 = [torch.zeros(shape, dtype=torch.float16, device=device) for _ in range(2)]
            tables = self._bufs16
        scales = [torch.zeros(2, dtype=torch.float32, device=device) for _ in range(2)]
        cur16 = _native.Shadow(tables[0][:N] if N else tables[0][:0], scales[0])
        with torch.no_grad():
            h_local = torch.empty((hi - lo, d), dtype=torch.float32, device=device)
            if hi > lo:
                _native.linear(self._local_rows(node_features), m.input_proj.weight, m.input_proj.bias, relu=True,
                               out=h_local)
                _native.absmax(h_local, cur16)
            # one scale for the whole shadow: the ranks agree on max |h0| first (4 bytes, stream-ordered).  This
            # collective also orders this forward after every rank's previous one (the tables are reused).
            dist.all_reduce(cur16.scale[1:2], op=dist.ReduceOp.MAX, group=self.group)
            _native.to_f16(h_local, cur16.rows(lo, hi), have_amax=True)      # writes the scale even with no rows
            pending = []
            if sym is not None:
                sym.push(0, lo, hi)
            else:
                pending = exchange_rows(tables[0], self.ranges, self.rank, self.group, async_op=True)
        return {"tables": tables, "scales": scales, "h_local": h_local, "pending": pending, "sym": sym}

    def _start_generators(self, packed, device, after=None):
        """Text encoder + every layer's generated weights in one native call on the side stream (they depend on the
        relation texts only).  `after`: event the side stream waits for (default: everything enqueued on the current
        stream so far).  -> (weights per layer, event after the last of them)."""
        m = self.model
        main = torch.cuda.current_stream(device)
        if getattr(self, "_gen_stream", None) is None:
            self._gen_stream = torch.cuda.Stream(device=device)
        with torch.no_grad():
            if after is not None:
                self._gen_stream.wait_event(after)
            else:
                self._gen_stream.wait_stream(main)
            with torch.cuda.stream(self._gen_stream):
                text_embs = m.text_encoder.encode_packed(packed)
                weights = m._generate_all(text_embs, packed.num_unique)
                done = torch.cuda.Event()
                done.record(self._gen_stream)
            for w in weights:                             # the tensors are consumed on the main stream
                for t in w.values():
                    t.record_stream(main)
        return weights, done

    def _layers_f16(self, graph, packed, st) -> torch.Tensor:
        """f16 engine: fp32 rows stay local, the fp16 shadow is what every rank reads and what travels.  The rank's
        super-blocks are run in `chunks` pieces; the finished rows of a piece are pushed while the next one runs.
        The generators of all layers run on a side stream (`_start_generators`, enqueued before graph build)."""
        from . import _native
        m = self.model
        N, d, lo, hi = self.num_nodes, m.hidden_dim, self.lo, self.hi
        tables, scales, sym = st["tables"], st["scales"], st["sym"]
        h_cur, pending = st["h_local"], st["pending"]
        device = h_cur.device
        main = torch.cuda.current_stream(device)
        with torch.no_grad():
            weights, done = st["weights"] if "weights" in st else self._start_generators(packed, device)
            ready = [done] * m.num_layers
            chunks = self._phase_chunks(graph) if graph.num_local else []
            self._mark("prep")
            for l in range(m.num_layers):
                ln, w = m.layer_norms[l], weights[l]
                last = l + 1 == m.num_layers
                cb, nb = l % 2, (l + 1) % 2
                # the rows of h_l are complete everywhere
                if sym is not None:
                    sym.close_layer(cb)
                else:
                    for work in pending:
                        work.wait()
                main.wait_event(ready[l])
                self._mark("wait")
                h_nxt = torch.empty_like(h_cur)
                h16 = _native.Shadow(tables[cb][:N], scales[cb])
                out16 = None if last or hi <= lo else _native.Shadow(tables[nb][lo:hi], scales[nb])
                mask = st.get("peer_mask")
                push = None
                if not last and sym is not None and mask is not None and hi > lo:
                    push = _native.PeerPush(mask, sym.table_ptrs[nb], self.rank)
                    self.rows_needed_by_peers = mask
                for p_lo, p_hi in ([(0, graph.num_phases)] if push is not None else chunks):
                    graph.mp_layer(h_cur, w["W_msg"], w["W_self"], w["bias"], ln.weight, ln.bias, ln.eps, _native.PREC_F16,
                                   out=h_nxt, h16=h16, out16=out16, h_row0=lo, phases=(p_lo, p_hi), push=push)
                    if not last and sym is not None and push is None:   # these rows travel while the next piece runs
                        r0, r1 = graph.phase_rows(p_lo, p_hi)
                        sym.push(nb, lo + r0, lo + r1)
                if not last and sym is None:
                    pending = exchange_rows(tables[nb], self.ranges, self.rank, self.group, async_op=True)
                self._mark("compute")
                h_cur = h_nxt
        return h_cur

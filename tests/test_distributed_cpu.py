"""Host-side logic of the multi-GPU path on CPU: destination-range planning and the in-place all-gather
(gloo, world_size 2).  Per-rank compute is done by the oracle restricted to the rank's destinations, so what
is tested is exactly what distributed.ShardedForward adds around the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hypergnn_oracle as O


def test_plan_partition_covers_nodes_once():
    from graph_hypernetwork_forge.distributed import plan_partition
    for n, w in ((10, 1), (10, 2), (10, 3), (7, 8), (0, 4), (2_500_000, 8), (16, 4)):
        rows, ranges = plan_partition(n, w)
        assert len(ranges) == w and rows * w >= n
        covered = []
        for lo, hi in ranges:
            assert 0 <= lo <= hi <= n and hi - lo <= rows
            covered += list(range(lo, hi)) if n < 100 else []
        if n < 100:
            assert covered == list(range(n))
        assert ranges[0][0] == 0 and max(hi for _, hi in ranges) == n
    with pytest.raises(ValueError):
        plan_partition(5, 0)


def test_plan_partition_by_edges_balances_work():
    from graph_hypernetwork_forge.distributed import plan_partition_by_edges
    rng = np.random.default_rng(0)
    n = 5000
    indeg = (rng.zipf(1.3, n) - 1).clip(0, 20000)            # power-law in-degree: a few hubs
    rowptr = np.concatenate([[0], np.cumsum(indeg)])
    for world in (1, 2, 3, 8):
        ranges = plan_partition_by_edges(rowptr, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])) and all(lo <= hi for lo, hi in ranges)
        work = [rowptr[hi] - rowptr[lo] + (hi - lo) for lo, hi in ranges]
        heaviest_node = indeg.max() + 1
        assert max(work) <= sum(work) / world + heaviest_node   # no range exceeds its share by more than one node
    assert plan_partition_by_edges(np.array([0]), 4) == [(0, 0)] * 4          # no nodes at all
    equal = plan_partition_by_edges(np.arange(0, 81, 1) * 7, 4)               # uniform degrees -> equal node counts
    assert [hi - lo for lo, hi in equal] == [20, 20, 20, 20]


def _exchange_worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "graph-hypernetwork-forge_b200")]
    from graph_hypernetwork_forge.distributed import exchange_rows, plan_partition
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d = 11, 3
    want = torch.arange(n * d, dtype=torch.float32).reshape(n, d)
    results = []
    rows, equal = plan_partition(n, world)
    for ranges, size in ((equal, rows * world), ([(0, 7), (7, 11)], n), ([(0, 0), (0, 11)], n)):
        buf = torch.full((size, d), -1.0)
        lo, hi = ranges[rank]
        buf[lo:hi] = want[lo:hi]
        for work in exchange_rows(buf, ranges, rank, None, async_op=True):
            work.wait()
        results.append(bool(torch.equal(buf[:n], want)))
    ret[rank] = results
    dist.destroy_process_group()


def test_exchange_rows_even_and_uneven_ranges_gloo():
    """The collective transport: in-place all-gather for equal padded ranges, one broadcast per rank otherwise
    (edge-balanced ranges, empty ranges)."""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_exchange_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert all(all(ret[r]) for r in range(world)), dict(ret)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "graph-hypernetwork-forge_b200"), os.path.join(root, "tests")]
    from graph_hypernetwork_forge.distributed import gather_rows, plan_partition
    from _util import build_model, load_case, model_params_numpy
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = load_case("synth_small")
    P = model_params_numpy(build_model(case))
    d, L = case["ctor"]["hidden_dim"], case["ctor"]["num_layers"]
    x, ei, texts = case["node_features"], case["edge_index"], case["edge_texts"]
    N = x.shape[0]
    rows, ranges = plan_partition(N, world)
    lo, hi = ranges[rank]
    src, dst = ei
    keep = (dst >= lo) & (dst < hi)                       # this rank's edges: destinations it owns
    unique, rel = O.dedup_texts([t for t, k in zip(texts, keep) if k])   # rank-local relation ids
    text = O.text_encode(unique, P["text_encoder.char_emb.weight"], P["text_encoder.proj.0.weight"],
                         P["text_encoder.proj.0.bias"])
    bufs = [torch.zeros(rows * world, d), torch.zeros(rows * world, d)]
    bufs[0][:N] = torch.from_numpy(np.maximum(x @ P["input_proj.weight"].T + P["input_proj.bias"], 0))
    for l in range(L):
        w = O.weight_generator(text, P, f"weight_generators.{l}.", d, d)
        h = bufs[0][:N].numpy()
        upd = O.message_passing(h, src[keep], dst[keep], rel, w["W_msg"], w["W_self"], w["bias"], num_nodes=N)
        out = O.layer_norm(np.maximum(upd + h, 0), P[f"layer_norms.{l}.weight"], P[f"layer_norms.{l}.bias"])
        bufs[1][lo:hi] = torch.from_numpy(out[lo:hi].astype(np.float32))
        gather_rows(bufs[1], rows, rank, None)
        bufs.reverse()
    if rank == 0:
        ret["out"] = bufs[0][:N].numpy().copy()
    dist.destroy_process_group()


def test_sharded_forward_matches_single_process_gloo():
    from _util import load_case
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        got = ret["out"]
    want = load_case("synth_small")["taps"]["out"]
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5)

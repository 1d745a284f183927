#!/usr/bin/env python
"""bench.py - HyperGNN forward, edges/s per layer, on the synthetic wikikg2-shaped KG (BASELINE config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|c4]

One "step" is one complete forward pass over the workload: relation dedup + text encoder +
input projection + graph build + L x (weight generation + message passing + LayerNorm epilogue).
  value  : E * L / step time with inputs resident in HBM (device-timed, CUDA events)
  e2e    : the same through the C ABI entry ghf_hypergnn_forward_host with HOST buffers
           (H2D of features / edges / strings and D2H of the embeddings inside the timed region)
  roofline: the dominant kernel (the fused layer kernel at hidden 128), SURVEY 8(d) algorithmic bytes / its
           event-timed duration; also the bytes at the operand width the kernel really reads, the whole layer
           and the whole step against the same roofline
  precision_alt: the same step on the tf32 and fp32 engines (the fp32-tolerance context of an f16 headline)
  cpu_baseline: the reference's CPU path on a bounded sample of the same workload, host cores
`--impl reference` times the UNMODIFIED reference (installed under baseline/_ref by baseline/install_reference.sh,
git-ignored, travels with the snapshot) on the host cores, on a sample of the workload with N and E scaled
together; without that install it falls back to the numpy port of the reference algorithm (oracle/).  Both
arms print the same `config`.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # T = 64 (so generator hidden = 128) and F = d for c2-c4
    "c2": dict(name="fb15k237-shaped", N=14_541, E=272_115, R=237, d=128, L=2, T=64, F=128),
    "c3": dict(name="wikikg2-shaped", N=2_500_000, E=16_000_000, R=535, d=128, L=3, T=64, F=128),
    # one eighth of BASELINE config 5 (50M nodes, 500M edges over 8 GPUs): the hidden-64 kernels at a rank's edge count
    "c5s": dict(name="large-shape-1/8", N=6_250_000, E=62_500_000, R=1000, d=64, L=2, T=64, F=64),
    "c4": dict(name="zero-shot-20k-rel", N=100_000, E=2_000_000, R=20_000, d=256, L=2, T=64, F=256),
    # BASELINE config 5: dst-partitioned across the GPUs of one box; every rank generates its own shard (features of
    # its rows, the edges pointing into them) and goes in through the ids-in entry - no global edge list exists
    "c5": dict(name="large-50M-500M", N=50_000_000, E=500_000_000, R=1000, d=64, L=2, T=64, F=64, sharded_inputs=True),
}
NAME_LEN = 14  # len("relation_00000")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        if os.environ.get("GHF_BENCH_NO_CLOCKS"):       # diagnostic: run without the sampler process
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi needs a few hundred milliseconds to initialise NVML and attach to the GPU, and CUDA calls
            # of this process can stall meanwhile: wait for its first sample so that none of that falls into the
            # timed region (a 5-step warm-up is far shorter than the start-up)
            t_end = time.time() + 5.0
            while not self.rows and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def window(self, t0, t1):
        """Keep the samples taken inside [t0, t1] (wall clock); if the region was shorter than the sampling
        period, keep the samples closest to it."""
        inside = [r[1:] for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05]
        if not inside and self.rows:
            mid = 0.5 * (t0 + t1)
            inside = [min(self.rows, key=lambda r: abs(r[0] - mid))[1:]]
        self.rows = inside

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        mhz = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz)}


def make_device_inputs(w, device, seed=0, skew=False):
    """Synthetic KG of the named shape, created on the device.  Headline: uniform src/dst/rel (the worst
    case for relation grouping).  skew=True: destinations and relations follow Zipf(1) (rank = n^u, u uniform; node
    ranks scattered over the id range by a fixed permutation) - closer to real knowledge graphs, secondary."""
    g = torch.Generator(device=device).manual_seed(seed)
    N, E, R, F = w["N"], w["E"], w["R"], w["F"]
    edge_index = torch.randint(0, N, (2, E), generator=g, device=device, dtype=torch.int64)
    rel = torch.randint(0, R, (E,), generator=g, device=device, dtype=torch.int64)
    if skew:
        node_of_rank = torch.randperm(N, generator=g, device=device)
        u = torch.rand(E, generator=g, device=device, dtype=torch.float64)
        edge_index[1] = node_of_rank[(torch.pow(float(N), u).long() - 1).clamp_(0, N - 1)]
        u = torch.rand(E, generator=g, device=device, dtype=torch.float64)
        rel = (torch.pow(float(R), u).long() - 1).clamp_(0, R - 1)
    x = torch.randn(N, F, generator=g, device=device, dtype=torch.float32)
    names = np.frombuffer("".join(f"relation_{r:05d}" for r in range(R)).encode(), dtype=np.uint8).reshape(R, NAME_LEN)
    utf8 = torch.from_numpy(names.copy()).to(device)[rel].reshape(-1).contiguous()
    offsets = torch.arange(E + 1, device=device, dtype=torch.int64) * NAME_LEN
    return x, edge_index, rel, utf8, offsets


def make_shard_inputs(w, device, lo, hi, n_edges, seed):
    """One rank's shard of a pre-sharded workload: features of rows [lo, hi), `n_edges` edges with uniform sources
    over all N nodes and uniform destinations in [lo, hi), uniform relation ids.  Seeded per shard, so any rank can
    regenerate any shard (the parity check on rank 0 does)."""
    g = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, w["N"], (n_edges,), generator=g, device=device, dtype=torch.int64)
    dst = torch.randint(lo, max(hi, lo + 1), (n_edges,), generator=g, device=device, dtype=torch.int64)
    rel = torch.randint(0, w["R"], (n_edges,), generator=g, device=device, dtype=torch.int32)
    x = torch.randn(hi - lo, w["F"], generator=g, device=device, dtype=torch.float32)
    return x, torch.stack([src, dst]), rel


def build_model(w, device, precision):
    from graph_hypernetwork_forge import HyperGNN
    torch.manual_seed(0)
    return HyperGNN(w["T"], w["F"], w["d"], w["L"], precision=precision).eval().to(device)


def algorithmic_bytes(w, E, N_local, N):
    """fp32 features, int32 ids.  contraction = row gathers + ids + h[dst] once + weights once;
    layer = contraction + write h' + rowptr/in-degree."""
    d, R = w["d"], w["R"]
    contraction = E * (4 * d + 8) + N_local * 4 * d + R * (2 * d * d + d) * 4
    layer = contraction + N_local * 4 * d + 4 * (N_local + 1)
    return contraction, layer


def bytes_as_read(w, E, N_local, precision, fused):
    """The same accounting at the operand width the engine really moves (second figure next to SURVEY's fp32 one):
    f16 engines gather fp16 rows and stream fp16 weight images; the fused layer kernel also reads the fp32 residual
    row and writes the fp32 row + its fp16 shadow; the unfused contraction writes and re-reads an fp32 accumulator."""
    d, R = w["d"], w["R"]
    row = 2 * d if precision == "f16" else 4 * d
    wbytes = R * (2 * d * d) * (2 if precision == "f16" else 4) + R * d * 4
    gathers = E * (2 * row + 8)                              # source row + destination row + two ids per edge
    contraction = gathers + wbytes + (0 if fused else N_local * 4 * d)
    layer = gathers + wbytes + N_local * (4 * d + 4 * d + 4) + (N_local * 2 * d if precision == "f16" else 0) \
        + (0 if fused else 2 * N_local * 4 * d)
    return contraction, layer


def make_config(w, world, skew):
    """`config` of the JSON line: identical in the b200 arm and the reference arm (the driver compares them)."""
    N, E, d, L = w["N"], w["E"], w["d"], w["L"]
    return {"workload": f"{w['name']} N={N} E={E} R={w['R']} d={d} L={L} T={w['T']} F={w['F']}"
                        + (" zipf-dst-rel" if skew else " uniform"),
            "step": "dedup + text encoder + input projection + graph build + L layers",
            "l2": (f"inputs larger than L2 (h {N * d * 4 / 1e9:.2f} GB, edges {E * 16 / 1e9:.2f} GB); "
                   "no explicit flush") if N * d * 4 > 256e6 else
                  "workload smaller than L2: 256 MB scratch written between steps",
            "parallelism": "single GPU" if world == 1 else
                           f"dst-range x{world}; fp16 rows of h exchanged per layer over NVLink; output stays sharded"}


def sampled_workload(w, edges, seed=0):
    """A bounded sample of the workload for the CPU legs: N and E scaled TOGETHER (same in-degree, same relation
    vocabulary, same generator), so per-node work (projection, LayerNorm) and per-edge work keep their proportion.
    One sampler for `cpu_baseline` and `--impl reference`."""
    from oracle import hypergnn_oracle as O
    k = int(min(w["E"], max(1, edges)))
    n = int(max(64, min(w["N"], round(w["N"] * k / w["E"]))))
    src, dst, rel, names, feats = O.synthetic_kg(n, k, w["R"], w["F"], seed=seed)
    return n, k, src, dst, [names[r] for r in rel], feats


REF_DIR = os.path.join(ROOT, "baseline", "_ref")
REF_EDGE_RATE = 3.0e4     # edge-layers/s of the literal reference on ~8 cores: sizes the sample only
REF_MAX_EDGES = 60_000    # the literal reference needs ~268 KB per edge at d = 128 (per-edge weight gathers)


def time_literal_reference(w, edges, steps, warmup):
    """The UNMODIFIED reference (baseline/_ref) on the host cores: model(x, edge_index, List[str]), eval, no_grad."""
    import importlib
    for name in [m for m in sys.modules if m.split(".")[0] == "graph_hypernetwork_forge"]:
        del sys.modules[name]                      # the drop-in shares the import path: make sure the reference loads
    sys.path[:] = [REF_DIR] + [q for q in sys.path if "graph-hypernetwork-forge_b200" not in q]
    ref = importlib.import_module("graph_hypernetwork_forge")
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REF_DIR)), ref.__file__
    torch.set_num_threads(os.cpu_count() or 1)
    n, k, src, dst, texts, feats = sampled_workload(w, edges)
    torch.manual_seed(0)
    model = ref.HyperGNN(w["T"], w["F"], w["d"], w["L"]).eval()
    x, ei = torch.from_numpy(feats), torch.from_numpy(np.stack([src, dst]))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(x, ei, texts)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return n, k, times


def time_port(w, edges, steps, warmup):
    """The numpy port of the reference algorithm (oracle/) on the same kind of sample."""
    from oracle import hypergnn_oracle as O
    n, k, src, dst, texts, feats = sampled_workload(w, edges)
    torch.manual_seed(0)
    sys.path.insert(0, os.path.join(ROOT, "graph-hypernetwork-forge_b200"))
    from graph_hypernetwork_forge import HyperGNN
    params = {kk: v.numpy() for kk, v in HyperGNN(w["T"], w["F"], w["d"], w["L"]).state_dict().items()}
    ei = np.stack([src, dst])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.hypergnn_forward(params, feats, ei, texts, w["d"], w["L"])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n, k, times


def run_reference(args, w):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, a bounded sample of
    the workload per step (sampled_workload).  Rank 0 alone runs; the other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    L = w["L"]
    budget = min(15.0, 120.0 / max(1, args.steps + args.warmup))      # seconds of CPU work per step
    literal = os.path.isdir(os.path.join(REF_DIR, "graph_hypernetwork_forge")) and not args.port
    port = None
    if not args.no_port_extra or not literal:
        # the memory-feasible port, one pass on a larger sample (it does not materialise per-edge weights)
        pn, pk, pt = time_port(w, args.sample_edges or int(min(w["E"], 4.0e5 * budget / 5.0)), 1 if literal else args.steps,
                               0 if literal else args.warmup)
        pms = 1e3 * sum(pt) / len(pt)
        port = {"value": pk * L / (pms / 1e3), "unit": "edges/s/layer", "cores": cores, "kind": "port",
                "ms_per_step": pms,
                "sample": f"{pk} edges, {pn} nodes (N and E scaled together), {L} layers: numpy port of the reference "
                          "algorithm (oracle/hypergnn_oracle.py)"}
    if literal:
        k = args.sample_edges or int(max(5_000, min(REF_MAX_EDGES, budget * REF_EDGE_RATE / L)))
        n, k, times = time_literal_reference(w, k, args.steps, args.warmup)
        ms = 1e3 * sum(times) / len(times)
        base = {"value": k * L / (ms / 1e3), "unit": "edges/s/layer", "cores": cores, "kind": "reference",
                "sample": f"{k} edges, {n} nodes (N and E of the workload scaled together, same R, same generator), "
                          f"{L} layers per step: the unmodified reference from baseline/_ref, "
                          f"torch {torch.__version__} CPU, {torch.get_num_threads()} threads",
                "port": port}
    else:
        ms, base = port["ms_per_step"], dict(port)
    line = {"impl": "reference", "metric": "hypergnn_fwd_edges_per_sec_per_layer", "value": base["value"],
            "unit": "edges/s/layer", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": make_config(w, args.gpus, args.skew), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "edges/s/layer", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    """The reference arm in a child process (the reference and the drop-in share an import path), one step."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
           "--warmup", "1", "--scale", str(args.scale)] + (["--skew"] if args.skew else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)    # 20 x ~12 ms at c3: single slow steps (allocator, clocks) average out
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "f16", "tf32", "fp32"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink N and E (debugging only; not a bench number)")
    ap.add_argument("--skew", action="store_true", help="Zipf destinations/relations (secondary workload)")
    ap.add_argument("--python-path", action="store_true",
                    help="enqueue the stages from Python (prepare_packed + forward_prepared) instead of one native call")
    ap.add_argument("--train", action="store_true",
                    help="also time a training step (forward + backward on a prepared graph) -> key 'train_step'")
    ap.add_argument("--chunks", type=int, default=0, help="multi-GPU, --push copy: pieces of a rank's rows whose exchange overlaps the next piece (0 = auto)")
    ap.add_argument("--push", default="auto", choices=["auto", "kernel", "copy"],
                    help="multi-GPU p2p: epilogue kernel stores rows to the peers that read them | whole ranges by copy engines")
    ap.add_argument("--transport", default=None, choices=["p2p", "collective"], help="multi-GPU row exchange")
    ap.add_argument("--balance", default="nodes", choices=["nodes", "edges"], help="multi-GPU: destination ranges by node count or by edges")
    ap.add_argument("--no-check", action="store_true", help="multi-GPU: skip the comparison with a single-GPU forward")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the tf32 / fp32 engines (precision_alt)")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the numpy port even if baseline/_ref exists")
    ap.add_argument("--no-port-extra", action="store_true", help="--impl reference: skip the extra pass of the numpy port")
    ap.add_argument("--sample-edges", type=int, default=0, help="--impl reference: edges of the per-step sample")
    args = ap.parse_args()
    torch.set_grad_enabled(False)   # inference benchmark: no autograd graph is recorded
    w = dict(WORKLOADS[args.workload])
    if args.scale != 1.0:
        w["N"], w["E"] = max(64, int(w["N"] * args.scale)), max(64, int(w["E"] * args.scale))
    if args.impl == "reference":
        return run_reference(args, w)

    from graph_hypernetwork_forge import _native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    dist = None
    nccl_env = {}
    if world > 1:
        # NVLink only (north_star): no InfiniBand, P2P through NVLink; NCCL's INFO lines go to a file that rank 0
        # summarises in the JSON line (stdout must stay one JSON line)
        for k, v in (("NCCL_IB_DISABLE", "1"), ("NCCL_P2P_LEVEL", "NVL"), ("NCCL_DEBUG", "INFO"),
                     ("NCCL_DEBUG_SUBSYS", "INIT,GRAPH"),
                     ("NCCL_DEBUG_FILE", f"/tmp/ghf_nccl_{os.getpid()}.log")):
            if k.startswith("NCCL_DEBUG") and not os.environ.get("GHF_KEEP_NCCL_DEBUG"):
                os.environ[k] = v                         # the image presets NCCL_DEBUG=VERSION
            else:
                os.environ.setdefault(k, v)
        nccl_env = {k: os.environ[k] for k in ("NCCL_IB_DISABLE", "NCCL_P2P_LEVEL", "NCCL_DEBUG", "NCCL_DEBUG_FILE")}
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    precision = args.precision
    if precision == "auto":
        precision = "f16" if w["d"] in (64, 128, 256) else "tf32"
    if w["d"] not in (32, 64, 128, 256):
        precision = "fp32"
    model = build_model(w, device, precision)
    N, E, L, d = w["N"], w["E"], w["L"], w["d"]
    pre_sharded = bool(w.get("sharded_inputs"))
    names = [f"relation_{r:05d}" for r in range(w["R"])]
    if not pre_sharded:
        x, edge_index, _rel, utf8, offsets = make_device_inputs(w, device, skew=args.skew)

    sharded = None
    if world == 1 and not pre_sharded:
        def step():
            if args.python_path:    # the same stages enqueued one by one from Python (~60 native calls per forward)
                with torch.no_grad():
                    return model.forward_prepared(x, model.prepare_packed(edge_index, utf8, offsets, N))
            return model.forward_packed(x, edge_index, utf8, offsets)   # one native call per forward
        n_local = N
    elif world == 1:
        # a pre-sharded workload on one GPU: the single shard is the whole graph, through the ids-in entry
        x, edge_index, rel_ids = make_shard_inputs(w, device, 0, N, E, seed=1000)
        n_local = N

        def step():
            return model.forward_prepared(x, model.prepare_ids(edge_index, rel_ids, names, N))
    else:
        from graph_hypernetwork_forge.distributed import ShardedForward, plan_partition, plan_partition_by_edges
        ranges = None
        if args.balance == "edges" and not pre_sharded:
            indeg = torch.bincount(edge_index[1], minlength=N)
            rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
            rowptr[1:] = torch.cumsum(indeg, 0)
            ranges = plan_partition_by_edges(rowptr, world)
        sharded = ShardedForward(model, N, dist.group.WORLD, ranges=ranges, transport=args.transport, chunks=args.chunks,
                                 push=args.push)
        n_local = sharded.hi - sharded.lo
        if pre_sharded:
            per = [E // world + (1 if r < E % world else 0) for r in range(world)]
            x, edge_index, rel_ids = make_shard_inputs(w, device, sharded.lo, sharded.hi, per[rank], seed=1000 + rank)

            def step():
                return sharded.forward_ids(x, edge_index, rel_ids, names)
        else:
            def step():
                return sharded.forward_packed(x, edge_index, utf8, offsets)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    flush = torch.empty(64 << 20, dtype=torch.float32, device=device) if N * d * 4 <= 256e6 else None

    def timed_step():
        if flush is not None:      # workload smaller than L2: evict it between steps
            flush.zero_()
        return step()

    # the clock sampler (nvidia-smi -lms) starts BEFORE the warm-up: its start-up (process spawn, NVML init) must not
    # fall into the timed region; only the samples taken inside the region are kept
    with ClockSampler(local) as clocks:
        out = None
        for _ in range(max(args.warmup, 3)):
            # bound to `out` exactly as in the timed loop: the previous result is still alive while the next one is
            # allocated, so the caching allocator reaches its steady state (two 1.28 GB result blocks at c3) during
            # the warm-up - otherwise the SECOND timed step pays a fresh cudaMalloc (27-160 ms observed)
            out = timed_step()
        barrier()
        _native.profile_enable(True)
        _native.profile_read()
        _native.launch_count(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = time.time()
        ev0.record()
        marks = []
        for _ in range(args.steps):
            out = timed_step()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        ev1.record()
        barrier()
        t_wall1 = time.time()
    clocks.window(t_wall0, t_wall1)
    ms = ev0.elapsed_time(ev1) / args.steps
    each = [a.elapsed_time(b) for a, b in zip([ev0] + marks[:-1], marks)]     # per-step device time (diagnostic)
    launches = _native.launch_count()
    prof, n_layers_timed = _native.profile_read()
    _native.profile_enable(False)
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = E * L / (ms / 1e3)

    hbm_peak, peak_src = peaks()
    local_edges = E if sharded is None else sharded.num_kept
    # a rank's layer may be several launches (chunks whose row exchange overlaps the next chunk's compute)
    per_layer = max(1, round(n_layers_timed / max(1, L * args.steps)))
    b_contr, b_layer = algorithmic_bytes(w, local_edges / per_layer, n_local / per_layer, N)
    fused = precision == "f16" and d == 128 and os.environ.get("GHF_MP_FUSED", "0") == "1"
    r_contr, r_layer = bytes_as_read(w, local_edges / per_layer, n_local / per_layer, precision, fused)
    t_contr = prof["contraction_ms"] / max(n_layers_timed, 1)
    t_layer = (prof["contraction_ms"] + prof["epilogue_ms"] + prof["prep_ms"]) / max(n_layers_timed, 1)
    # the dominant kernel: the fused layer kernel does the whole layer, so it is charged the layer's bytes
    b_kernel, r_kernel = (b_layer, r_layer) if fused else (b_contr, r_contr)
    gbs = lambda nbytes, t_ms: nbytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0   # noqa: E731
    achieved = gbs(b_kernel, t_contr)
    traffic = None
    try:  # per-launch DRAM bytes of that kernel from the committed ncu capture, when present (1 GPU, full graph)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{args.workload}:{precision}" + (":fused" if fused else "")) if world == 1 else None
    except Exception:
        pass
    kernel_name = {"f16": ("mp_f16_fused_kernel" if fused else "mp_f16_kernel") if d == 128 else "mp_f16_ss_kernel",
                   "tf32": "mp_umma_ts_kernel" if d == 128 else f"mp_umma_kernel<{d}>", "fp32": "mp_fp32_kernel"}[precision]
    roofline = {"bound": "hbm", "kernel": kernel_name,
                "what": "whole layer (contraction + mean + residual + ReLU + LayerNorm) in one kernel" if fused
                        else "contraction kernel",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b_kernel,
                "ms_per_launch": t_contr,
                "as_read": {"bytes_per_launch": r_kernel, "achieved": gbs(r_kernel, t_contr),
                            "frac": gbs(r_kernel, t_contr) / hbm_peak,
                            "what": "same accounting at the operand width the kernel moves (fp16 rows and weight "
                                    "images on the f16 engine; fp32 residual in, fp32 + fp16 rows out)"},
                "layer": {"algorithmic_bytes": b_layer, "ms": t_layer, "achieved": gbs(b_layer, t_layer),
                          "frac": gbs(b_layer, t_layer) / hbm_peak,
                          "what": "weight-image packing + layer kernel(s), per layer"},
                "step": {"algorithmic_bytes": algorithmic_bytes(w, E, N, N)[1] * L, "ms": ms,
                         "achieved": gbs(algorithmic_bytes(w, E, N, N)[1] * L, ms),
                         "frac": gbs(algorithmic_bytes(w, E, N, N)[1] * L, ms) / (hbm_peak * world),
                         "what": "the whole forward (dedup, text encoder, projection, graph build, generators, L "
                                 "layers) charged only the L layers' algorithmic bytes: value / (E / (B_layer / "
                                 "peak)); the north-star target is 0.60 of this"}}

    # ---- multi-GPU: parity against a single-GPU forward, stage split, transport, end to end
    multi = None
    if sharded is not None:
        out_local = step()
        parity = None
        if not args.no_check:
            # every rank holds the model; rank 0 (any rank for the replicated workloads) runs the same inputs on ONE
            # GPU and compares its own rows.  Pre-sharded: rank 0 regenerates every shard.
            err = torch.zeros(1, device=device, dtype=torch.float64)
            rows = torch.zeros(1, device=device, dtype=torch.float64)
            if not pre_sharded:
                single = model.forward_packed(x, edge_index, utf8, offsets)
                if n_local:
                    err[0] = float((single[sharded.lo:sharded.hi] - out_local).abs().max())
                rows[0] = n_local
                del single
            elif rank == 0:
                from graph_hypernetwork_forge.distributed import plan_partition as _pp
                per = [E // world + (1 if r < E % world else 0) for r in range(world)]
                shards = [make_shard_inputs(w, device, lo_r, hi_r, per[r], seed=1000 + r)
                          for r, (lo_r, hi_r) in enumerate(sharded.ranges)]
                xf = torch.cat([t[0] for t in shards])
                eif = torch.cat([t[1] for t in shards], dim=1)
                relf = torch.cat([t[2] for t in shards])
                del shards
                single = model.forward_prepared(xf, model.prepare_ids(eif, relf, names, N))
                err[0] = float((single[sharded.lo:sharded.hi] - out_local).abs().max())
                rows[0] = n_local
                del single, xf, eif, relf
                torch.cuda.empty_cache()
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            dist.all_reduce(rows)
            parity = {"max_abs_diff": float(err.item()), "rows_compared": int(rows.item()),
                      "against": "the single-GPU forward of the same inputs" +
                                 (" (rank 0 regenerates all shards; its own rows)" if pre_sharded else " (every rank, its own rows)"),
                      "tolerance": 1.5e-4, "ok": bool(err.item() <= 1.5e-4)}
        pushed = None
        if sharded._sym is not None:
            sharded._sym.bytes_pushed = 0
            out_local = step()
            pushed = sharded._sym.bytes_pushed
        sharded.profile = True
        for _ in range(3):
            out_local = step()
        stages = {k: v / 3 for k, v in sharded.stage_ms().items()}
        sharded.profile = False
        st = torch.tensor([stages["prep"], stages["compute"], stages["wait"]], device=device, dtype=torch.float64)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        # end to end: the rank's inputs in pinned host memory -> device, forward, own rows back to the host
        e2e_multi = None
        if not args.no_e2e:
            host = [t.cpu().pin_memory() for t in ((x, edge_index, rel_ids) if pre_sharded else
                                                     (x[sharded.lo:sharded.hi], edge_index, utf8, offsets))]
            hout = torch.empty(n_local, d, dtype=torch.float32).pin_memory()

            def e2e_step():
                dev_in = [t.to(device, non_blocking=True) for t in host]
                if pre_sharded:
                    o = sharded.forward_ids(dev_in[0], dev_in[1], dev_in[2], names)
                else:
                    o = sharded.forward_packed(dev_in[0], dev_in[1], dev_in[2], dev_in[3])
                hout.copy_(o, non_blocking=True)
                torch.cuda.synchronize(device)
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            k = max(1, min(args.steps, 3))
            for _ in range(k):
                e2e_step()
            barrier()
            tt = torch.tensor([1e3 * (time.perf_counter() - t0) / k], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            h2d = sum(t.numel() * t.element_size() for t in host)
            e2e_multi = {"value": E * L / (float(tt.item()) / 1e3), "unit": "edges/s/layer", "ms_per_step": float(tt.item()),
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(hout.numel() * 4),
                         "api": "ShardedForward.forward_" + ("ids" if pre_sharded else "packed") +
                                " per rank, pinned host buffers, own rows back; bytes are per rank"}
        transport_lines = []
        try:
            with open(os.environ.get("NCCL_DEBUG_FILE", "")) as f:
                keep = [ln.strip() for ln in f if any(t in ln for t in ("via P2P", "NVLS", "via NET", "via SHM", "Connected", "NCCL version"))]
            transport_lines = keep[:3] + keep[-5:]
        except Exception:
            pass
        sel = sharded.rows_needed_by_peers
        multi = {"transport": sharded.transport, "push": sharded.push_used if sharded.transport == "p2p" else None,
                 "rows_sent_fraction": (float((sel.float().sum() - sel[rank].float().sum()) /
                                              max(1, sel.numel() - sel.shape[1])) if sel is not None else None),
                 "chunks": per_layer, "balance": args.balance,
                 "ranges_rows": [hi_r - lo_r for lo_r, hi_r in sharded.ranges],
                 "edges_per_rank_max": int(local_edges), "parity": parity,
                 "stage_ms_max_over_ranks": {"prep": float(st[0]), "compute": float(st[1]), "exposed_exchange": float(st[2])},
                 "bytes_pushed_per_step_per_rank": pushed,
                 "output": "sharded: every rank keeps its own rows (SURVEY 8e)",
                 "nccl_env": nccl_env, "nccl_log": transport_lines}
        if rank == 0:
            for ln in transport_lines:
                print("[nccl] " + ln, file=sys.stderr)
        e2e_override = e2e_multi
    else:
        e2e_override = None

    # ---- end to end through the C ABI with host buffers (rank 0 of a 1-GPU run)
    e2e = e2e_override
    if world == 1 and not pre_sharded and not args.no_e2e:
        from graph_hypernetwork_forge import _native as nat
        hx = x.cpu().pin_memory()
        hei = edge_index.cpu().pin_memory()
        hutf8 = utf8.cpu().pin_memory()
        hoff = offsets.cpu().pin_memory()
        hout = torch.empty(N, d, dtype=torch.float32).pin_memory()
        desc = nat.ModelDesc(w["T"], w["F"], d, L, 32, max(64, 2 * w["T"]), 2, nat.precision_code(precision), 1e-5)
        params = model.flat_parameters()
        del out
        for _ in range(2):
            nat.forward_host(desc, params, hx, hei, hutf8, hoff, hout, device)
        k = max(1, min(args.steps, 3))
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(k):
            nat.forward_host(desc, params, hx, hei, hutf8, hoff, hout, device)
        torch.cuda.synchronize(device)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / k
        h2d = hx.numel() * 4 + hei.numel() * 8 + hutf8.numel() + hoff.numel() * 8
        e2e = {"value": E * L / (e2e_ms / 1e3), "unit": "edges/s/layer", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(hout.numel() * 4),
               "api": "ghf_hypergnn_forward_host (C ABI, pinned host buffers)"}

    # ---- optional: one training step (forward with the autograd graph + backward through the gradient kernels)
    train = None
    if world == 1 and args.train:
        with torch.enable_grad():
            model.train()
            prepared = model.prepare_packed(edge_index, utf8, offsets, N)
            loss_w = torch.randn(N, d, device=device)

            def train_step():
                model.zero_grad(set_to_none=True)
                (model.forward_prepared(x, prepared) * loss_w).sum().backward()
            for _ in range(2):
                train_step()
            k = max(1, min(args.steps, 5))
            tv0, tv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(device)
            tv0.record()
            for _ in range(k):
                train_step()
            tv1.record()
            torch.cuda.synchronize(device)
            t_ms = tv0.elapsed_time(tv1) / k
            model.eval()
            model.zero_grad(set_to_none=True)
            del prepared, loss_w
        train = {"ms_per_step": t_ms, "value": E * L / (t_ms / 1e3), "unit": "edges/s/layer, forward + backward",
                 "peak_memory_gib": torch.cuda.max_memory_allocated(device) / 2**30,
                 "what": "prepared graph reused; loss = (out * W).sum(); gradients of every parameter"}

    # ---- the graph reused ("edges are sorted once", north_star item 3): dedup + graph build done once, every timed
    # step = input projection + text encoder + generators + L layers on the prepared graph.  A second figure, NOT the
    # headline: `value` above rebuilds everything every step.
    prepared_fig = None
    if world == 1 and not pre_sharded and not args.no_alt:
        prep = model.prepare_packed(edge_index, utf8, offsets, N)
        for _ in range(3):
            o3 = model.forward_prepared(x, prep)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        p0.record()
        for _ in range(args.steps):
            o3 = model.forward_prepared(x, prep)
        p1.record()
        torch.cuda.synchronize(device)
        pms = p0.elapsed_time(p1) / args.steps
        prepared_fig = {"ms_per_step": pms, "value": E * L / (pms / 1e3), "unit": "edges/s/layer",
                        "step_frac": gbs(algorithmic_bytes(w, E, N, N)[1] * L, pms) / hbm_peak,
                        "what": "HyperGNN.forward_prepared on a graph prepared once (relation dedup + graph build "
                                "outside the timed region); projection, text encoder, generators and layers inside"}
        del prep, o3

    # ---- the same step on the other engines (fp32-tolerance context of an f16 / tf32 headline)
    alt = None
    if world == 1 and not pre_sharded and not args.no_alt and precision != "fp32":
        alt = {}
        try:
            del out
        except NameError:
            pass
        for other in [q for q in ("tf32", "fp32") if q != precision and (q != "tf32" or d in (32, 64, 128))]:
            m2 = build_model(w, device, other)
            torch.cuda.synchronize(device)
            for _ in range(2):
                o2 = m2.forward_packed(x, edge_index, utf8, offsets)
            _native.profile_enable(True)
            _native.profile_read()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k = 3
            a0.record()
            for _ in range(k):
                o2 = m2.forward_packed(x, edge_index, utf8, offsets)
            a1.record()
            torch.cuda.synchronize(device)
            prof2, n2 = _native.profile_read()
            _native.profile_enable(False)
            ms2 = a0.elapsed_time(a1) / k
            tc2 = prof2["contraction_ms"] / max(n2, 1)
            tl2 = (prof2["contraction_ms"] + prof2["epilogue_ms"] + prof2["prep_ms"]) / max(n2, 1)
            alt[other] = {"ms_per_step": ms2, "value": E * L / (ms2 / 1e3), "unit": "edges/s/layer",
                          "contraction_ms": tc2, "contraction_frac": gbs(b_contr, tc2) / hbm_peak,
                          "layer_ms": tl2, "layer_frac": gbs(b_layer, tl2) / hbm_peak,
                          "step_frac": gbs(b_layer * L, ms2) / hbm_peak,
                          "tolerance": {"tf32": "upd 2e-3 of max, h 8e-4 abs (tests/_util.py)",
                                        "fp32": "rtol 1e-4 / atol 2e-5 on upd, h (fp32 end to end)"}[other]}
            del m2, o2
        alt["note"] = ("engines: f16 = fp16 operand transport + tcgen05 kind::f16, tf32 = tcgen05 kind::tf32 on fp32 rows, "
                       "fp32 = CUDA-core FFMA; all accumulate in fp32")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_subprocess(args)

    if rank == 0:
        line = {"metric": "hypergnn_fwd_edges_per_sec_per_layer", "value": value, "unit": "edges/s/layer",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": precision,
                "data": "synthetic",
                "config": make_config(w, world, args.skew),
                "roofline": roofline, "prepared_graph": prepared_fig, "precision_alt": alt, "cpu_baseline": cpu, "e2e": e2e, "multi_gpu": multi,
                "gpu_launches": launches,
                "ms_each_step": [round(v, 3) for v in each],
                "clocks": clocks.summary()}
        if train is not None:
            line["train_step"] = train
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

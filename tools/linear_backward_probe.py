"""ghf_linear_backward on the two shapes that matter at c3 (for event timing and for `ncu -k regex:linear_bwd`):
the input projection (2.5M x 128 -> 128, ReLU; only dL/dW and dL/db are needed) and a generator head
(535 x 128 -> 16384, exp(log_scale); all four gradients).        python tools/linear_backward_probe.py   GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200")]
import torch  # noqa: E402

from graph_hypernetwork_forge import _native  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for name, (M, K, N, relu, scaled, need_x) in {"input projection": (2_500_000, 128, 128, True, False, False),
                                              "generator head": (535, 128, 16384, False, True, True)}.items():
    x = torch.randn(M, K, generator=g, device=dev)
    w = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
    b = torch.randn(N, generator=g, device=dev)
    ls = torch.tensor([-1.0], device=dev) if scaled else None
    g_y = torch.randn(M, N, generator=g, device=dev)
    y = _native.linear(x, w, b, relu=relu, log_scale=ls)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.linear_backward(x, w, ls, y, g_y, relu, need_x=need_x, need_ls=scaled)
        e1.record()
        torch.cuda.synchronize()
    flop = 2.0 * M * N * K * (2 if need_x else 1)
    gbytes = 4.0 * (M * N * (2 if relu or scaled else 1) + M * K + N * K * (2 if need_x else 1)) / 1e9
    ms = e0.elapsed_time(e1)
    print(f"{name}: M={M} K={K} N={N}: {ms:.3f} ms, {flop / ms / 1e9:.1f} TFLOP/s fp32, {gbytes / ms * 1e3:.0f} GB/s of "
          f"{gbytes:.2f} GB algorithmic")

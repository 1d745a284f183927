"""Measured dense TF32 (and bf16) tensor throughput of this GPU through cuBLAS: the denominator for fractions quoted
on the tf32 engine (SURVEY 8d asks for a measured TF32 peak; MEASURED_PEAKS.json holds bf16 and HBM only).
    python tools/tf32_peak.py        GPU box only."""
import torch

torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
n = 8192
for name, dtype in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {n}^3 matmul best of 10: {best:.3f} ms = {2 * n ** 3 / best / 1e9:.1f} TFLOP/s (cuBLAS, burst)")

"""Which Philox counters does torch's CUDA dropout use for element i?  (GPU box only.)

The native epilogue applies training-mode dropout (HG:293-294) itself and must drop exactly the elements
`F.dropout` would drop under the same generator state, so that the drop-in follows the reference's random stream.
This probe restates the mapping on the host (numpy Philox4x32-10) and compares it with `F.dropout` on the device:

    element i of a contiguous fp32 tensor with numel % 4 == 0 (vectorised kernel, 4 elements per thread and step):
      T = 256 * min(SMs * (max threads per SM / 256), ceil(numel / 256))      threads of the launch
      v = i // 4;  thread t = v % T;  step s = v // T;  component c = i % 4
      r = Philox4x32-10(key = seed, counter = (offset / 4 + s, subsequence = t))[c]
      kept  <=>  fmaf(float(r), 2^-32, 2^-33) < 1 - p
    and the generator offset advances by 4 * ceil(numel / (4 * T)).
"""
import sys

import numpy as np
import torch

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox(seed, ctr_lo, subseq):
    """ctr_lo, subseq: uint64 arrays -> [4, n] uint32 outputs of Philox4x32-10."""
    c = [(ctr_lo & 0xFFFFFFFF).astype(np.uint64), (ctr_lo >> np.uint64(32)).astype(np.uint64),
         (subseq & 0xFFFFFFFF).astype(np.uint64), (subseq >> np.uint64(32)).astype(np.uint64)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for r in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c).astype(np.uint32)


def predicted_keep(numel, p, seed, offset, sms, max_threads_per_sm):
    T = 256 * min(sms * (max_threads_per_sm // 256), -(-numel // 256))
    i = np.arange(numel, dtype=np.uint64)
    v = i // np.uint64(4)
    t, s = v % np.uint64(T), v // np.uint64(T)
    out = philox(seed, np.uint64(offset // 4) + s, t)
    r = out[(i % np.uint64(4)).astype(np.int64), np.arange(numel)]
    u = (r.astype(np.float64) .astype(np.float32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(np.float32)
    keep = u < np.float32(1.0 - p)
    steps = -(-numel // (4 * T))
    return keep, 4 * steps


def main():
    dev = torch.device("cuda:0")
    props = torch.cuda.get_device_properties(dev)
    sms, mt = props.multi_processor_count, props.max_threads_per_multi_processor
    print(f"torch {torch.__version__}  SMs {sms}  max threads/SM {mt}")
    gen = torch.cuda.default_generators[0]
    ok = True
    for shape, p in [((8, 128), 0.1), ((1000, 128), 0.25), ((20000, 128), 0.1), ((777, 64), 0.5), ((33, 24), 0.3),
                     ((2_500_000 // 8, 128), 0.1)]:
        torch.manual_seed(1234 + shape[0])
        torch.rand(5, device=dev)                       # move the offset off zero
        seed, off = gen.initial_seed(), gen.get_offset()
        x = torch.ones(shape, device=dev)
        y = torch.nn.functional.dropout(x, p=p, training=True)
        off_after = gen.get_offset()
        keep, adv = predicted_keep(x.numel(), p, seed, off, sms, mt)
        got = (y.flatten() != 0).cpu().numpy()
        same = float((got == keep).mean())
        scale_ok = bool(torch.all((y == 0) | (y == torch.tensor(np.float32(1.0 / np.float64(np.float32(1.0 - p))), device=dev))))
        print(f"shape {shape} p {p}: mask agreement {same:.6f}  kept {got.mean():.4f}  offset {off} -> {off_after} "
              f"(predicted +{adv}: {'ok' if off + adv == off_after else 'MISMATCH'})  scale {'ok' if scale_ok else 'MISMATCH'}")
        ok = ok and same == 1.0 and off + adv == off_after and scale_ok
    print("PROBE", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

"""Host-side pieces of bench.py that need no GPU: the shared sampler of the two CPU legs, the `config` both arms must
print identically, the byte accounting behind the roofline figures, and the reference arm end to end on a tiny sample
(the unmodified reference when baseline/_ref is installed, the numpy port otherwise)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_sampled_workload_scales_nodes_with_edges():
    w = bench.WORKLOADS["c3"]
    n, k, src, dst, texts, feats = bench.sampled_workload(w, 60_000)
    assert k == 60_000 and n == round(w["N"] * k / w["E"]) == 9375           # same in-degree as the workload
    assert src.shape == dst.shape == (k,) and feats.shape == (n, w["F"]) and len(texts) == k
    assert int(src.max()) < n and int(dst.max()) < n and len(set(texts)) <= w["R"]
    n2, k2, *_ = bench.sampled_workload(w, 10 ** 12)                           # never more than the workload itself
    assert (n2, k2) == (w["N"], w["E"])


def test_config_is_the_same_in_both_arms_and_names_the_l2_policy():
    for name in ("c2", "c3", "c5"):
        w = bench.WORKLOADS[name]
        a, b = bench.make_config(w, 1, False), bench.make_config(dict(w), 1, False)
        assert a == b and set(a) == {"workload", "step", "l2", "parallelism"}
        assert ("flush" in a["l2"]) or ("scratch written between steps" in a["l2"])
    assert "x8" in bench.make_config(bench.WORKLOADS["c3"], 8, False)["parallelism"]


def test_byte_accounting_matches_survey_8d():
    w = bench.WORKLOADS["c3"]
    contraction, layer = bench.algorithmic_bytes(w, w["E"], w["N"], w["N"])
    assert abs(layer / 1e9 - 10.96) < 0.01 and abs(contraction / 1e9 - 9.67) < 0.01       # SURVEY 8(d) worked numbers
    c16, l16 = bench.bytes_as_read(w, w["E"], w["N"], "f16", fused=False)
    c32, l32 = bench.bytes_as_read(w, w["E"], w["N"], "tf32", fused=False)
    assert c16 < c32 and l16 < l32                                                          # fp16 rows and images
    assert bench.bytes_as_read(w, w["E"], w["N"], "f16", fused=True)[1] < l16              # no accumulator round trip


def test_reference_arm_prints_one_line_with_the_contract_keys():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3", "--steps", "1",
           "--warmup", "0", "--sample-edges", "3000", "--no-port-extra"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and len(lines) == 1, out.stderr[-800:]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "hypergnn_fwd_edges_per_sec_per_layer" and j["value"] > 0
    assert j["config"] == bench.make_config(bench.WORKLOADS["c3"], 1, False)
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"] and j["gpu_launches"] == 0

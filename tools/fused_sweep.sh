#!/bin/bash
# round-2 A/B of the fused layer kernel (mp_f16_fused.cu) at c3: super-block size, ring slots, cache hints
cd "$(dirname "$0")/.."
B="python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu"
run() { echo "== $1"; shift; env "$@" timeout 300 $B 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        j = json.loads(l); r = j["roofline"]
        print("  ms/step %.3f  contraction %.3f ms  layer %.3f ms  steps %s" % (j["ms_per_step"], r["ms_per_launch"], r["layer"]["ms"], j["ms_each_step"]))
    elif "rror" in l or "Traceback" in l: print("  " + l.rstrip())
'; }
run "unfused (round 1 path)" GHF_MP_FUSED=0
run "fused default (slots 3, sb 49152)" GHF_MP_FUSED=1
run "fused slots 2" GHF_FUSED_SLOTS=2
run "fused slots 4" GHF_FUSED_SLOTS=4
run "fused sb 32768 slots 2" GHF_SB_NODES=32768 GHF_FUSED_SLOTS=2
run "fused sb 32768 slots 3" GHF_SB_NODES=32768 GHF_FUSED_SLOTS=3
run "fused sb 32768 slots 4" GHF_SB_NODES=32768 GHF_FUSED_SLOTS=4
run "fused sb 24576 slots 3" GHF_SB_NODES=24576 GHF_FUSED_SLOTS=3
run "fused sb 24576 slots 4" GHF_SB_NODES=24576 GHF_FUSED_SLOTS=4
run "fused sb 65536 slots 2" GHF_SB_NODES=65536 GHF_FUSED_SLOTS=2
run "fused default, no ring evict_last" GHF_FUSED_FLAGS=1
run "fused default, dst rows evict_last" GHF_FUSED_FLAGS=7
run "fused default, weights evict_last" GHF_FUSED_FLAGS=13

#!/bin/bash
# diagnostics of the fused layer kernel at c3: where does the time go?
cd "$(dirname "$0")/.."
B="python bench.py --workload c3 --steps 3 --warmup 3 --no-e2e --no-cpu"
run() { echo "== $1"; shift; env "$@" timeout 300 $B 2>&1 | python -c '
import sys, json
seen = 0
for l in sys.stdin:
    if l.startswith("{"):
        j = json.loads(l); r = j["roofline"]
        print("  ms/step %.3f  contraction %.3f ms  layer %.3f ms" % (j["ms_per_step"], r["ms_per_launch"], r["layer"]["ms"]))
    elif "trace" in l:
        seen += 1
        if seen in (10, 11, 12): print("  " + l.rstrip())
    elif "rror" in l or "Traceback" in l: print("  " + l.rstrip())
'; }
run "fused default" GHF_MP_FUSED=1
run "fused, 4 producer warps" GHF_FUSED_PROD=4
run "fused, row epilogue without body (timing only)" GHF_FUSED_FLAGS=37
run "fused, no reductions (timing only)" GHF_FUSED_FLAGS=69
run "fused, neither (timing only)" GHF_FUSED_FLAGS=101
run "fused trace" GHF_FUSED_TRACE=1
run "fused trace, slots 2" GHF_FUSED_TRACE=1 GHF_FUSED_SLOTS=2

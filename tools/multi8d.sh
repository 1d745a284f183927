#!/bin/bash
# round-2 fourth 8-GPU session: c3 on 8 and 4 GPUs after the pre-sync hooks (h0 under the edge selection, generators
# under graph build)
cd "$(dirname "$0")/.."
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); m = j["multi_gpu"]
    print(sys.argv[1], "ms/step %.3f" % j["ms_per_step"], "value %.3e" % j["value"], {k: round(v, 2) for k, v in m["stage_ms_max_over_ranks"].items()}, "parity", m["parity"] and m["parity"]["max_abs_diff"],
          "e2e", j["e2e"] and round(j["e2e"]["ms_per_step"], 2), "push", m["push"], "chunks", m["chunks"], "sent", m["rows_sent_fraction"], j["ms_each_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", repr(e))
PY
}
run() { out=$1; n=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; show gpurun_out/$out.json; grep -E "Error|error:" gpurun_out/$out.err | tail -3; }
run r2d_c3_n8 8 --steps 10 --warmup 4
run r2d_c3_n4 4 --steps 10 --warmup 4 --no-e2e --no-check

#!/bin/bash
# round-2 8-GPU session: parity of the sharded forward, c3 at 8 ranks (chunk sweep), c5 at full size
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); m = j["multi_gpu"]
    print(sys.argv[1], "ms/step %.3f" % j["ms_per_step"], "value %.3e" % j["value"], m["stage_ms_max_over_ranks"], "parity", m["parity"] and m["parity"]["max_abs_diff"],
          "e2e", j["e2e"] and round(j["e2e"]["ms_per_step"], 2), "chunks", m["chunks"], "pushed", m["bytes_pushed_per_step_per_rank"])
except Exception as e:
    print(sys.argv[1], "FAILED", repr(e))
PY
}
timeout 300 $TR --master-port 29531 tools/check_sharded.py > gpurun_out/check$N.log 2>&1; echo "check exit=$?"; grep -E "world" gpurun_out/check$N.log | tail -12
timeout 300 $TR --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --chunks 2 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; show gpurun_out/bench_c3_n$N.json
for C in 1 4; do
  timeout 300 $TR --master-port 2953$((3+C)) bench.py --gpus $N --steps 5 --warmup 3 --chunks $C --no-e2e --no-check > gpurun_out/bench_c3_n${N}_c$C.json 2> gpurun_out/bench_c3_n${N}_c$C.err; show gpurun_out/bench_c3_n${N}_c$C.json
done
timeout 300 $TR --master-port 29539 bench.py --gpus $N --steps 5 --warmup 3 --transport collective --no-e2e --no-check > gpurun_out/bench_c3_n${N}_coll.json 2> gpurun_out/bench_c3_n${N}_coll.err; show gpurun_out/bench_c3_n${N}_coll.json
timeout 600 $TR --master-port 29540 bench.py --gpus $N --workload c5 --steps 4 --warmup 3 --chunks 4 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; show gpurun_out/bench_c5_n$N.json
grep -vE "^\s*$|OMP_NUM|\*\*\*|FutureWarning|enable_symm|\[nccl\]" gpurun_out/bench_c5_n$N.err | tail -8

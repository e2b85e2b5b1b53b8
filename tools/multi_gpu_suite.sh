#!/bin/bash
# Multi-GPU evidence on one box:  gpurun --gpus N --timeout 1800 -- bash tools/multi_gpu_suite.sh N [tag]
# Runs the multi-GPU tests, then bench.py under torchrun for the configs that name N GPUs; JSON lines land in
# gpurun_out/<tag>_n<N>_<config>_<scaling>.json
N=${1:-2}
TAG=${2:-r02}
WHAT=${3:-"tests c3:weak c3:strong c4:strong c5:weak c5:strong"}
PORT=29517
run() {  # config scaling extra-args...
  local c=$1 s=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N \
    --config $c --scaling $s --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_n${N}_${c}_${s}.json 2> gpurun_out/${TAG}_n${N}_${c}_${s}.err
  PORT=$((PORT+1))
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_n${N}_${c}_${s}.json"))
    print("${c} ${s} n=${N}: value %.4g %s, ms/step %.3f, e2e %.4g, kernel_ms %s, pairs/s %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("kernel_ms"), d["config"].get("pairs_per_s")))
except Exception as e:
    print("${c} ${s} FAILED", e); print(open("gpurun_out/${TAG}_n${N}_${c}_${s}.err").read()[-1500:])
PY
}
nvidia-smi -L | head -8
for w in $WHAT; do
  if [ $w = tests ]; then python -m pytest tests -m gpu -x -q -k "multi_gpu or allgather" 2>&1 | tail -5; else run ${w%%:*} ${w##*:}; fi
done

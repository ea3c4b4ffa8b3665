#!/bin/bash
# usage: run_scaling.sh N nx ny nz label [extra bench args]   (on the GPU box, N GPUs)
N=$1; shift; NX=$1; NY=$2; NZ=$3; LABEL=$4; shift 4
OUT=gpurun_out/bench_r02_${LABEL}_${N}gpu.json
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 30 --warmup 5 --mesh $NX $NY $NZ --no-cpu "$@" > $OUT 2> ${OUT%.json}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 30 --warmup 5 --mesh $NX $NY $NZ "$@" > $OUT 2> ${OUT%.json}.err
fi
python - <<PY
import json
try:
    d = json.load(open("$OUT"))
    print("$LABEL", d["n_gpus"], "GPUs", d["config"]["workload"][:40], "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["config"]["schedule"], "build %.1fs" % d["config"]["host_operator_build_s"], d.get("slab_balance"), d["parity_check"]["equal"])
except Exception as e:
    print("FAILED", e); print(open("${OUT%.json}.err").read()[-1500:])
PY

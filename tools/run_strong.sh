#!/bin/bash
# strong scaling of the headline mesh on N GPUs of this box: tools/run_strong.sh N [N ...]
for N in "$@"; do
  OUT=gpurun_out/bench_r02_final_${N}gpu.json
  if [ "$N" = "1" ]; then python bench.py --steps 20 --warmup 5 --no-cpu > $OUT 2> ${OUT%.json}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > $OUT 2> ${OUT%.json}.err; fi
  python - <<PY
import json
try:
    d = json.load(open("$OUT"))
    print(d["n_gpus"], "GPUs value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "build %.1fs" % d["config"]["host_operator_build_s"], d.get("slab_balance"), d["parity_check"]["equal"], d["parity_check"]["full_size_digest"]["E"], d["clocks"])
except Exception as e:
    print("FAILED", e); print(open("${OUT%.json}.err").read()[-1500:])
PY
done

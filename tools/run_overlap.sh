#!/bin/bash
# experiment: halo work on the side stream (option overlap_halo) at N GPUs
N=$1
for O in ${2:-1 0}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$O bench.py --gpus $N --steps 40 --warmup 5 --opt overlap_halo=$O > gpurun_out/ov_${N}_$O.json 2> gpurun_out/ov_${N}_$O.err
  python -c "
import json
d=json.load(open('gpurun_out/ov_${N}_$O.json')); print('N=$N overlap_halo=$O: %.0f MCells/s  %.4f ms/step' % (d['value'], d['ms_per_step']), d['slab_balance']['busy_ms_per_rank'] if d.get('slab_balance') else '')"
done

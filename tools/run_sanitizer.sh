#!/bin/bash
# compute-sanitizer record of the round-2 kernels (GPU box); writes gpurun_out/sanitizer_r02.txt
OUT=gpurun_out/sanitizer_r02.txt
mkdir -p gpurun_out
echo "compute-sanitizer runs on a B200 (round 2, after the last kernel change)" > $OUT
run() { # tool, description, command...
  TOOL=$1; shift
  echo "" >> $OUT
  echo "$TOOL: $*" >> $OUT
  compute-sanitizer --tool $TOOL $EXTRA "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|ok " | tail -8 >> $OUT
}
run memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "x_slab_windows or odd_sizes or programmatic or graded"
run memcheck python -m pytest tests/test_gpu_readout_ext.py tests/test_gpu_slab_readout.py tests/test_gpu_multi.py -q -m gpu -k "lorentz or dispersive or slabs or one_gpu"
EXTRA="--racecheck-report all" run racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "x_slab_windows and n0"
EXTRA="--racecheck-report all" run racecheck python -m pytest tests/test_gpu_readout_ext.py -q -m gpu -k "one_pass_schedule and c4"
cat $OUT

python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "x_slab_windows" 2>&1 | tail -3
for v in "" s3b6 s2b6 s2b8; do
  if [ -n "$v" ]; then export OPENEMS_B200_LIB=/root/repo/openems_b200/lib/variants/lib_$v.so; fi
  echo "variant: $v"
  python tools/xslab_time.py 2>&1 | grep -E "xslab=2" | head -3
done

timeout 600 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/sr_tests.txt
for r in new base new base; do
  if [ $r = new ]; then so=""; else so="/root/repo/stateline_b200/libslgpu_targets_base.so"; fi
  echo "== $r"
  SLGPU_TARGETS_SO=$so python tools/probe.py --workload config3 --warm 200 --rounds 1000 2>&1 | grep us_per_round | tail -2 | cut -c60-200
done > gpurun_out/sr_probe.txt 2>&1

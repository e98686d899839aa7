for tool in memcheck racecheck; do
  for w in "config3 --stacks 64 --warm 12 --rounds 12" "config4 --stacks 8 --warm 11 --rounds 3" "config1 --warm 12 --rounds 12" "config5 --stacks 4 --warm 11 --rounds 3"; do
    echo "== $tool $w"
    timeout 300 compute-sanitizer --tool $tool --print-limit 5 python tools/probe.py --workload $w --sink host 2>&1 | grep -v "^{" | grep "ERROR SUMMARY\|RACECHECK SUMMARY\|Invalid\|hazard\|rhat" | head -8
  done
done > gpurun_out/sanitizer.txt 2>&1

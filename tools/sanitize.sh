#!/bin/bash
# compute-sanitizer over small slices of BASELINE configs #1-#3 (one tool per gpurun call, as the profiling recipe asks):
#   gpurun -- 'bash tools/sanitize.sh memcheck'     (or racecheck / synccheck / initcheck)
# Writes gpurun_out/sanitize_<tool>.log; the summaries kept under profiles/ come from there.
tool=${1:-memcheck}
out=gpurun_out/sanitize_${tool}.log
: > "$out"
for wl in config1 config2:16 config3:24 config4:2; do
  name=${wl%%:*}; S=${wl#*:}; [ "$S" = "$wl" ] && S=0
  echo "== $tool $name S=$S" >> "$out"
  timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 python - "$name" "$S" >> "$out" 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
from stateline_b200 import Engine, configs
name, S = sys.argv[1], int(sys.argv[2])
w = configs.workload(name, S=S or None)
with Engine(configs.engine_config(w, sink="host", seed=3, hostRingRounds=64), plugin=w["plugin"], payload=w["payload"]) as e:
    e.init_chains()
    e.step_rounds(23)
    rows = e.drain()
    print("rows", rows.shape, "launches", e.launch_count)
PY
  echo "rc=$?" >> "$out"
done
grep -E "^== |ERROR SUMMARY|rc=|rows" "$out"

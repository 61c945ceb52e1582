#!/bin/bash
# GPU round trip for the big grids: tests, C2 bench, launch lists of one steady-state build at C4 / C5 size.
# usage: tools/gpu_cycle_big.sh <tag> [notest]
tag=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/test_$tag.txt
  tail -5 gpurun_out/test_$tag.txt
fi
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["stage_ms"], "launches", d.get("gpu_launches_per_step"), "e2e_cli", d.get("e2e_cli"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-2000:])
PY
for c in c4 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${c}$tag.csv python tools/profile_big.py $c > gpurun_out/big_${c}$tag.txt 2>&1
  grep "^$c" gpurun_out/big_${c}$tag.txt
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_${c}$tag.csv")) if len(r)>5 and r[0].isdigit()]
half=rows[len(rows)//2:]
for r in half: print("  %-60s %10s" % (r[4][:60], r[-1]))
PY
done

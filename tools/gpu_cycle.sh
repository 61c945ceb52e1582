#!/bin/bash
# One GPU round trip of the development loop: tests, bench, launch list, full ncu capture of the top kernels.
# usage: tools/gpu_cycle.sh <tag> [notest]
tag=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/test_$tag.txt
  tail -5 gpurun_out/test_$tag.txt
fi
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["stage_ms"], "launches", d.get("gpu_launches_per_step"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-2000:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py c2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vox_warp|k_emit_leaf|k_brick|k_dense_scan|k_emit_upper" -s 10 -c 14 -o gpurun_out/prof_$tag -f python tools/profile_step.py c2 > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep

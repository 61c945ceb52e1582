timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.json 2>/dev/null
timeout 600 python tests/scale_parity.py --golden c4 c5 2>&1 | tail -3
cp gpurun_out/scale_parity.json gpurun_out/scale_parity_golden_1gpu.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or steady or second or speculative" 2>&1 | tail -8 > gpurun_out/sanitizer_memcheck.txt; tail -4 gpurun_out/sanitizer_memcheck.txt

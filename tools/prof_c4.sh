timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit_leaf|k_brick|k_emit_upper_fast" -s 6 -c 6 -o gpurun_out/prof_c4l -f python tools/profile_big.py c4 > gpurun_out/ncu_c4l.log 2>&1
ls -la gpurun_out/prof_c4l.ncu-rep; tail -3 gpurun_out/ncu_c4l.log

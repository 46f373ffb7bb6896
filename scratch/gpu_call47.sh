mkdir -p /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_kernel" -c 1 -o /tmp/ncu/fwd python scratch/bench_attn.py > gpurun_out/ncu_fwd.log 2>&1; tail -2 gpurun_out/ncu_fwd.log
ncu -i /tmp/ncu/fwd.ncu-rep --page raw --csv > gpurun_out/r2_ncu_attn_fwd_raw.csv 2>/dev/null
ncu -i /tmp/ncu/fwd.ncu-rep --page source --csv > gpurun_out/r2_ncu_attn_fwd_source.csv 2>/dev/null
ls -la gpurun_out/r2_ncu_attn_fwd_raw.csv gpurun_out/r2_ncu_attn_fwd_source.csv

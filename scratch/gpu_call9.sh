timeout 300 python -m pytest tests/test_step_gpu.py -q -k cuda_graph -s 2>&1 | grep -v "LoRA target" | tail -4
timeout 200 python scratch/bench_graph.py tiny 2>&1 | grep -v "LoRA target" | tee gpurun_out/graph_tiny.txt
timeout 300 python scratch/bench_graph.py 13.6b 2>&1 | grep -v "LoRA target" | tee gpurun_out/graph_13b.txt
timeout 200 python scratch/bench_elem.py 2>&1 | tee gpurun_out/elem.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_mod|qk_norm|gate_mul" -c 12 -o gpurun_out/r2_elem python scratch/bench_elem.py > gpurun_out/ncu_elem.log 2>&1; tail -2 gpurun_out/ncu_elem.log
B200TTA_ATTN_BWD=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_kernel|dq_kernel|dkv_kernel" -c 3 -o gpurun_out/r2_attn python scratch/bench_attn.py > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
B200TTA_ATTN_BWD=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dkvq_kernel" -c 1 -o gpurun_out/r2_attn_fused python scratch/bench_attn.py > gpurun_out/ncu_attn_fused.log 2>&1; tail -2 gpurun_out/ncu_attn_fused.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/r2_launches.csv
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck.log

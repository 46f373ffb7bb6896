timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "colsum or sgd or adamw or sumsq or clip" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_tta_gpu.py -q -x 2>&1 | tail -3
mkdir -p /tmp/ncu
B200TTA_ATTN_BWD=split timeout 600 ncu --set full --clock-control none -k regex:"dq_kernel|dkv_kernel" -c 2 -o /tmp/ncu/attn_bwd python scratch/bench_attn.py > /tmp/ncu/a.log 2>&1; tail -1 /tmp/ncu/a.log
timeout 600 ncu --set full --clock-control none -k regex:"ln_mod|qk_norm" -c 26 -o /tmp/ncu/elem_after python scratch/bench_elem.py > /tmp/ncu/b.log 2>&1; tail -1 /tmp/ncu/b.log
timeout 600 ncu --set full --clock-control none -k regex:"gemm2_kernel" -c 1 -o /tmp/ncu/gemm2 python scratch/bench_gemm.py > /tmp/ncu/c.log 2>&1; tail -1 /tmp/ncu/c.log
for f in attn_bwd elem_after gemm2; do ncu -i /tmp/ncu/$f.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$f.csv 2>/dev/null; ls -la gpurun_out/r2_ncu_$f.csv; done

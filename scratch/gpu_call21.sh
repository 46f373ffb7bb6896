for g in 8 16 32 4; do
  B200TTA_GEMM_GROUP_M2=$g timeout 120 python scratch/bench_gemm.py 2>&1 | sed "s/^/g=$g /"
done

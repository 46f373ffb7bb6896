for i in 1 2; do
timeout 120 python scratch/bench_attn.py 2>&1 | head -1
B200TTA_LIB=longcat_video_tta_b200/libb200tta_nostats.so timeout 120 python scratch/bench_attn.py 2>&1 | head -1 | sed 's/^/nostats /'
done

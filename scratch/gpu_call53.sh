for v in "" _st16 "" _st16; do
  echo "== lib$v"
  B200TTA_LIB=longcat_video_tta_b200/libb200tta$v.so timeout 200 python scratch/bench_attn.py 2>&1 | grep "^attn_fwd\|spot-check" | head -2
done

for v in "" _sleep150 _sleep300; do
  echo "== lib$v"
  B200TTA_LIB=longcat_video_tta_b200/libb200tta$v.so timeout 200 python scratch/bench_attn.py 2>&1 | grep "^attn_fwd"
done

for lib in debug nostats; do
  echo "== $lib"
  B200TTA_LIB=longcat_video_tta_b200/libb200tta_$lib.so B200TTA_ATTN_BWD=split timeout 200 python scratch/bench_attn.py 2>&1 | grep -v "^fwd\|spot" | tail -6
done

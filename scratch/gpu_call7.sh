B200TTA_ATTN_BWD=split timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err; tail -2 gpurun_out/bench_split.err
B200TTA_ATTN_BWD=fused timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -2 gpurun_out/bench_fused.err
python - <<'PY'
import json
for n in ("split","fused"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d["ms_per_step"], d["clocks"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernel_ms_per_step"].items() if "attn" in k})
    except Exception as e: print(n, "failed", e)
PY

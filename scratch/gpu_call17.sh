timeout 600 python -m pytest tests/test_bsa_gpu.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 2 --warmup 3 --lat-h 96 --lat-w 160 --bsa-sparsity 0.9375 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_720p_bsa_n1.json 2> gpurun_out/bench_720p_bsa_n1.err || tail gpurun_out/bench_720p_bsa_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_720p_bsa_n1.json")); print(d["ms_per_step"], d["clocks"]); s=0
for k,v in d["kernel_ms_per_step"].items():
    s+=v["ms"]; print("  ",k,v["ms"],v["n"],v.get("tflops"))
print("sum of ABI families", s)
PY

timeout 900 python bench.py --method full --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err || tail -20 gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_full.json"))
    print("full", d["ms_per_step"], d["e2e"], d["config"]["adapter_params"], d["loss_first_last"], d["achieved_tflops_per_gpu"], d["config"]["recompute"][-80:])
    for k,v in d["kernel_ms_per_step"].items(): print("  ",k,v["ms"],v["n"],v.get("tflops"))
except Exception as e: print("failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv | tail -1

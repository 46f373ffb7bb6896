timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default_final2.json 2> gpurun_out/bench_default_final2.err; tail -2 gpurun_out/bench_default_final2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default_final2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["roofline"], d["library_baseline"]["ms_per_step"], d["gpu_launches"])
for k,v in list(d["kernel_ms_per_step"].items())[:8]: print("  ",k,v["ms"],v["n"],v.get("tflops"))
PY

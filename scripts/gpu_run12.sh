mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -2 gpurun_out/bench_r1f.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1f.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"cmp",d["cmp"]["value"],"cmp e2e",d["cmp"]["e2e"]["value"], d["phases_ms_per_step"], d["gpu_launches"], d["clocks"])
PY

mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; tail -2 gpurun_out/bench_last.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_last.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"cmp",d["cmp"]["value"],"cmp e2e",d["cmp"]["e2e"]["value"], d["gpu_launches"], d["roofline"]["traffic"])
PY

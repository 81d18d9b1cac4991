mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err
tail -3 gpurun_out/bench_n8b.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n8b.json") if l.startswith("{")][-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"cmp",d["cmp"]["value"],"cmp e2e",d["cmp"]["e2e"], d["phases_ms_per_step"], d["clocks"])
PY

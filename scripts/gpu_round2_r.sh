#!/bin/bash
# eight GPUs, final state: bench at N = 8 (sharded compare inside the library, per-rank host threads)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2r_gpus.txt; nproc >> gpurun_out/r2r_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 3 --no-cli > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err
tail -3 gpurun_out/r2r_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2r_bench_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=8 value %.1f e2e %.1f packed %.1f cmp %.2f e2ecmp %.2f host_threads %s f %.2f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["packed"]["value"] / 1e9, d["cmp"]["value"] / 1e9, d["cmp"]["e2e"]["value"] / 1e9, d["e2e"]["host_threads"], d["e2e"]["host_packed_fraction"]))
print(d["phases_ms_per_step"], d["cmp"]["roofline"].get("code_prep_ms_per_step"), d["verify"])
PY

#!/bin/bash
# two GPUs: bench at N = 2 with the sharded host compare in the e2e leg; --gpu-stats test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "gpu_stats or gpus_option" 2>&1 | tail -4 > gpurun_out/r2w_pytest_cli.txt; tail -2 gpurun_out/r2w_pytest_cli.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cli > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err
tail -3 gpurun_out/r2w_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2w_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=2 value %.1f e2e %.1f packed %.1f cmp %.2f e2ecmp %.2f host_threads %s f %.2f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["packed"]["value"] / 1e9, d["cmp"]["value"] / 1e9, d["cmp"]["e2e"]["value"] / 1e9, d["e2e"]["host_threads"], d["e2e"]["host_packed_fraction"]), d["verify"])
PY

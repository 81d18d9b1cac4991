#!/bin/bash
# A/B of the fast-kernel key + barrier change: 2048-genome kernel timing, windowed parity tests, headline bench
mkdir -p gpurun_out
timeout 600 python scripts/sketch_only_bench.py 2048 5000000 4 fss 4096 51 > gpurun_out/r2n_sketch_only.txt 2>&1; tail -4 gpurun_out/r2n_sketch_only.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_scale.py -q --tb=short -rf -k "window or fast or golden or config2 or tie or redo or keymask" 2>&1 | tail -6 > gpurun_out/r2n_pytest.txt; tail -3 gpurun_out/r2n_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cli > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fast_kernel -c 1 -o gpurun_out/r2n_fast python scripts/sketch_only_bench.py 2048 5000000 1 fss 4096 51 > gpurun_out/r2n_ncu.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f packed %.1f launch_ms %.2f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["packed"]["value"] / 1e9, d["roofline"]["launch_ms"]), d["verify"])
PY

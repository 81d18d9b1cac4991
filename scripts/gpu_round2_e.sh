#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "sharded or baseline_scale or config or packed_entry or chunked" 2>&1 | tail -15 > gpurun_out/r2e_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cli > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
D2G_HYBRID_F=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cli --no-cpu-baseline --no-verify > gpurun_out/r2e_bench_f1.json 2>> gpurun_out/r2e_bench.err
tail -6 gpurun_out/r2e_pytest.txt; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2e_bench.json", "gpurun_out/r2e_bench_f1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f e2e %.1f packed %.1f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["packed"]["value"] / 1e9))
    except Exception as e:
        print(f, e)
PY

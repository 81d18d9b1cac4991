#!/bin/bash
# final state, one GPU: whole suite, headline bench, launch list with DRAM traffic
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -15 > gpurun_out/r2o_pytest.txt; tail -3 gpurun_out/r2o_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2o_bench.err
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cli --no-verify --e2e-genomes 64 > gpurun_out/r2o_bench_under_ncu.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f packed %.1f launch_ms %.2f cli %.2f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["packed"]["value"] / 1e9, d["roofline"]["launch_ms"], d.get("e2e_cli", {}).get("speedup", 0)), d["verify"], d["e2e"]["host_threads"], d["e2e"]["host_packed_fraction"])
PY

#!/bin/bash
# final tree: whole GPU suite, smoke, config 3 line, headline bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -8 > gpurun_out/r2u_pytest.txt; tail -3 gpurun_out/r2u_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.txt 2>&1; tail -1 gpurun_out/r2u_smoke.txt
timeout 1500 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2u_bench_c3.json 2> gpurun_out/r2u_bench_c3.err; echo "c3 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r2u_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_bench_c3.json").read().strip().splitlines()[-1])
print("c3 value %.2f G kmers/s ms %.0f" % (d["value"] / 1e9, d["ms_per_step"]), d["roofline"]["phases_ms_per_step"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
d = json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f packed %.1f launch_ms %.2f traffic %s cli %.2f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["packed"]["value"] / 1e9, d["roofline"]["launch_ms"], d["roofline"]["traffic"], d.get("e2e_cli", {}).get("speedup", 0)), d["verify"])
PY

mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_final.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:sketch_kernel|cmp|c16|fss_|fill_|RadixSort|densify' -c 120 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-genomes 64 > gpurun_out/bench_under_ncu.log 2>&1
name=prof_sk_main2k_f
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -c 1 -o /tmp/$name python scripts/sketch_only_bench.py 2048 5000000 1 fss > gpurun_out/ncu_sk.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/$name.details.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"cmp",d["cmp"]["value"],"cmp e2e",d["cmp"]["e2e"]["value"], d["phases_ms_per_step"], d["gpu_launches"], d["clocks"])
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cmp"]["value"], r["cpu_baseline"]["cores"])
PY

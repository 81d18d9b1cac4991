#!/bin/bash
# contain, --nLSH 4..9, filter set through the CLI; config 3 line with the reference baseline; headline bench + reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -rf -k "contain or nlsh or topk or filterset or graph or threshold" 2>&1 | tail -30 > gpurun_out/r2m_pytest.txt
tail -6 gpurun_out/r2m_pytest.txt
timeout 1500 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err; echo "c3 rc=$?"; tail -c 300 gpurun_out/r2m_bench_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2m_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.txt 2>&1; tail -2 gpurun_out/r2m_smoke.txt
for f in r2m_bench_c3 r2m_bench_ref r2m_bench; do head -c 600 gpurun_out/$f.json; echo; done

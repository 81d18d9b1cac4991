python -m pytest tests -m gpu -x -q -k "compare or cmp or compress" 2>&1 | tail -4
python scripts/cmp_only_bench.py 10000 4096 3 codes 2>&1 | tail -2
D2G_C16_NO_HASH=1 python scripts/cmp_only_bench.py 10000 4096 2 codes 2>&1 | tail -1
python scripts/cmp_only_bench.py 16000 1024 2 codes 2>&1 | tail -1
python scripts/cmp_only_bench.py 3000 4096 2 auto 2>&1 | tail -1
D2G_CMP_PATH=f64 python scripts/cmp_only_bench.py 3000 4096 2 auto 2>&1 | tail -1
python scripts/cmp_only_bench.py 1000 1024 2 codes 2>&1 | tail -1
python scripts/cmp_only_bench.py 1000 1024 2 f64 2>&1 | tail -1

python -m pytest tests -m gpu -x -q -k "compare or cmp or compress" 2>&1 | tail -3
python scripts/cmp_only_bench.py 10000 4096 2 codes 2>&1 | tail -1
python scripts/cmp_only_bench.py 20000 4096 2 codes 2>&1 | tail -1
D2G_C16_NO_HASH=1 python scripts/cmp_only_bench.py 20000 4096 2 codes 2>&1 | tail -1
python scripts/cmp_only_bench.py 40000 1024 2 codes 2>&1 | tail -1
D2G_C16_NO_HASH=1 python scripts/cmp_only_bench.py 40000 1024 2 codes 2>&1 | tail -1
python scripts/config_bench.py c4 50000 100000 1024 2>&1 | tail -1

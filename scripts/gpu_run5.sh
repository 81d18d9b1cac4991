python scripts/sketch_only_bench.py 2048 5000000 2 fss 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 5000000 2 opmh 4096 51 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 5000000 2 opmh 1024 0 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 5000000 2 opmh 1024 51 2>&1 | tail -1

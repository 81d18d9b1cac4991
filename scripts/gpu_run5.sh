python -m pytest tests -m gpu -x -q -k "sketch or smoke or fss" 2>&1 | tail -3
D2G_DEBUG=1 python scripts/sketch_only_bench.py 2048 5000000 2 fss 4096 200 2>&1 | tail -2
D2G_DEBUG=1 python scripts/sketch_only_bench.py 2048 5000000 2 fss 4096 100 2>&1 | tail -2
python scripts/sketch_only_bench.py 2048 5000000 2 fss 8192 51 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 1000000 2 fss 4096 51 2>&1 | tail -1

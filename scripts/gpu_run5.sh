python -m pytest tests -m gpu -x -q -k "sketch or smoke or fss or cli" 2>&1 | tail -3
python scripts/sketch_only_bench.py 2048 5000000 2 fss 2>&1 | tail -1
python scripts/sketch_only_bench.py 512 5000000 2 fss 4096 0 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 5000000 2 fss 8192 51 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 5000000 2 fss 1024 51 2>&1 | tail -1

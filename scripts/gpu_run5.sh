python -m pytest tests -m gpu -x -q -k "sketch or smoke or cli or weighted or merge" 2>&1 | tail -3
python scripts/sketch_only_bench.py 2048 5000000 2 fss 2>&1 | tail -1
python scripts/sketch_only_bench.py 2048 5000000 2 opmh 4096 51 2>&1 | tail -1

python -m pytest tests -m gpu -x -q -k "weighted or multiset or pmh or bmh or golden" 2>&1 | tail -3
python scripts/config_bench.py c3 8 20000000 8192 pmh 2>&1 | tail -1
python scripts/config_bench.py c3 8 20000000 8192 bmh 2>&1 | tail -1
python scripts/config_bench.py c3 48 20000000 8192 pmh 2>&1 | tail -1

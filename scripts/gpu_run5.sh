D2G_DEBUG=1 python -m pytest tests -m gpu -x -q -s -k "fss_guessed" 2>&1 | grep -v "^$" | tail -12
python -m pytest tests -m gpu -x -q -k "sketch or smoke or cli or weighted or merge" 2>&1 | tail -3
D2G_DEBUG=1 python scripts/sketch_only_bench.py 2048 5000000 2 fss 2>&1 | tail -2
D2G_FSS_NO_GUESS=1 python scripts/sketch_only_bench.py 2048 5000000 2 fss 2>&1 | tail -1
python scripts/sketch_only_bench.py 512 5000000 2 fss 4096 0 2>&1 | tail -1

mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sketch or smoke or cli or weighted or merge" 2>&1 | tail -3
for st in 0 8 16 32 64; do
  echo "boot stride $st"; if [ $st = 0 ]; then unset D2G_FSS_BOOT_STRIDE; else export D2G_FSS_BOOT_STRIDE=$st; fi
  python scripts/sketch_only_bench.py 2048 5000000 2 fss 2>&1 | tail -1
done
unset D2G_FSS_BOOT_STRIDE
python scripts/sketch_only_bench.py 2048 5000000 2 opmh 4096 51 2>&1 | tail -1

mkdir -p gpurun_out
L=gpurun_out/cmpbench2.log; : > $L
for acc in 0 1 2; do
  echo "== ACC=$acc gt/lt" >> $L
  D2G_C16_ACC=$acc D2G_C16_NO_NE=1 python scripts/cmp_only_bench.py 10000 4096 3 codes 2>&1 | tail -1 >> $L
  echo "== ACC=$acc ne-only" >> $L
  D2G_C16_ACC=$acc python scripts/cmp_only_bench.py 10000 4096 3 codes 2>&1 | tail -1 >> $L
done
D2G_C16_ACC=1 python scripts/cmp_only_bench.py 10000 4096 2 codes containment 2>&1 | tail -1 >> $L
D2G_C16_ACC=1 python scripts/cmp_only_bench.py 20000 1024 3 codes 2>&1 | tail -1 >> $L
D2G_C16_ACC=1 python scripts/cmp_only_bench.py 20000 1024 3 codes poisson_llr 2>&1 | tail -1 >> $L
cat $L
python -m pytest tests -m gpu -x -q -k "compare or topk" 2>&1 | tail -3

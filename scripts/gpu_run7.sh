mkdir -p gpurun_out; free -g | head -2; nproc
L=gpurun_out/configs.log; : > $L
timeout 300 python scripts/config_bench.py c4 50000 100000 1024 >> $L 2>&1; tail -2 $L
timeout 300 python scripts/config_bench.py c5 250000 1024 32 >> $L 2>&1; tail -3 $L
timeout 300 python scripts/config_bench.py c3 8 20000000 8192 pmh >> $L 2>&1; tail -3 $L
timeout 400 python scripts/config_bench.py c3 8 20000000 8192 bmh >> $L 2>&1; tail -3 $L

mkdir -p gpurun_out
name=prof_sk_main2k_c
ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 1 -c 1 -o /tmp/$name python scripts/sketch_only_bench.py 2048 5000000 1 fss > gpurun_out/ncu_sk.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.sass.csv 2>/dev/null
tail -2 gpurun_out/ncu_sk.log

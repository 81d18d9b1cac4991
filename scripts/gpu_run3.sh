mkdir -p gpurun_out
prof() { # name, env...
  name=$1; shift
  env "$@" ncu --set full --clock-control none --import-source on -k regex:cmp16_tile -c 1 -o /tmp/$name python scripts/cmp_only_bench.py 4000 4096 1 codes > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.source.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/$name.details.txt
}
prof prof_cmp16_m0a1 D2G_C16_ACC=1 D2G_C16_NO_NE=1
prof prof_cmp16_m1a1 D2G_C16_ACC=1
ls -la gpurun_out

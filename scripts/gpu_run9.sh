mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "topk" 2>&1 | tail -3
python scripts/config_bench.py c5 250000 1024 32 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:lsh_|RadixSort|cmp' -c 200 --csv --log-file gpurun_out/launches_topk.csv python scripts/config_bench.py c5 250000 1024 32 > gpurun_out/topk_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_topk.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(",",""))/1e6
for k,(n,ms) in agg.items(): print(f"{n:4d} {ms:9.2f} ms  {k}")
PY

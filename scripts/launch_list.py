#!/usr/bin/env python
"""ncu launch list (csv of gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> the text table committed under
profiles/ (kernel, grid, ms, DRAM MB, share of the library kernels' time) and profiles/roofline_traffic.json (DRAM bytes per launch of the two
dominant kernels of BASELINE configs[1], which bench.py reports as roofline.traffic).
usage: launch_list.py <ncu.csv> <out.txt> [--traffic-json profiles/roofline_traffic.json --genomes 10000 --genome-len 5000000 --sketchsize 4096]"""
import argparse, collections, csv, json, sys
ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("out")
ap.add_argument("--traffic-json"); ap.add_argument("--genomes", type=int, default=10000); ap.add_argument("--genome-len", type=int, default=5000000)
ap.add_argument("--sketchsize", type=int, default=4096); ap.add_argument("--title", default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.csv, errors="replace")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0].isdigit():
        continue
    d = launches.setdefault(int(r[0]), {"kernel": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]]})
    v = float(r[ci["Metric Value"]].replace(",", "")); unit = r[ci["Metric Unit"]]
    name = r[ci["Metric Name"]]
    if name == "gpu__time_duration.sum":
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    else:
        mb = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        d["rd" if "read" in name else "wr"] = mb
lib = [d for d in launches.values() if "d2g::" in d["kernel"] or "cub::" in d["kernel"] or "fill_u64" in d["kernel"] or "kernel(" in d["kernel"] and "at::" not in d["kernel"]]
tot = sum(d.get("ms", 0) for d in lib)
with open(a.out, "w") as f:
    f.write(f"# {a.title}\n# id  kernel  grid  ms  dram_read_MB  dram_write_MB  share of the listed kernels' time\n")
    for i, d in enumerate(lib):
        f.write(f"{i} {d['kernel'][:110].replace(' ', '')} {d['grid'].replace(' ', '')} {d.get('ms', 0):.3f} {d.get('rd', 0):.1f} {d.get('wr', 0):.1f} {100 * d.get('ms', 0) / tot:.1f}%\n")
    f.write(f"# total {tot:.3f} ms\n")
if a.traffic_json:
    def biggest(pat):
        c = [d for d in lib if pat in d["kernel"]]
        return max(c, key=lambda d: d.get("ms", 0)) if c else None
    sk = biggest("sketch_fast_kernel") or biggest("sketch_kernel"); cm = biggest("cmp16_tile_kernel")
    json.dump({"_source": f"{a.out} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, full BASELINE configs[1] on one B200)",
               "config": {"genomes_per_gpu": a.genomes, "genome_len": a.genome_len, "sketchsize": a.sketchsize, "n_gpus": 1},
               "sketch_main_bytes_per_launch": int((sk.get("rd", 0) + sk.get("wr", 0)) * 1e6) if sk else None,
               "cmp_tile_bytes_per_launch": int((cm.get("rd", 0) + cm.get("wr", 0)) * 1e6) if cm else None,
               "sketch_main_kernel": sk["kernel"][:80] if sk else None}, open(a.traffic_json, "w"), indent=1)

#!/usr/bin/env python
"""bench.py -- the dashing2 hot paths on B200: k-mers hashed/s (sketch) + pairwise compares/s (cmp).

Workload = BASELINE.json configs[1]: 10 000 synthetic genomes x 5 Mbp, k=31 w=51, Full SetSketch
S=4096, then all-pairs symmetric comparison of the 10 000 sketches.  One STEP is one pass of BOTH hot
paths over the whole workload of this rank, inputs resident in HBM:
    sketch : all genomes (50 G k-mer positions) -> f64[10000][4096] registers + cardinalities
    cmp    : 49 995 000 pairs of those registers -> float32 condensed matrix
`value` is the sketch throughput (k-mers hashed/s, whole job); the cmp throughput and its own roofline,
e2e and cpu_baseline ride along under "cmp".  Weak scaling: every rank owns the same number of genomes;
for cmp the ranks all-gather their registers (NCCL over NVLink, the path's one exchange step) and each
computes an equal-area block of rows of the (N*10000)^2 triangle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
Under torchrun (N > 1) one rank per GPU; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, W, S = 31, 51, 4096
METRIC = "k-mers hashed/s (sketch) + pairwise compares/s (cmp)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE config to run, 1-based: 2 = configs[1], the headline (default); 1, 3, 4, 5 = the other configs' own lines (bench_configs.py, one GPU)")
    ap.add_argument("--genomes", type=int, default=None, help="genomes per rank (config 2: 10000; config 1: 64; config 3: 2000)")
    ap.add_argument("--genome-len", type=int, default=None, help="bases per genome (config 2: 5 000 000; config 1: 1 000 000; config 3: 20 000 000)")
    ap.add_argument("--n", type=int, default=None, help="config 4: reference sketches (queries = 2n; default 50 000); config 5: sketches (default 1 000 000)")
    ap.add_argument("--batch-genomes", type=int, default=64, help="config 3: genomes per library call (a batch of the counting sketches is sorted at once)")
    ap.add_argument("--e2e-genomes", type=int, default=1024, help="genomes per host-buffer call in the e2e leg")
    ap.add_argument("--cpu-genomes", type=int, default=64, help="genomes in the CPU-baseline sketch sample (at least 4 per host thread are used)")
    ap.add_argument("--cpu-cmp-n", type=int, default=5200, help="sketches in the CPU-baseline cmp sample (13.5 M pairs: seconds of reference time)")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of one genome's registers and 1000 sampled pairs")
    ap.add_argument("--no-cli", action="store_true", help="skip the CLI-vs-CLI leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.config == 2:
        args.genomes = args.genomes or 10000
        args.genome_len = args.genome_len or 5_000_000
    return args


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own OpenMP path (oracle/_ref binary) on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class RefSample:
    """A bounded sample of the bench workload on disk (tmpfs), timed with the unmodified reference binary: `dashing2 sketch` over
    n_genomes FASTA files (file-parallel OpenMP loop, so at least 4 files per thread keep every thread busy) and `dashing2 cmp` over
    cmp_n sketches (large enough to run for seconds).  Every timing returns (wall s, CPU s of the child): CPU / wall = threads busy."""

    def __init__(self, n_genomes, genome_len, cmp_n, threads):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refbin
        from dashing2_b200 import synth
        self.refbin, self.threads = refbin, threads
        self.exe = refbin.ref_binary()
        self.n_genomes, self.genome_len, self.cmp_n = n_genomes, genome_len, cmp_n
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        self.work = tempfile.mkdtemp(prefix="d2bench", dir=base)
        if self.exe is None:
            return
        self.paths = synth.write_fasta_set(os.path.join(self.work, "fa"), n_genomes, genome_len, seed=2, n_families=max(1, n_genomes // 4))
        self.flist = os.path.join(self.work, "files.txt")
        open(self.flist, "w").write("\n".join(self.paths) + "\n")
        regs, cards = synth.synthetic_sketches(cmp_n, S, seed=4, n_families=max(1, cmp_n // 64))
        self.stk = os.path.join(self.work, "cmp.ss")
        synth.write_stacked(self.stk, regs, cards, names=[f"s{i}" for i in range(cmp_n)])
        self.sk_cmd = ["sketch", "-k", str(K), "-w", str(W), "--full-setsketch", "-S", str(S), "-p", str(threads),
                       "-F", self.flist, "-o", os.path.join(self.work, "out.ss")]
        self.cmp_cmd = ["cmp", "--presketched", "--binary-output", "--cmpout", os.path.join(self.work, "out.f32"), "-p", str(threads), self.stk]
        self.kmers = n_genomes * (genome_len - K + 1)
        self.pairs = cmp_n * (cmp_n - 1) // 2

    def _timed(self, cmd):
        t = os.times(); c0 = t.children_user + t.children_system
        t0 = time.perf_counter()
        self.refbin.run_ref(cmd, threads=self.threads)
        wall = time.perf_counter() - t0
        t = os.times()
        return wall, t.children_user + t.children_system - c0

    def run_sketch(self):
        return self._timed(self.sk_cmd)

    def run_cmp(self):
        return self._timed(self.cmp_cmd)

    def describe(self, busy_s, busy_c):
        return (f"sketch: {self.n_genomes} genomes x {self.genome_len} bp (-k31 -w51 --full-setsketch -S4096; FASTA on tmpfs), "
                f"cmp: {self.cmp_n} sketches S=4096 all-pairs symmetric binary; dashing2 v2.1.20 {os.path.basename(self.exe)} -p {self.threads}, "
                f"wall clock of the whole process, median of the timed runs after one discarded run; threads busy (CPU s / wall s): "
                f"sketch {busy_s:.1f}, cmp {busy_c:.1f}")

    def close(self):
        shutil.rmtree(self.work, ignore_errors=True)


def cpu_reference_run(n_genomes, genome_len, cmp_n, threads, repeats=3):
    """Median-of-`repeats` rates of the reference binary on the sample (one discarded warm-up run each)."""
    rs = RefSample(n_genomes, genome_len, cmp_n, threads)
    try:
        if rs.exe is None:
            return cpu_port_run(n_genomes, genome_len, cmp_n)
        rs.run_sketch(); rs.run_cmp()
        ts = [rs.run_sketch() for _ in range(repeats)]
        tc = [rs.run_cmp() for _ in range(repeats)]
        ws = float(np.median([x[0] for x in ts])); wc = float(np.median([x[0] for x in tc]))
        busy_s = float(np.median([x[1] / x[0] for x in ts])); busy_c = float(np.median([x[1] / x[0] for x in tc]))
        return dict(sketch_kmers_s=rs.kmers / ws, cmp_pairs_s=rs.pairs / wc, sketch_s=ws, cmp_s=wc, kind="reference", cores=threads,
                    threads_busy={"sketch": busy_s, "cmp": busy_c}, sample=rs.describe(busy_s, busy_c))
    finally:
        rs.close()


def cpu_port_run(n_genomes, genome_len, cmp_n):
    """Fallback when the reference binary cannot run on this CPU: the single-threaded oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from dashing2_b200 import synth
    L = O.lib()
    n_genomes = min(n_genomes, 2); cmp_n = min(cmp_n, 400)
    t0 = time.perf_counter(); kmers = 0
    for _, s in synth.family_genomes(n_genomes, genome_len, seed=2):
        hv = O.hash_stream(s.tobytes(), K, W)
        regs = np.empty(2 * S - 1); L.d2o_css_reset(regs, S); L.d2o_css_update(regs, S, hv, len(hv), None)
        kmers += genome_len - K + 1
    ts = time.perf_counter() - t0
    regs, cards = synth.synthetic_sketches(cmp_n, S, seed=4)
    t0 = time.perf_counter(); O.allpairs(regs, cards); tc = time.perf_counter() - t0
    return dict(sketch_kmers_s=kmers / ts, cmp_pairs_s=cmp_n * (cmp_n - 1) / 2 / tc, sketch_s=ts, cmp_s=tc, kind="port", cores=1,
                threads_busy={"sketch": 1.0, "cmp": 1.0},
                sample=f"oracle port, 1 thread: {n_genomes} genomes x {genome_len} bp; {cmp_n} sketches")


def sample_sizes(args, cores):
    """The reference's sketch loop is file-parallel: at least 4 genomes per host thread (and at least 64)."""
    return max(args.cpu_genomes, 4 * cores), args.cpu_cmp_n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n_genomes, cmp_n = sample_sizes(args, cores)
    rs = RefSample(n_genomes, args.genome_len, cmp_n, cores)
    try:
        if rs.exe is None:
            res = cpu_port_run(n_genomes, args.genome_len, cmp_n)
            v, vc, ms, sample, kind, busy = res["sketch_kmers_s"], res["cmp_pairs_s"], (res["sketch_s"] + res["cmp_s"]) * 1e3, res["sample"], "port", res["threads_busy"]
            cores = 1
        else:
            ts, tc = [], []
            for i in range(args.warmup + args.steps):        # one step = one run of both reference commands over the sample
                a = rs.run_sketch(); b = rs.run_cmp()
                if i >= args.warmup:
                    ts.append(a); tc.append(b)
            ws = float(np.mean([x[0] for x in ts])); wc = float(np.mean([x[0] for x in tc]))
            busy = {"sketch": float(np.median([x[1] / x[0] for x in ts])), "cmp": float(np.median([x[1] / x[0] for x in tc]))}
            v, vc, ms, kind = rs.kmers / ws, rs.pairs / wc, (ws + wc) * 1e3, "reference"
            sample = rs.describe(busy["sketch"], busy["cmp"])
    finally:
        rs.close()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "kmers/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": "kmers/s", "cores": cores, "kind": kind, "sample": sample, "threads_busy": busy},
            "e2e": {"value": v, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cmp": {"value": vc, "unit": "pairs/s", "e2e": {"value": vc, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cli_vs_cli(n_genomes, genome_len, threads):
    """Whole programs, same argv, FASTA on tmpfs: the reference binary against the drop-in front-end (dashing2-gpu), sketch + all-pairs
    compare with binary output; reports wall clocks (median of 2 after a warm-up) and whether the output files are byte-identical."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refbin
    from dashing2_b200 import synth
    ref = refbin.ref_binary()
    ours = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")
    if ref is None or not os.path.exists(ours):
        return {"unavailable": "reference binary or front-end missing"}
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    work = tempfile.mkdtemp(prefix="d2cli", dir=base)
    try:
        paths = synth.write_fasta_set(os.path.join(work, "fa"), n_genomes, genome_len, seed=2, n_families=max(1, n_genomes // 4))
        flist = os.path.join(work, "files.txt"); open(flist, "w").write("\n".join(paths) + "\n")
        res = {}
        for tag, exe in (("reference", ref), ("ours", ours)):
            argv = ["sketch", "-k", str(K), "-w", str(W), "--full-setsketch", "-S", str(S), "-p", str(threads), "-F", flist,
                    "-o", os.path.join(work, tag + ".ss"), "--cmpout", os.path.join(work, tag + ".f32"), "--binary-output"]
            env = dict(os.environ, OMP_NUM_THREADS=str(threads))
            ts = []
            for i in range(3):
                t0 = time.perf_counter()
                r = subprocess.run([exe] + argv, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                ts.append(time.perf_counter() - t0)
                if r.returncode:
                    return {"unavailable": f"{tag} failed: {r.stderr.decode(errors='replace')[-300:]}"}
            res[tag] = float(np.median(ts[1:]))
        same_mat = open(os.path.join(work, "ours.f32"), "rb").read() == open(os.path.join(work, "reference.f32"), "rb").read()
        a = np.fromfile(os.path.join(work, "ours.ss"), dtype=np.uint64); b = np.fromfile(os.path.join(work, "reference.ss"), dtype=np.uint64)
        n = n_genomes
        same_regs = len(a) == len(b) and bool(np.array_equal(a[2 + n:], b[2 + n:]))     # registers bit for bit; cardinalities to 1e-12 (summation order)
        kmers = n_genomes * (genome_len - K + 1)
        return {"argv": "sketch -k31 -w51 --full-setsketch -S4096 -p %d -F files.txt -o X.ss --cmpout X.f32 --binary-output" % threads,
                "sample": "%d genomes x %d bp, FASTA on tmpfs" % (n_genomes, genome_len),
                "reference_s": res["reference"], "ours_s": res["ours"], "speedup": res["reference"] / res["ours"],
                "reference_kmers_s": kmers / res["reference"], "ours_kmers_s": kmers / res["ours"],
                "matrix_bytes_identical": same_mat, "registers_bit_identical": same_regs,
                "note": "whole-process wall clock; ours includes CUDA context creation and teardown (about 1 s on these boxes)"}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def workload_config(args, n_gpus):
    return {"workload": "BASELINE configs[1]: sketch+cmp, %d synthetic genomes x %d bp per GPU, k=31 w=51, Full SetSketch S=4096, "
                        "all-pairs symmetric" % (args.genomes, args.genome_len),
            "genomes_per_gpu": args.genomes, "genome_len": args.genome_len, "k": K, "w": W, "sketchsize": S,
            "sketch_mode": "--full-setsketch", "cmp": "symmetric all-pairs, similarity, f64 registers",
            "parallelism": "files sharded over %d GPU(s), no collective; cmp rows equal-area sharded, one exchange step inside the library over its own NCCL communicator "
                           "(register all-to-all, each GPU ranks S/N register positions, all-gather of the 32-bit ranks)" % n_gpus,
            "l2": "inputs larger than L2 (sequence buffer %.1f GB, register matrix %.0f MB per GPU)" %
                  (args.genomes * args.genome_len / 1e9, args.genomes * S * 8 / 1e6)}


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []; self.proc = None; self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def make_genomes_on_device(torch, dev, n, length, seed, n_families):
    """ASCII genomes on the device, SURVEY 8(d) recipe (family ancestor + i.i.d. substitutions at rate
    0.001*(g%64+1)); returns uint8[n*length] (+256 B pad)."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    anc = torch.randint(0, 4, (n_families, length), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty(n * length + 256, dtype=torch.uint8, device=dev)
    out[n * length:] = 0
    for i in range(n):
        a = anc[i % n_families]
        rate = 0.001 * (i % 64 + 1)
        mut = torch.rand(length, device=dev, generator=g) < rate
        sub = torch.randint(1, 4, (length,), dtype=torch.uint8, device=dev, generator=g)
        codes = torch.where(mut, (a + sub) & 3, a)
        out[i * length:(i + 1) * length] = lut[codes.long()]
    return out


from dashing2_b200.shard import equal_area_rows  # noqa: E402


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.config != 2:
        import bench_configs
        return bench_configs.run(args, sys.modules[__name__])
    import torch
    import torch.distributed as dist
    from dashing2_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libd2gpu has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Host buffers of the end-to-end legs should live on the NUMA node next to this rank's GPU: bind the process to the GPU's
    # CPU set while they are allocated and used (first-touch placement), restore it for the CPU baseline.
    affinity0 = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:                      # no NVML / not permitted: keep the inherited affinity
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one process per GPU: the ranks of a node share its cores, so each rank's host packer takes its share (the library sizes its
    # host / device packing split from that; a single process keeps every core)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world > 1 and "D2G_HOST_THREADS" not in os.environ:
        os.environ["D2G_HOST_THREADS"] = str(max(1, len(affinity0) // local_world))
    # From three ranks up the host's DRAM, not the PCIe links, bounds the upload (r1: 183 G k-mers/s at N = 8 with plain DMA of ASCII; r2r: 155 with
    # 29 % of every chunk packed by four threads per rank -- packing adds 0.75 B of host memory traffic per base): send ASCII, pack on the device
    if local_world >= 3:
        os.environ.setdefault("D2G_HYBRID_F", "0")
    # the library's split between host packing and device packing of an ASCII chunk (api_sketch.cu: same formula, same inputs)
    host_thr = int(os.environ.get("D2G_HOST_THREADS", "0")) or min(64, len(os.sched_getaffinity(0)))
    hyb_f = float(os.environ["D2G_HYBRID_F"]) if "D2G_HYBRID_F" in os.environ else max(0.1, min(0.9, (1. / 50e9) / (1. / (4.6e9 * host_thr) + 0.75 / 50e9)))
    ctx = capi.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    if world > 1:
        # the library owns its communicator; torch.distributed only carries the 128-byte id to the other ranks
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_rank(world, rank, uid[0])

    G, Lg = args.genomes, args.genome_len
    n_all = G * world
    n_fam = max(1, G * 157 // 10000)
    seq = make_genomes_on_device(torch, dev, G, Lg, seed=2 + rank, n_families=n_fam)
    rec_off = (torch.arange(G + 1, dtype=torch.int64, device=dev) * Lg)
    rec_ent = torch.arange(G, dtype=torch.int32, device=dev)
    sig = torch.empty((G, S), dtype=torch.float64, device=dev)
    card = torch.empty(G, dtype=torch.float64, device=dev)
    all_sig = torch.empty((n_all, S), dtype=torch.float64, device=dev) if world > 1 else sig
    all_card = torch.empty(n_all, dtype=torch.float64, device=dev) if world > 1 else card
    bounds = equal_area_rows(n_all, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    p_sk = ctx.params(mode="fss", S=S, k=K, w=W)
    p_cmp = ctx.cmp_params(S, n_all, "symmetric", "similarity", k=K)
    my_pairs = ctx.cmp_rows_size(p_cmp, r0, r1)
    out = torch.empty(my_pairs, dtype=torch.float32, device=dev)
    kmers_rank = G * max(0, Lg - K + 1)
    torch.cuda.synchronize()

    def step(timed):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record(ext)
        ctx.sketch_batch_dev(p_sk, seq.data_ptr(), rec_off.data_ptr(), rec_ent.data_ptr(), G, G, G * Lg,
                             sig_d=sig.data_ptr(), card_d=card.data_ptr())
        evs[1].record(ext)
        evs[2].record(ext)
        if world > 1:
            # the path's one exchange step, inside the library: register all-to-all + 1/N of the ranking + all-gather of 32-bit ranks (NCCL)
            ctx.cmp_rows_sharded_dev(p_cmp, sig.data_ptr(), card.data_ptr(), rank * G, G, r0, r1, out.data_ptr())
        else:
            ctx.cmp_rows_dev(p_cmp, all_sig.data_ptr(), all_card.data_ptr(), r0, r1, out.data_ptr())
        evs[3].record(ext)
        ext.synchronize()
        return evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    ctx.set_timing(True)
    for c in range(5):
        ctx.get_timing(c)
    sampler = ClockSampler(local); sampler.start()
    l0 = ctx.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    ts, tg, tc = [], [], []
    for _ in range(args.steps):
        a, b, c = step(True)
        ts.append(a); tg.append(b); tc.append(c)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    k_main_ms, k_main_n = ctx.get_timing(0)
    k_boot_ms, k_boot_n = ctx.get_timing(1)
    k_cmp_ms, k_cmp_n = ctx.get_timing(2)
    k_prep_ms, k_prep_n = ctx.get_timing(3)
    k_pack_ms, k_pack_n = ctx.get_timing(4)
    ctx.set_timing(False)

    if world > 1:        # outside the timed region: the full register matrix for the host-buffer leg and the oracle check
        dist.all_gather_into_tensor(all_sig, sig)
        dist.all_gather_into_tensor(all_card, card)
        torch.cuda.synchronize()
    # max over ranks of the device-timed totals
    tot = torch.tensor([sum(ts), sum(tg), sum(tc), sum(ts) + sum(tg) + sum(tc)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    sk_ms, ag_ms, cmp_ms, all_ms = (float(x) for x in tot.tolist())

    # ---- e2e: the public C-ABI call with HOST buffers (pinned), H2D + kernels + D2H inside the timed region
    Ge = min(args.e2e_genomes, G)
    h_seq = torch.empty(Ge * Lg, dtype=torch.uint8).pin_memory()
    h_seq.copy_(seq[:Ge * Lg])
    h_off = (np.arange(Ge + 1, dtype=np.uint64) * np.uint64(Lg))
    h_ent = np.arange(Ge, dtype=np.uint32)
    h_sig = torch.empty((Ge, S), dtype=torch.float64).pin_memory()
    h_card = torch.empty(Ge, dtype=torch.float64).pin_memory()
    import ctypes as C

    def e2e_sketch():
        nk = C.c_uint64(0)
        rc = ctx.L.d2g_sketch_batch(ctx.h, C.byref(p_sk), h_seq.data_ptr(), h_off.ctypes.data, h_ent.ctypes.data, Ge, Ge,
                                    None, h_sig.data_ptr(), h_card.data_ptr(), None, C.byref(nk))
        if rc:
            raise RuntimeError(ctx.L.d2g_last_error().decode())
        return nk.value

    n_e2e_cmp = n_all                      # the same (weak-scaled) triangle as the resident leg, from host registers
    h_regs = torch.empty((n_e2e_cmp, S), dtype=torch.float64).pin_memory(); h_regs.copy_(all_sig[:n_e2e_cmp])
    h_cards = torch.empty(n_e2e_cmp, dtype=torch.float64).pin_memory(); h_cards.copy_(all_card[:n_e2e_cmp])
    p_e2e_cmp = ctx.cmp_params(S, n_e2e_cmp, "symmetric", "similarity", k=K)
    eb = equal_area_rows(n_e2e_cmp, world)
    e_pairs = ctx.cmp_rows_size(p_e2e_cmp, eb[rank], eb[rank + 1])
    h_out = torch.empty(e_pairs, dtype=torch.float32).pin_memory()
    out_np = h_out.numpy()

    # N > 1: every rank hands over only the registers of the sketches it made (its block of the host matrix); the exchange step runs inside
    # the library (d2g_cmp_rows_sharded).  N = 1: the plain host entry point.
    h_regs_np, h_cards_np = h_regs.numpy(), h_cards.numpy()
    loc0, loc1 = rank * G, (rank + 1) * G

    def e2e_cmp():
        if world > 1:
            ctx.cmp_rows_sharded(p_e2e_cmp, h_regs_np[loc0:loc1], h_cards_np[loc0:loc1], loc0, eb[rank], eb[rank + 1], out_np)
        else:
            ctx.cmp_rows(h_regs_np, h_cards_np, p_e2e_cmp, eb[rank], eb[rank + 1], out=out_np)

    e2e_sketch(); e2e_cmp()   # warm (allocations)
    barrier(); t0 = time.perf_counter()
    ne = max(1, min(args.steps, 3))
    for _ in range(ne):
        nk = e2e_sketch()
    barrier(); t_e2e_sk = (time.perf_counter() - t0) / ne
    t0 = time.perf_counter()
    for _ in range(ne):
        e2e_cmp()
    barrier(); t_e2e_cmp = (time.perf_counter() - t0) / ne
    te = torch.tensor([t_e2e_sk, t_e2e_cmp], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e_sk, t_e2e_cmp = (float(x) for x in te.tolist())

    # ---- e2e with the sequence packed by the caller (d2g_sketch_batch_packed): what the front-end sends once it packs while parsing
    from dashing2_b200 import capi as _capi
    Lc = ctx.L
    nwp = int(Lc.d2g_packed_words(Ge * Lg))
    h_codes = torch.empty(nwp, dtype=torch.int64).pin_memory(); h_mask = torch.empty(nwp, dtype=torch.int32).pin_memory()
    pp = (C.c_void_p * 1)(h_seq.data_ptr()); pl = np.array([Ge * Lg], dtype=np.uint64); nzw = C.c_uint64(0)
    t0 = time.perf_counter()
    if Lc.d2g_pack_sequences(pp, pl.ctypes.data, 1, h_codes.data_ptr(), h_mask.data_ptr(), C.byref(nzw)):
        raise RuntimeError(Lc.d2g_last_error().decode())
    t_host_pack = time.perf_counter() - t0

    def e2e_sketch_packed():
        nk = C.c_uint64(0)
        rc = Lc.d2g_sketch_batch_packed(ctx.h, C.byref(p_sk), h_codes.data_ptr(), None if nzw.value == 0 else h_mask.data_ptr(), h_off.ctypes.data,
                                        h_ent.ctypes.data, Ge, Ge, None, h_sig.data_ptr(), h_card.data_ptr(), None, C.byref(nk))
        if rc:
            raise RuntimeError(Lc.d2g_last_error().decode())
    e2e_sketch_packed()
    barrier(); t0 = time.perf_counter()
    for _ in range(ne):
        e2e_sketch_packed()
    barrier(); t_e2e_pk = (time.perf_counter() - t0) / ne
    tp = torch.tensor([t_e2e_pk], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    t_e2e_pk = float(tp.item())

    # ---- the run checks its own outputs against the oracle (test infrastructure, never the thing timed): the registers of one genome
    # of this rank, and 1000 sampled pairs of the matrix rows this rank computed
    verify = None
    if not args.no_verify and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        Lo = O.lib()
        g = G // 2
        hv = O.hash_stream(seq[g * Lg:(g + 1) * Lg].cpu().numpy().tobytes(), K, W)
        oregs = np.empty(2 * S - 1); Lo.d2o_css_reset(oregs, S); Lo.d2o_css_update(oregs, S, hv, len(hv), None)
        regs_ok = bool(np.array_equal(sig[g].cpu().numpy().view(np.uint64), oregs[:S].view(np.uint64)))
        card_ok = bool(abs(float(card[g]) - Lo.d2o_css_card(oregs, S)) <= 1e-12 * abs(float(card[g])))
        rng = np.random.default_rng(7)
        ii = rng.integers(r0, max(r0 + 1, r1), size=1000); jj = rng.integers(0, n_all, size=1000)
        keep = (ii < jj) & (ii < r1)
        ii, jj = ii[keep], jj[keep]
        tri0 = r0 * n_all - r0 * (r0 + 1) // 2
        idx = ii * n_all - ii * (ii + 1) // 2 + (jj - ii - 1) - tri0
        got = out[torch.from_numpy(idx).to(dev)].cpu().numpy()
        ra = all_sig[torch.from_numpy(ii).to(dev)].cpu().numpy(); rb = all_sig[torch.from_numpy(jj).to(dev)].cpu().numpy()
        ca = all_card[torch.from_numpy(ii).to(dev)].cpu().numpy(); cb2 = all_card[torch.from_numpy(jj).to(dev)].cpu().numpy()
        Lo.d2o_finalize.restype = C.c_float
        exp = np.array([Lo.d2o_finalize(int((ra[t] > rb[t]).sum()), int((ra[t] < rb[t]).sum()), S, float(ca[t]), float(cb2[t]), 0, K, 0)
                        for t in range(len(ii))], dtype=np.float32)
        pairs_ok = bool(np.array_equal(got.view(np.uint32), exp.view(np.uint32)))
        # the host-buffer legs produced the same bytes as the resident ones (registers of the first Ge genomes; this rank's rows of the matrix)
        e2e_regs_ok = bool(torch.equal(h_sig.view(torch.int64), sig[:Ge].cpu().view(torch.int64)))
        e2e_rows_ok = bool(torch.equal(h_out.view(torch.int32), out.cpu().view(torch.int32)))
        verify = {"genome": int(g), "registers_bit_identical_to_oracle": regs_ok, "cardinality_within_1e-12": card_ok,
                  "pairs_checked": int(len(ii)), "pairs_bit_identical_to_oracle": pairs_ok,
                  "e2e_registers_equal_resident": e2e_regs_ok, "e2e_rows_equal_resident": e2e_rows_ok}
        regs_ok = regs_ok and e2e_regs_ok; pairs_ok = pairs_ok and e2e_rows_ok
        if not (regs_ok and card_ok and pairs_ok):
            raise SystemExit("bench.py: outputs differ from the oracle: %s" % json.dumps(verify))

    try:
        os.sched_setaffinity(0, affinity0)
    except OSError:
        pass
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # DRAM traffic of the two dominant kernels, measured once under ncu at exactly this configuration (profiles/)
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        tc = tj.get("config", {})
        if tc.get("genomes_per_gpu") == G and tc.get("genome_len") == Lg and tc.get("sketchsize") == S and tc.get("n_gpus") == world:
            traffic = tj
    except (OSError, ValueError):
        pass
    steps = args.steps
    total_kmers = kmers_rank * world
    total_pairs = n_all * (n_all - 1) // 2
    sketch_rate = total_kmers * steps / (sk_ms / 1e3)
    cmp_rate = total_pairs * steps / (cmp_ms / 1e3)
    # roofline of the dominant kernel of each path: algorithmic bytes per launch / average launch duration
    sk_bytes = kmers_rank * 1.0                                   # 1 B per k-mer
    sk_ach = sk_bytes / (k_main_ms / max(1, k_main_n) / 1e3) / 1e9
    cmp_bytes = n_all * S * 8 + 4.0 * my_pairs                    # every register read once + every result written once
    cmp_ach = cmp_bytes / (k_cmp_ms / max(1, k_cmp_n) / 1e3) / 1e9
    line = {
        "metric": METRIC, "value": sketch_rate, "unit": "kmers/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": all_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64+f64", "data": "synthetic", "config": workload_config(args, world),
        "phases_ms_per_step": {"sketch": sk_ms / steps, "allgather": ag_ms / steps, "cmp": cmp_ms / steps, "wall": wall_ms / steps},
        "roofline": {"bound": "hbm", "achieved": sk_ach, "peak": hbm_peak, "unit": "GB/s", "frac": sk_ach / hbm_peak, "traffic": traffic.get("sketch_main_bytes_per_launch"),
                     "kernel": "sketch_fast_kernel<21, FssMainConsumer> (32-bit window keys)", "launch_ms": k_main_ms / max(1, k_main_n),
                     "algorithmic_bytes_per_launch": sk_bytes, "peak_source": peak_src,
                     "note": "1 B per k-mer position; the kernel is integer-ALU bound (~hundreds of int ops per k-mer), see DESIGN.md",
                     "boot_kernel_ms": k_boot_ms / max(1, k_boot_n)},
        "cmp": {"value": cmp_rate, "unit": "pairs/s", "pairs_per_step": total_pairs, "ms_per_step": cmp_ms / steps,
                "roofline": {"bound": "hbm", "achieved": cmp_ach, "peak": hbm_peak, "unit": "GB/s", "frac": cmp_ach / hbm_peak, "traffic": traffic.get("cmp_tile_bytes_per_launch"),
                             "kernel": "cmp16_tile_kernel<ne, imad> (16-bit order codes; S is a power of two)", "launch_ms": k_cmp_ms / max(1, k_cmp_n),
                             "algorithmic_bytes_per_launch": cmp_bytes,
                             "no_reuse_bytes_per_pair": 2 * S * 8,
                             "code_prep_ms_per_step": k_prep_ms / steps,
                             "note": "code_prep = keys + per-register segmented radix sort + rank kernels that turn f64 registers into order codes; "
                                     "included in cmp.ms_per_step and cmp.value, not in launch_ms"},
                "e2e": {"value": (n_e2e_cmp * (n_e2e_cmp - 1) // 2) / t_e2e_cmp, "unit": "pairs/s",
                        "h2d_bytes_per_step": (G if world > 1 else n_e2e_cmp) * (S * 8 + 8), "d2h_bytes_per_step": e_pairs * 4,
                        "call": ("d2g_cmp_rows_sharded (this rank's block of pinned host registers in, exchange inside the library, its rows of the matrix copied into a pinned host buffer while later rows compute)"
                                 if world > 1 else "d2g_cmp_rows (pinned host registers in, float32 rows copied into a pinned host buffer while later rows compute)"), "n": n_e2e_cmp}},
        "e2e": {"value": Ge * (Lg - K + 1) * world / t_e2e_sk, "unit": "kmers/s",
                "h2d_bytes_per_step": int(Ge * Lg * (hyb_f * 0.25 + (1. - hyb_f))) + (Ge + 1) * 8 + Ge * 4,
                "d2h_bytes_per_step": Ge * S * 8 + Ge * 8,
                "call": "d2g_sketch_batch (pinned host ASCII buffers in, host registers out; of every 384 Mi-base chunk the library packs the first %.0f %% to 2 bits "
                        "per base on its %d host threads and sends the rest as ASCII by DMA to be packed on the device, while earlier chunks sketch)" % (100 * hyb_f, host_thr),
                "host_threads": host_thr, "host_packed_fraction": hyb_f,
                "batch": "%d genomes x %d bp per call" % (Ge, Lg)},
        "gpu_launches": launches, "clocks": clocks,
    }
    line["roofline"]["pack_kernel_ms"] = k_pack_ms / max(1, k_pack_n)
    line["roofline"]["pack_kernel_note"] = ("ASCII -> 2 bit + invalid bit packing on the device (reads 1 B, writes 0.375 B per base) runs before the main kernel "
                                            "inside every step; its bytes are extra, not a discount (SURVEY 8(d))")
    line["e2e"]["packed"] = {"value": Ge * (Lg - K + 1) * world / t_e2e_pk, "unit": "kmers/s", "h2d_bytes_per_step": nwp * (8 if nzw.value == 0 else 12),
                             "call": "d2g_sketch_batch_packed (sequence packed once by d2g_pack_sequences, outside this leg: %.3f s for the batch = %.1f G bases/s on the host threads)"
                                     % (t_host_pack, Ge * Lg / t_host_pack / 1e9)}
    if verify is not None:
        line["verify"] = verify
    if not args.no_cpu_baseline:
        cores = host_cores()
        ng, nc = sample_sizes(args, cores)
        cb = cpu_reference_run(ng, Lg, nc, cores, repeats=3)
        line["cpu_baseline"] = {"value": cb["sketch_kmers_s"], "unit": "kmers/s", "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
                                "threads_busy": cb["threads_busy"]["sketch"]}
        line["cmp"]["cpu_baseline"] = {"value": cb["cmp_pairs_s"], "unit": "pairs/s", "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
                                       "threads_busy": cb["threads_busy"]["cmp"]}
        if not args.no_cli and world == 1:
            line["e2e_cli"] = cli_vs_cli(ng, Lg, cores)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

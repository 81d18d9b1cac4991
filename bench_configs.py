"""bench.py --config {1,3,4,5}: one JSON line each for the BASELINE configs that are not the headline (configs[1] is bench.py's own
default leg).  Same contract as bench.py: W untimed steps, K steps timed on the device (CUDA events on the library's stream), inputs
resident in HBM for `value`, the C-ABI call with pinned HOST buffers for `e2e`, `roofline` for the dominant kernel from the library's
per-class kernel timers, `cpu_baseline` = the unmodified reference binary (oracle/_ref) on a bounded sample on this box's host cores.
Single GPU (under torchrun only rank 0 works).  Sizes default to the configs' stated sizes; --genomes / --genome-len / --n shrink them."""
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
K = 31
METRIC = "k-mers hashed/s (sketch) + pairwise compares/s (cmp)"


def _peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d.get("hbm_gbs") or d.get("hbm_gbs_burst") or 6539.2), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6539.2, "fallback"


def _roof(kernel, algo_bytes, launch_ms, note, extra=None):
    peak, src = _peaks()
    ach = algo_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "kernel": kernel,
         "launch_ms": launch_ms, "algorithmic_bytes_per_launch": algo_bytes, "peak_source": src, "note": note}
    if extra:
        r.update(extra)
    return r


def _ref_timed(refbin, cmd, threads, repeats=3, before=None):
    """median wall seconds (after one discarded run) and threads busy of the reference binary"""
    def once():
        if before:
            before()
        t = os.times(); c0 = t.children_user + t.children_system
        t0 = time.perf_counter(); refbin.run_ref(cmd, threads=threads); w = time.perf_counter() - t0
        t = os.times(); return w, (t.children_user + t.children_system - c0) / w
    once()
    rs = [once() for _ in range(repeats)]
    return float(np.median([r[0] for r in rs])), float(np.median([r[1] for r in rs]))


def _workdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="d2cfg", dir=base)


def _sketches_on_device(torch, dev, n, S, seed, n_fam):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    base = torch.rand((n_fam, S), dtype=torch.float64, device=dev, generator=g)
    out = torch.empty((n, S), dtype=torch.float64, device=dev)
    step = 8192
    for i in range(0, n, step):
        m = min(step, n - i)
        fam = (torch.arange(i, i + m, device=dev) % n_fam)
        p = 0.05 + 0.9 * torch.rand((m, 1), dtype=torch.float64, device=dev, generator=g)
        fresh = torch.rand((m, S), dtype=torch.float64, device=dev, generator=g)
        keep = torch.rand((m, S), dtype=torch.float64, device=dev, generator=g) >= p
        out[i:i + m] = torch.where(keep, base[fam], fresh)
    return out, torch.full((n,), 1e6, dtype=torch.float64, device=dev)


def run(args, bench):
    import torch
    from dashing2_b200 import capi, synth
    sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbin
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libd2gpu has no CPU fallback)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = capi.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    cores = bench.host_cores()
    cfg = args.config
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # larger than the 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)

    def timed_steps(step_fn, flush):
        for _ in range(args.warmup):
            step_fn()
        ctx.sync(); torch.cuda.synchronize()
        ctx.set_timing(True)
        for c in range(7):
            ctx.get_timing(c)
        sampler = bench.ClockSampler(local); sampler.start()
        l0 = ctx.launch_count()
        tot = 0.
        for _ in range(args.steps):
            if flush:
                flush_l2(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(ext); step_fn(); e1.record(ext); ext.synchronize()
            tot += e0.elapsed_time(e1)
        launches = ctx.launch_count() - l0
        clocks = sampler.stop()
        kt = [ctx.get_timing(c) for c in range(7)]
        ctx.set_timing(False)
        return tot / args.steps, launches, clocks, kt

    line = {"metric": METRIC, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "data": "synthetic"}
    work = _workdir()
    try:
        threads = cores
        if cfg in (1, 3):
            if cfg == 1:
                G, Lg, S, mode = args.genomes or 64, args.genome_len or 1_000_000, 1024, "opmh"
            else:
                G, Lg, S, mode = args.genomes or 2000, args.genome_len or 20_000_000, 8192, "bmh"
            seq = bench.make_genomes_on_device(torch, dev, G, Lg, seed=2 + cfg, n_families=max(1, G // 16))
            m = S + (S & 1)
            p = ctx.params(mode=mode, S=S, k=K)
            per = G if cfg == 1 else min(G, args.batch_genomes)            # counting sketches sort a whole batch at once: batches of genomes
            rec_off = torch.arange(per + 1, dtype=torch.int64, device=dev) * Lg
            rec_ent = torch.arange(per, dtype=torch.int32, device=dev)
            regs = torch.empty((G, m), dtype=torch.int64, device=dev)
            sig = torch.empty((G, S), dtype=torch.float64, device=dev); card = torch.empty(G, dtype=torch.float64, device=dev)

            def step():
                for g0 in range(0, G, per):
                    ng = min(per, G - g0)
                    sp = seq.data_ptr() + g0 * Lg
                    if mode == "opmh":
                        ctx.sketch_batch_dev(p, sp, rec_off.data_ptr(), rec_ent.data_ptr(), ng, ng, ng * Lg, regs_u64_d=regs.data_ptr() + g0 * m * 8)
                    else:
                        ctx.sketch_batch_dev(p, sp, rec_off.data_ptr(), rec_ent.data_ptr(), ng, ng, ng * Lg,
                                             sig_d=sig.data_ptr() + g0 * S * 8, card_d=card.data_ptr() + g0 * 8)
            kmers = G * (Lg - K + 1)
            small = G * Lg < (200 << 20)
            ms, launches, clocks, kt = timed_steps(step, flush=small)
            n_launch = G // per + (1 if G % per else 0)
            if cfg == 1:
                roof = _roof("sketch_kernel<false, OpmhConsumer> (unwindowed exact 2-bit k-mers, bucket minima in shared memory)", float(kmers),
                             kt[0][0] / max(1, kt[0][1]), "1 B per k-mer position (SURVEY 8(d)); integer-ALU bound",
                             {"pack_kernel_ms": kt[4][0] / args.steps})
            else:
                el_ms, sort_ms, emit_ms = kt[1][0] / args.steps, kt[5][0] / args.steps, kt[0][0] / args.steps
                roof = _roof("bmh_kernel (BagMinHash2 split-tree walk, one lane per distinct k-mer, lanes refill from a global counter)",
                             float(per * (Lg - K + 1)) * 12., el_ms / n_launch,
                             "per step: emit %.1f ms, radix sorts + run-length encode %.1f ms (CUB, 12 B per k-mer x 12 passes), element kernel %.1f ms; "
                             "algorithmic bytes = 12 B per k-mer of a batch (hash + entity of the sorted stream read once); the kernel is latency / divergence bound" % (emit_ms, sort_ms, el_ms),
                             {"phases_ms_per_step": {"emit": emit_ms, "sort_rle": sort_ms, "element_kernel": el_ms, "pack": kt[4][0] / args.steps}})
            # e2e: d2g_sketch_batch with pinned host ASCII (one batch), host registers out
            Ge = min(per, G)
            h_seq = torch.empty(Ge * Lg, dtype=torch.uint8).pin_memory(); h_seq.copy_(seq[:Ge * Lg])
            h_off = np.arange(Ge + 1, dtype=np.uint64) * np.uint64(Lg); h_ent = np.arange(Ge, dtype=np.uint32)
            h_sig = torch.empty((Ge, S), dtype=torch.float64).pin_memory(); h_card = torch.empty(Ge, dtype=torch.float64).pin_memory()

            def e2e():
                nk = C.c_uint64(0)
                rc = ctx.L.d2g_sketch_batch(ctx.h, C.byref(p), h_seq.data_ptr(), h_off.ctypes.data, h_ent.ctypes.data, Ge, Ge,
                                            None, h_sig.data_ptr(), h_card.data_ptr(), None, C.byref(nk))
                if rc:
                    raise RuntimeError(ctx.L.d2g_last_error().decode())
            e2e(); t0 = time.perf_counter()
            ne = 3
            for _ in range(ne):
                e2e()
            t_e2e = (time.perf_counter() - t0) / ne
            e2e_line = {"value": Ge * (Lg - K + 1) / t_e2e, "unit": "kmers/s", "h2d_bytes_per_step": Ge * Lg + (Ge + 1) * 8 + Ge * 4,
                        "d2h_bytes_per_step": Ge * S * 8 + Ge * 8,
                        "call": "d2g_sketch_batch (pinned host ASCII in, host f64 registers + cardinalities out%s)" % (", x87 finalisation on the host threads" if cfg == 1 else ""),
                        "batch": f"{Ge} genomes x {Lg} bp per call"}
            # CPU: the reference binary on the same kind of files
            ng = G if cfg == 1 else max(16, 2 * cores)
            paths = synth.write_fasta_set(os.path.join(work, "fa"), ng, Lg, seed=2, n_families=max(1, ng // 4))
            fl = os.path.join(work, "files.txt"); open(fl, "w").write("\n".join(paths) + "\n")
            if cfg == 1:
                cmd = ["sketch", "-k", "31", "-S", "1024", "-p", str(threads), "-F", fl, "-o", os.path.join(work, "o.ss"), "--cmpout", os.path.join(work, "o.phy"), "--phylip"]
            else:
                cmd = ["sketch", "-k", "31", "-S", "8192", "--multiset", "-p", str(threads), "-F", fl, "-o", os.path.join(work, "o.ss")]
            cpu = None
            if refbin.ref_binary() is not None and cfg == 3:
                # without --cache the reference binary (v2.1.20) ends `sketch --multiset` with a segmentation fault after the work is done;
                # with it the per-file sketches are written (and must be removed between runs, or the next run only loads them)
                cdir = os.path.join(work, "cache")

                def wipe():
                    shutil.rmtree(cdir, ignore_errors=True); os.makedirs(cdir)
                cmd = ["sketch", "-k", "31", "-S", "8192", "--multiset", "-p", str(threads), "-F", fl, "-o", os.path.join(work, "o.ss"), "--cache", "--outprefix", cdir]
                w, busy = _ref_timed(refbin, cmd, threads, repeats=2, before=wipe)
                cpu = {"value": ng * (Lg - K + 1) / w, "unit": "kmers/s", "cores": threads, "kind": "reference", "threads_busy": busy,
                       "sample": f"dashing2 sketch -k 31 -S 8192 --multiset -p {threads} --cache ... over {ng} genomes x {Lg} bp (FASTA on tmpfs), whole process wall clock {w:.2f} s, median after one discarded run"}
            elif refbin.ref_binary() is not None:
                w, busy = _ref_timed(refbin, cmd, threads, repeats=3 if cfg == 1 else 2)
                cpu = {"value": ng * (Lg - K + 1) / w, "unit": "kmers/s", "cores": threads, "kind": "reference", "threads_busy": busy,
                       "sample": f"dashing2 {' '.join(cmd[:7])} ... over {ng} genomes x {Lg} bp (FASTA on tmpfs), whole process wall clock {w:.2f} s, median after one discarded run"}
                if cfg == 1:   # the drop-in front-end on the same argv
                    import subprocess
                    exe = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")
                    ocmd = [exe] + cmd[:-5] + ["-o", os.path.join(work, "g.ss"), "--cmpout", os.path.join(work, "g.phy"), "--phylip"]
                    subprocess.run(ocmd, check=True, capture_output=True)
                    t0 = time.perf_counter(); subprocess.run(ocmd, check=True, capture_output=True); tw = time.perf_counter() - t0
                    line["e2e_cli"] = {"reference_s": w, "ours_s": tw, "speedup": w / tw,
                                       "matrix_bytes_identical": open(os.path.join(work, "o.phy"), "rb").read() == open(os.path.join(work, "g.phy"), "rb").read(),
                                       "sketch_file_bytes_identical": open(os.path.join(work, "o.ss"), "rb").read() == open(os.path.join(work, "g.ss"), "rb").read(),
                                       "note": "whole-process wall clock incl. CUDA context creation (about 1 s)"}
            line.update({"value": kmers / (ms * 1e-3), "unit": "kmers/s", "ms_per_step": ms, "dtype": "u64" if cfg == 1 else "u64+f64",
                         "config": {"workload": ("BASELINE configs[0]: 64 synthetic genomes x 1 Mbp, k=31, OPMH S=1024 (sketch leg; the all-pairs PHYLIP matrix is in e2e_cli)" if cfg == 1 else
                                                 "BASELINE configs[2]: --multiset BagMinHash, %d genomes x %d bp, S=8192, batches of %d genomes (a batch is sorted at once)" % (G, Lg, per)),
                                    "genomes": G, "genome_len": Lg, "k": K, "sketchsize": S, "sketch_mode": "--oneperm" if cfg == 1 else "--multiset",
                                    "l2": "L2 flushed between timed steps (256 MB write)" if small else "inputs larger than L2"},
                         "roofline": roof, "e2e": e2e_line, "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu})
        elif cfg == 4:
            nf, nq, S = args.n or 50_000, (args.n * 2 if args.n else 100_000), 1024
            regs, cards = _sketches_on_device(torch, dev, nf + nq, S, 4, max(1, (nf + nq) // 100))
            p = ctx.cmp_params(S, nf + nq, "panel", "similarity", k=K, nq=nq)
            out = torch.empty(nf * nq, dtype=torch.float32, device=dev)

            def step():
                ctx.cmp_rows_dev(p, regs.data_ptr(), cards.data_ptr(), 0, nf, out.data_ptr())
            ms, launches, clocks, kt = timed_steps(step, flush=False)
            pairs = nf * nq
            tile_ms, tile_n = kt[2]; prep_ms = kt[3][0] / args.steps
            roof = _roof("cmp16_tile_kernel (16-bit order codes, 64 x 64 pair tiles staged with cp.async.bulk)",
                         float(pairs) * 4. + float(nf + nq) * S * 2., tile_ms / args.steps,
                         "all tile launches of a step together; algorithmic bytes = 4 B per pair written + 2 B per register code read once; "
                         "the kernel is integer-ALU bound (S compares per pair); code preparation %.1f ms per step on top" % prep_ms,
                         {"tile_launches_per_step": tile_n / args.steps, "code_prep_ms_per_step": prep_ms, "no_reuse_bytes_per_pair": 2 * S * 8})
            # e2e: host registers in, a block of rows out into pinned host memory
            rows_e = min(nf, max(1, int(4e8 // nq)))
            h_regs = torch.empty((nf + nq, S), dtype=torch.float64).pin_memory(); h_regs.copy_(regs)
            h_cards = torch.empty(nf + nq, dtype=torch.float64).pin_memory(); h_cards.copy_(cards)
            h_out = torch.empty(rows_e * nq, dtype=torch.float32).pin_memory(); out_np = h_out.numpy()
            ctx.cmp_rows(h_regs.numpy(), h_cards.numpy(), p, 0, rows_e, out=out_np)
            t0 = time.perf_counter()
            for _ in range(2):
                ctx.cmp_rows(h_regs.numpy(), h_cards.numpy(), p, 0, rows_e, out=out_np)
            t_e2e = (time.perf_counter() - t0) / 2
            e2e_line = {"value": rows_e * nq / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": (nf + nq) * S * 8 + (nf + nq) * 8, "d2h_bytes_per_step": rows_e * nq * 4,
                        "call": f"d2g_cmp_rows (pinned host registers of all {nf + nq} sketches in, rows [0, {rows_e}) of the panel out into pinned host memory)"}
            cpu = None
            if refbin.ref_binary() is not None:
                cn = 8000
                r2, c2 = synth.synthetic_sketches(cn, S, seed=4, n_families=max(1, cn // 64))
                stk = os.path.join(work, "c.ss"); synth.write_stacked(stk, r2, c2, names=[f"s{i}" for i in range(cn)])
                cmd = ["cmp", "--presketched", "--binary-output", "--cmpout", os.path.join(work, "o.f32"), "-p", str(threads), stk]
                w, busy = _ref_timed(refbin, cmd, threads)
                cpu = {"value": cn * (cn - 1) / 2 / w, "unit": "pairs/s", "cores": threads, "kind": "reference", "threads_busy": busy,
                       "sample": f"dashing2 cmp --presketched over {cn} sketches S={S}, all-pairs symmetric binary (the same compare() per pair as the panel), wall {w:.2f} s, median after one discarded run"}
            line.update({"value": pairs / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "dtype": "f64 registers -> u16 order codes -> f32",
                         "config": {"workload": f"BASELINE configs[3]: panel cmp, {nq} query x {nf} reference sketches S={S}, rectangular float32 matrix, one GPU (rows shard over GPUs in bench.py --gpus N)",
                                    "n_ref": nf, "n_query": nq, "sketchsize": S, "l2": "register matrix (1.2 GB) and output (20 GB) larger than L2"},
                         "roofline": roof, "e2e": e2e_line, "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu})
        elif cfg == 5:
            n, S, topk = args.n or 1_000_000, 1024, 32
            regs, cards = _sketches_on_device(torch, dev, n, S, 5, max(1, n // 1000))
            h_regs_t = torch.empty(regs.shape, dtype=torch.float64).pin_memory(); h_regs_t.copy_(regs)
            h_regs = h_regs_t.numpy(); h_cards = cards.cpu().numpy()
            del regs; torch.cuda.empty_cache()
            nnz = [0]; refined = [0]

            p5 = ctx.cmp_params(S, n, "symmetric", "similarity", k=K)
            indptr = np.zeros(n + 1, dtype=np.uint64)

            def step():      # the C entry point itself: host registers in, malloc'ed CSR out (released with d2g_free, not copied again)
                pi = C.POINTER(C.c_uint32)(); pv = C.POINTER(C.c_float)()
                rc = ctx.L.d2g_lsh_topk(ctx.h, C.byref(p5), h_regs.ctypes.data, h_cards.ctypes.data, topk, indptr.ctypes.data, C.byref(pi), C.byref(pv))
                if rc:
                    raise RuntimeError(ctx.L.d2g_last_error().decode())
                nnz[0] = int(indptr[-1]); refined[0] = ctx.stat(0)
                ctx.L.d2g_free(pi); ctx.L.d2g_free(pv)
            # the call is synchronous and host-in / host-out: wall clock == device timeline + copies; events on the stream bracket it too
            ms, launches, clocks, kt = timed_steps(step, flush=False)
            ntab = S + S // 2
            q_ms = kt[2][0] / args.steps; sort_ms = kt[5][0] / args.steps; ref_ms = kt[6][0] / args.steps
            roof = _roof("lsh_refine_kernel (one warp per neighbour-list entry; the owner's row is served by L1 / L2, the neighbour's row streams from HBM)",
                         float(refined[0]) * S * 8. + float(n) * S * 8., ref_ms,
                         "per step: per-table key sort %.1f ms (CUB segmented radix sort), ordered candidate scan %.1f ms, refine %.1f ms; algorithmic bytes of refine = "
                         "S x 8 B per list entry compared (%d entries before trimming) + every sketch's own row once; a million 8 KiB rows do not fit the L2, so this "
                         "kernel is the one HBM-bound kernel of the repository" % (sort_ms, q_ms, ref_ms, refined[0]),
                         {"phases_ms_per_step": {"table_sort": sort_ms, "candidate_scan": q_ms, "refine": ref_ms}, "tables": ntab, "nnz": nnz[0], "entries_refined": refined[0]})
            cpu = None
            if refbin.ref_binary() is not None:
                cn = 20000
                r2, c2 = synth.synthetic_sketches(cn, S, seed=5, n_families=max(1, cn // 1000))
                stk = os.path.join(work, "c.ss"); synth.write_stacked(stk, r2, c2, names=[f"s{i}" for i in range(cn)])
                cmd = ["cmp", "--presketched", "--binary-output", "--topk", "32", "--cmpout", os.path.join(work, "o.csr"), "-p", str(threads), stk]
                w, busy = _ref_timed(refbin, cmd, threads, repeats=2)
                cpu = {"value": cn / w, "unit": "sketches/s", "cores": threads, "kind": "reference", "threads_busy": busy,
                       "sample": f"dashing2 cmp --presketched --topk 32 over {cn} sketches S={S} (index build + candidate scan + refine), wall {w:.2f} s, median after one discarded run"}
            v = n / (ms * 1e-3)
            line.update({"value": v, "unit": "sketches/s", "ms_per_step": ms, "dtype": "f64 registers, u32 LSH keys, f32 distances",
                         "config": {"workload": f"BASELINE configs[4]: --topk {topk} LSH neighbour graph over {n} pre-built sketches S={S} (candidate generation + refine), one GPU",
                                    "n": n, "sketchsize": S, "topk": topk, "l2": "register matrix (8 GB) larger than L2"},
                         "roofline": roof,
                         "e2e": {"value": v, "unit": "sketches/s", "h2d_bytes_per_step": n * S * 8 + n * 8, "d2h_bytes_per_step": nnz[0] * 8 + (n + 1) * 8,
                                 "call": "d2g_lsh_topk (pinned host registers in, CSR out): the entry point is host-in / host-out, so value and e2e are the same measurement"},
                         "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu})
        else:
            raise SystemExit("--config: 1, 3, 4 or 5 (2 is bench.py's default leg)")
    finally:
        shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(line), flush=True)

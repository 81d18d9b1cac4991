#!/usr/bin/env python
"""Golden vectors for the compressed-register compare (--fastcmp N [--bbit-sigs]); UNMODIFIED reference binary.
Dev container only (needs oracle/_ref); writes tests/golden/expected/cmpc_*.npy and the fitted (a, b) the
reference printed.  Same inputs as the other presketched goldens (inputs/sk48x256.npz)."""
import json, os, re, shutil, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402
from make_golden import CMP_CASES  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")


def main():
    work = tempfile.mkdtemp(prefix="d2goldc")
    z = np.load(os.path.join(INP, "sk48x256.npz"))
    stk = os.path.join(work, "sk48.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(48)])
    ab = {}
    exe = refbin.ref_binary()
    for fd in ("1", "2", "4"):
        for bbit in (False, True):
            for kind in ("sim_sym", "sim_asym", "containment_sym", "symcontainment_sym", "mash_sym", "isz_sym", "usz_sym"):
                mat = os.path.join(work, "o.f32")
                argv = [exe, "cmp", "--presketched", "-p1", "--binary-output", "--cmpout", mat, "--fastcmp", fd] + (["--bbit-sigs"] if bbit else []) + CMP_CASES[kind] + [stk]
                r = subprocess.run(argv, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
                assert r.returncode == 0, r.stderr[-2000:]
                tag = f"cmpc_sk48_fd{fd}_{'bbit' if bbit else 'ss'}_{kind}"
                np.save(os.path.join(EXP, tag + ".npy"), np.fromfile(mat, dtype=np.float32))
                m = re.search(r"a = ([0-9.eE+-]+) and b = ([0-9.eE+-]+)", r.stderr)
                if m: ab[f"fd{fd}"] = [m.group(1), m.group(2)]
    json.dump(ab, open(os.path.join(EXP, "cmpc_fitted_ab.json"), "w"), indent=1)
    shutil.rmtree(work)
    print("fitted a,b:", ab)


if __name__ == "__main__":
    main()

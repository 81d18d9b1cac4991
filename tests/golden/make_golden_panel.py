#!/usr/bin/env python
"""Golden vectors for the PANEL shape through the command line (`-F refs -Q queries`, src/emitrect.cpp:229-246: rows = references,
columns = queries); UNMODIFIED reference binary.  Dev container only (needs oracle/_ref).  Inputs: the committed FASTA fixtures, the
first four as references, the last four as queries."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
from make_golden import materialise  # noqa: E402
EXP = os.path.join(HERE, "expected")
CASES = {"panel_opmh_k31_S1024_sim": ["-k31", "-S1024"], "panel_opmh_k31_S1024_containment": ["-k31", "-S1024", "--containment"],
         "panel_fss_k31_S256_mash": ["-k31", "-S256", "--full-setsketch", "--mash-distance"]}


def main():
    work = tempfile.mkdtemp(prefix="d2goldp")
    names, paths = materialise(work)
    nf = 4
    ff = os.path.join(work, "refs.txt"); open(ff, "w").write("\n".join(paths[:nf]) + "\n")
    qf = os.path.join(work, "queries.txt"); open(qf, "w").write("\n".join(paths[nf:]) + "\n")
    for name, argv in CASES.items():
        mat = os.path.join(work, name + ".f32"); txt = os.path.join(work, name + ".txt")
        refbin.run_ref(["sketch", "-p1", "-F", ff, "-Q", qf, "--binary-output", "--cmpout", mat] + argv, threads=1)
        refbin.run_ref(["sketch", "-p1", "-F", ff, "-Q", qf, "--cmpout", txt] + argv, threads=1, cwd=work)
        m = np.fromfile(mat, dtype=np.float32)
        assert m.size == nf * (len(paths) - nf)
        np.save(os.path.join(EXP, name + ".npy"), m)
        open(os.path.join(EXP, name + ".txt"), "w").write(open(txt).read().replace(work + "/", ""))
        print(name, m[:6])
    shutil.rmtree(work)


if __name__ == "__main__":
    main()

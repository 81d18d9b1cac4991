#!/bin/sh
# Regenerates every golden fixture from the UNMODIFIED reference binary (dev container only: needs /root/reference to build oracle/_ref).
# The committed fixtures under tests/golden/expected were produced by exactly these scripts.
set -e
cd "$(dirname "$0")/../.."
make -s -f oracle/Makefile.ref -j8
for s in make_golden make_golden_compressed make_golden_byseq make_golden_mincount make_golden_panel make_golden_countsketch \
         make_golden_weighted_ids make_golden_fss_ids make_golden_nlsh make_golden_rolling make_golden_threshold make_golden_protein make_golden_kmercounts make_golden_topk_fastcmp; do
    echo "== $s"; python tests/golden/$s.py
done

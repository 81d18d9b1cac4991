#!/bin/bash
# entity-bit sort passes: the tests that go emit -> sort, and the config-3 kernel timing on 8 genomes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -rf -k "count or weighted or bmh or pmh or multiset or distinct or filterset or mincount or byseq or parse_by_seq or contain or kmer or config3 or stream" 2>&1 | tail -8 > gpurun_out/r2s_pytest.txt; tail -3 gpurun_out/r2s_pytest.txt
timeout 300 python scripts/sketch_only_bench.py 8 20000000 3 bmh 8192 -1 > gpurun_out/r2s_weighted.txt 2>&1; tail -2 gpurun_out/r2s_weighted.txt
timeout 300 python scripts/sketch_only_bench.py 64 20000000 2 bmh 8192 -1 >> gpurun_out/r2s_weighted.txt 2>&1; tail -1 gpurun_out/r2s_weighted.txt

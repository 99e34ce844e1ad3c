#!/bin/bash
# Last evidence pass of round 1: full GPU suite, smoke, the default bench line, the per-GPU share of config 5
# (ML-KEM-1024, 2^21 ops) and the reference arm, all with the final library.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c.log 2>&1; tail -1 gpurun_out/pytest_gpu_c.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
python bench.py --workload mlkem1024 --batch-log2 21 --no-ntt > gpurun_out/bench_mlkem1024_r01c.json 2> gpurun_out/bench_mlkem1024_r01c.err
python bench.py --impl reference --steps 3 > gpurun_out/bench_ref_r01c.json 2>&1
tail -c 300 gpurun_out/bench_mlkem1024_r01c.json

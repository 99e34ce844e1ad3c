#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/ubench_pipes.cu -o /tmp/ubench && /tmp/ubench > gpurun_out/ubench_pipes.txt 2>&1
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --batch-log2 17 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full capture: forward/inverse NTT kernels (2^20 batch)
ncu --set full --clock-control none --import-source on -k regex:ntt_kernel -s 4 -c 2 -o gpurun_out/prof_ntt \
    python scripts/time_ring.py > gpurun_out/prof_ntt.log 2>&1
# full capture: ML-KEM sampler + encrypt kernels
ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|encrypt_kernel|hash_ek" -s 6 -c 3 -o gpurun_out/prof_mlkem \
    python bench.py --steps 1 --warmup 3 --batch-log2 16 --no-cpu-baseline --no-ntt > gpurun_out/prof_mlkem.log 2>&1
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
ls -la gpurun_out

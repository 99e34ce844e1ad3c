#!/bin/bash
# Round-1 (second pass) profiling recipe: run under gpurun on one GPU; outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_b.csv \
    python bench.py --steps 1 --warmup 3 --batch-log2 17 --no-cpu-baseline > gpurun_out/bench_under_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|encrypt_kernel|hash_ek" -s 6 -c 3 -o gpurun_out/prof_mlkem_b \
    python bench.py --steps 1 --warmup 3 --batch-log2 16 --no-cpu-baseline --no-ntt > gpurun_out/prof_mlkem_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mldsa" -s 3 -c 9 -o gpurun_out/prof_mldsa \
    python bench.py --workload mldsa65 --steps 1 --warmup 3 --batch-log2 14 --no-cpu-baseline > gpurun_out/prof_mldsa.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ntt_kernel" -s 4 -c 4 -o gpurun_out/prof_ntt_b \
    python scripts/time_ring.py > gpurun_out/prof_ntt_b.log 2>&1
python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err
python bench.py --workload mldsa65 > gpurun_out/bench_mldsa_r01b.json 2> gpurun_out/bench_mldsa_r01b.err
python bench.py --impl reference --steps 3 > gpurun_out/bench_ref_r01b.json 2>&1
python scripts/bench_ops.py > gpurun_out/bench_ops_r01b.json 2>&1
ls -la gpurun_out

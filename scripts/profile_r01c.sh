#!/bin/bash
# Round-1 (third pass) evidence run: one GPU under gpurun; everything lands in gpurun_out/ and is summarised into profiles/r01c_*.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c.log 2>&1
tail -3 gpurun_out/pytest_gpu_c.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_c.csv \
    python bench.py --steps 1 --warmup 3 --batch-log2 17 --no-cpu-baseline > gpurun_out/bench_under_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|encrypt_kernel|hash_ek" -s 6 -c 3 -o gpurun_out/prof_mlkem_c \
    python bench.py --steps 1 --warmup 3 --batch-log2 16 --no-cpu-baseline --no-ntt > gpurun_out/prof_mlkem_c.log 2>&1
# (the ncu reports of the signing kernels and of x25519_kernel are taken by scripts/profile_r01c_sign.sh: gpurun_out/ is
#  limited to 64 MiB per call)
ncu --set full --clock-control none --import-source on -k regex:"ntt_kernel|ntt_fwd_tma_kernel" -s 4 -c 6 -o gpurun_out/prof_ntt_c \
    python scripts/time_ring.py > gpurun_out/prof_ntt_c.log 2>&1
python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
python bench.py --workload mldsa65 > gpurun_out/bench_mldsa_r01c.json 2> gpurun_out/bench_mldsa_r01c.err
python bench.py --impl reference --steps 3 > gpurun_out/bench_ref_r01c.json 2>&1
python scripts/bench_ops.py > gpurun_out/bench_ops_r01c.json 2>&1
tail -c 1500 gpurun_out/bench_r01c.json
ls -la gpurun_out | tail -20

#!/bin/bash
# Round-2 evidence run: one GPU under gpurun; everything lands in gpurun_out/ and is summarised into profiles/r02z_*.
# (ncu numbers are for the kernel analysis only; every bench value comes from the un-profiled bench.py runs below.)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02z_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err
timeout 600 python bench.py --workload mlkem1024 --batch-log2 21 --no-cpu-baseline --no-ntt --no-extras > gpurun_out/r02z_bench_mlkem1024.json 2> gpurun_out/r02z_mlkem1024.err
timeout 600 python scripts/bench_ops.py > gpurun_out/r02z_bench_ops.json 2>&1
# launch list of the default bench command (cold-cache, serialised per-launch times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02z_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --batch-log2 17 --no-cpu-baseline --no-extras > gpurun_out/r02z_bench_under_ncu.log 2>&1
# full captures of the pipeline kernels (one sub-batch each), the ring kernels and the signing round
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|sample_fix_kernel|encrypt_dp_kernel|hash_ek_kernel" -s 8 -c 4 \
    -o gpurun_out/r02z_prof_mlkem python bench.py --steps 1 --warmup 3 --batch-log2 16 --no-cpu-baseline --no-ntt --no-extras > gpurun_out/r02z_prof_mlkem.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ntt_fwd_kernel|ntt_inv_kernel" -s 1 -c 2 \
    -o gpurun_out/r02z_prof_ntt python scripts/ntt_once.py > gpurun_out/r02z_prof_ntt.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"^ntt_kernel$" -s 12 -c 2 \
    -o gpurun_out/r02z_prof_ring python scripts/time_ring.py > gpurun_out/r02z_prof_ring.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"mask_kernel|yntt_kernel|w_kernel|challenge_kernel|response_kernel" -c 7 \
    -o gpurun_out/r02z_prof_sign python scripts/sign_once.py > gpurun_out/r02z_prof_sign.log 2>&1
# DRAM traffic of the signing kernels over the WHOLE rejection loop of one 2^15 batch (all rounds: attempts of one op that
# run in the same round share its A-hat through L2)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"w_kernel|yntt_kernel|response_kernel|mask_kernel|expand_a_kernel|challenge_kernel" --csv \
    --log-file gpurun_out/r02z_sign_traffic.csv python scripts/sign_once.py > gpurun_out/r02z_sign_traffic.log 2>&1
# compute-sanitizer: scripts/sanitize_r02.sh, run as a gpurun call of its own
tail -c 1200 gpurun_out/r02z_bench.json
ls -la gpurun_out | tail -20

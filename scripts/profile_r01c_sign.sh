#!/bin/bash
# Round-1 (third pass), second call: ncu reports of the ML-DSA signing kernels (first round of a 2^15 batch) and of x25519_kernel.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:"mask_kernel|yntt_kernel|w_kernel|challenge_kernel|cntt_mask|response_kernel|finalize_kernel" -c 10 \
    -o gpurun_out/prof_sign_e python scripts/sign_once.py > gpurun_out/prof_sign_e.log 2>&1
ncu --set full --clock-control none -k regex:"x25519_kernel" -c 1 -o gpurun_out/prof_x25519 \
    python scripts/x25519_once.py 65536 > gpurun_out/prof_x25519.log 2>&1
ls -la gpurun_out | tail

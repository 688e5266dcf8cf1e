#!/usr/bin/env bash
# scripts/gpu_profile_aux.sh -- ncu evidence for the kernels outside bench.py's pipeline: the ECC iteration
# (scripts/ecc_bench.py) and the lossy pre-conditioner (scripts/lossy_bench.py).  Each command runs plain first
# and must exit 0 (B200_PROFILING.md); the numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
TAG="${TAG:-r1}"
ECC="python scripts/ecc_bench.py --frames 40"
$ECC > gpurun_out/ecc_plain_${TAG}.jsonl 2> gpurun_out/ecc_plain_${TAG}.err &&
ncu --set full --clock-control none --import-source on -k "regex:ecc_solve_kernel|ecc_iter_kernel" -s 60 -c 4 -f \
    -o gpurun_out/prof_ecc_${TAG} $ECC > gpurun_out/ncu_ecc_${TAG}.log 2>&1
echo "ecc capture rc=$?"
tail -2 gpurun_out/ncu_ecc_${TAG}.log
ls -la gpurun_out/ | tail -8

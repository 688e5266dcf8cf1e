#!/usr/bin/env bash
# scripts/gpu_profile.sh -- ncu evidence for one round: (1) the launch list of a short bench run
# (per-launch device time), (2) one `--set full` capture of the hot kernels.  The same command
# runs plain first and must exit 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
TAG="${TAG:-r1}"
CMD="${CMD:-python bench.py --chunk 1000 --steps 2 --warmup 3 --no-cpu --e2e-frames 250 --no-configs --no-e2e-all --stream-frames 0}"
KERN="${KERN:-bp_correct_kernel|gauss_tile_kernel|translate_u16_rows_kernel|split_stats_kernel}"
NK="${NK:-4}"
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:${KERN}|bp_detect_kernel|hist_frame_kernel|stats_init_kernel|hist_quantile_kernel" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.json 2>> gpurun_out/plain_${TAG}.err &&
ncu --set full --clock-control none --import-source on -k "regex:${KERN}" -s $((3 * NK)) -c ${NK} -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/

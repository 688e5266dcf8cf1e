#!/usr/bin/env bash
# scripts/gpu_check.sh -- what one gpurun call does on the B200 box: smoke, GPU parity tests, a
# short bench.  Every step is bounded by `timeout` so a bad kernel cannot hold the box.
# Logs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 ${PYTEST_EXTRA:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps "${BENCH_STEPS:-5}" --warmup 3 ${BENCH_EXTRA:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err

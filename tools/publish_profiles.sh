#!/bin/bash
# Copies the evidence of tools/refresh_profiles.sh from gpurun_out/ (scratch) into profiles/ (tracked) and rebuilds the summary.
set -eu
for f in r2_bench_default.json r2_bench_c2.json r2_bench_c4.json r2_bench_reference.json r2_launches_bench.csv r2_ncu_full.csv; do
  cp gpurun_out/$f profiles/$f
done
python tools/make_profile_summary.py profiles/r2_launches_bench.csv profiles/r2_ncu_full.csv "Round 2, final kernels" > profiles/r2_summary.md
tail -2 gpurun_out/r2_pytest_gpu.log > profiles/r2_pytest_gpu.txt
tail -1 gpurun_out/r2_smoke.log >> profiles/r2_pytest_gpu.txt

#!/bin/bash
# Round-end evidence in ONE gpurun call (B200, 1 GPU): GPU tests, the bench lines, the ncu launch list of the bench command
# and one `ncu --set full` capture of a whole step.  Everything lands in gpurun_out/; tools/publish_profiles.sh copies the
# summaries into profiles/.   usage: gpurun --timeout 1500 -- 'bash tools/refresh_profiles.sh'
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1; tail -1 $O/r2_pytest_gpu.log
python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1; tail -1 $O/r2_smoke.log
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err
python bench.py --config 2 --no-cpu-baseline > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err
python bench.py --config 4 > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_launches_bench.log 2>&1
K='^k_(ba_solve|copy_level0|resize_quads|fast_cells|blur|quadtree|describe|expand|hamming_umma|hamming_decode)$'
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 48 -c 16 -f -o $O/r2_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_full.log 2>&1
ncu -i $O/r2_full.ncu-rep --page raw --csv > $O/r2_ncu_full.csv 2> /dev/null
python - <<'PY'
import json
for f in ("default", "c2", "c4", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), d["unit"], "e2e", round(d.get("e2e", {}).get("value", 0)))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
wc -l $O/r2_launches_bench.csv $O/r2_ncu_full.csv

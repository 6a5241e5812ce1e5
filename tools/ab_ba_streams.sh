#!/bin/bash
# A/B: register-capped BA builds (tools/libslamb200_mbN.so) x back-end batches in flight (BENCH_BA_STREAMS)
for v in ${VARIANTS:-base mb2 mb3}; do for n in ${STREAMS:-1 2 3}; do
  if [ "$v" = "base" ]; then unset SLAMB200_LIB; else export SLAMB200_LIB=$PWD/tools/libslamb200_$v.so; fi
  BENCH_BA_STREAMS=$n python bench.py --steps 120 --warmup 4 --no-cpu-baseline > gpurun_out/abs_${v}_$n.json 2> gpurun_out/abs_${v}_$n.err
  python - "$v" "$n" <<'PY'
import json, sys
v, n = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/abs_{v}_{n}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"{v:5s} streams {n}: value {d['value']:8.0f} e2e {d['e2e']['value']:8.0f} ms/step {d['ms_per_step']:.3f}  ba ser {r['stage_ms_per_step_serialised']['local_ba']:.3f} conc {r['stage_ms_per_step_concurrent']['local_ba']:.3f}")
except Exception as ex:
    print(v, n, "FAILED", ex, open(f"gpurun_out/abs_{v}_{n}.err").read()[-400:])
PY
done; done

#!/bin/bash
# A/B of experimental library builds (tools/libslamb200_<name>.so, see csrc/Makefile) through bench.py: prints value / e2e / BA stage per variant.
for v in "$@"; do
  if [ "$v" = "base" ]; then unset SLAMB200_LIB; else export SLAMB200_LIB=$PWD/tools/libslamb200_$v.so; fi
  python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{v}.json").read().strip().split("\n")[-1])
    st = d["roofline"].get("stage_ms_per_step_serialised") or d["roofline"]["stage_ms_per_step"]
    print(f"{v:8s} value {d['value']:9.0f}  e2e {d['e2e']['value']:9.0f}  ms/step {d['ms_per_step']:.3f}  ba {st.get('local_ba', 0):.3f}  fast {st['fast_cells']:.3f}  describe {st['describe']:.3f} blur {st['gauss_blur']:.3f}")
except Exception as ex:
    print(v, "FAILED", ex, open(f"gpurun_out/ab_{v}.err").read()[-500:])
PY
done

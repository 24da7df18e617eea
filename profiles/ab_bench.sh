#!/bin/bash
# A/B timing helper: runs bench.py (device-resident arm only) for the given workloads and prints one short line each.
# usage: profiles/ab_bench.sh "<label>" "<env assignments>" c4 c4s ...
label="$1"; envs="$2"; shift 2
for wl in "$@"; do
  out=$(env $envs python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$out" > gpurun_out/ab_${label}_${wl}.json
  python - "$label" "$wl" <<'PY' "$out"
import json, sys
label, wl, raw = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.loads(raw)
    print(f"{label:10s} {wl:4s} ms/frame {d['ms_per_step']:8.3f}  splat {d['phases_ms']['splat']:8.3f}  Gp/s {d['value']:7.2f}  hbm frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(label, wl, "FAILED", raw[-300:])
PY
done

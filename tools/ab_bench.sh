#!/bin/bash
# A/B runs of bench.py under environment switches: `tools/ab_bench.sh name "ENV=.. ENV=.." "bench args" ...` (triples).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
while [ $# -ge 3 ]; do
  name=$1; envs=$2; bargs=$3; shift 3
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $bargs > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    line = [l for l in open(f"gpurun_out/ab_{name}.json") if l.startswith("{")][-1]
    j = json.loads(line)
    k = j.get("kernels") or {}
    top = ", ".join(f"{n}={v['ms_per_step']:.3f}" for n, v in list(k.items())[:7])
    print(f"{name}: ms/step {j['ms_per_step']:.3f} serial {j['pipeline']['serial_ms_per_step']} value {j['value']/1e6:.2f}M "
          f"e2e {(j.get('e2e') or {}).get('value', 0)/1e6:.2f}M launches {j['gpu_launches']} | {top}")
except Exception as e:
    print(name, "FAILED", e)
    print(open(f"gpurun_out/ab_{name}.err").read()[-1500:])
PY
done

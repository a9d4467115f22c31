#!/bin/bash
# Final measurement pass of a round (run on the GPU box through gpurun): the whole GPU test-suite, then every named
# workload with the default settings; JSON lines land in gpurun_out/ and are copied to profiles/ by hand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/${R}_bench_iqap_b1024.json 2> gpurun_out/${R}_b.err
python bench.py --workload fa --steps 8 > gpurun_out/${R}_bench_fa_b4096.json 2> gpurun_out/${R}_fa.err
python bench.py --workload e2e --steps 6 > gpurun_out/${R}_bench_e2e_b4096.json 2> gpurun_out/${R}_e2e.err
python bench.py --pipeline-depth 1 --blocks 1 --no-cpu-baseline > gpurun_out/${R}_bench_iqap_b1024_serial.json 2> gpurun_out/${R}_serial.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_iqap_reference_cpu.json 2> gpurun_out/${R}_ref.err
python - "$R" <<'PY'
import json, sys
R = sys.argv[1]
for n in ["bench_iqap_b1024", "bench_fa_b4096", "bench_e2e_b4096", "bench_iqap_b1024_serial", "bench_iqap_reference_cpu"]:
    try:
        j = json.loads([l for l in open(f"gpurun_out/{R}_{n}.json") if l.startswith("{")][-1])
        e = j.get("e2e") or {}
        print(n, "ms/step", round(j.get("ms_per_step", 0), 3), "value", round(j["value"]), "e2e", e.get("value") and round(e["value"]),
              e.get("upload"), "median", (j.get("blocks") or {}).get("ms_per_step_median"), "roof", (j.get("roofline") or {}).get("frac"),
              "flops", (j.get("roofline") or {}).get("model_flops_frac_of_bf16_peak"), "clocks", j.get("clocks"))
    except Exception as ex:
        print(n, "FAILED", ex)
PY

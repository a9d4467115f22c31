timeout 200 python -m pytest tests/test_iqap_gpu.py -m gpu -x -q 2>&1 | tail -2
for m in auto fp32; do python bench.py --blocks 2 --no-cpu-baseline --e2e-upload $m > gpurun_out/e2e_$m.json 2> gpurun_out/e2e_$m.err; done
python bench.py --blocks 2 --no-cpu-baseline --pipeline-depth 3 > gpurun_out/e2e_auto_d3.json 2>&1
python bench.py --blocks 2 --no-cpu-baseline --pipeline-depth 1 > gpurun_out/e2e_auto_d1.json 2>&1
python - <<PY
import json
for n in ["e2e_auto","e2e_fp32","e2e_auto_d3","e2e_auto_d1"]:
    try:
        j=json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "ms/step", round(j["ms_per_step"],3), "value", round(j["value"]), "e2e", j["e2e"])
    except Exception as e: print(n, "FAILED", e)
PY

timeout 600 python -m pytest tests/test_parity_full_size_gpu.py tests/test_fa_gpu.py tests/test_iqap_gpu.py -m gpu -x -q 2>&1 | tail -2
for m in auto fp32; do python bench.py --workload fa --steps 8 --blocks 2 --no-cpu-baseline --e2e-upload $m > gpurun_out/fa_up_$m.json 2> gpurun_out/fa_up_$m.err; done
python bench.py --blocks 2 --no-cpu-baseline > gpurun_out/iq_up.json 2>/dev/null
python - <<'PY'
import json
for n in ["fa_up_auto", "fa_up_fp32", "iq_up"]:
    try:
        j = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "ms/step", round(j["ms_per_step"], 2), "value", round(j["value"]), "e2e", j["e2e"])
    except Exception as ex:
        print(n, "FAILED", ex); print(open(f"gpurun_out/{n}.err").read()[-800:])
PY

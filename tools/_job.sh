TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2_bench_iqap_8gpu.json 2> gpurun_out/r2_bench_iqap_8gpu.err
$TR bench.py --gpus 8 --workload e2e --batch 18750 --steps 3 --skip-host-e2e --no-cpu-baseline > gpurun_out/r2_bench_e2e_150k_8gpu.json 2> gpurun_out/r2_bench_e2e_150k_8gpu.err
$TR bench.py --gpus 8 --workload fa --steps 6 --blocks 3 --no-cpu-baseline > gpurun_out/r2_bench_fa_8gpu.json 2> gpurun_out/r2_bench_fa_8gpu.err
python - <<'PY'
import json
for n in ["r2_bench_iqap_8gpu", "r2_bench_e2e_150k_8gpu", "r2_bench_fa_8gpu"]:
    try:
        j = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        e = j.get("e2e") or {}
        print(n, "ms/step", round(j["ms_per_step"], 3), "value", round(j["value"]), "e2e", e.get("value") and round(e["value"]), e.get("upload"),
              "ms e2e", e.get("ms_per_step"), "h2d", (j.get("context") or {}).get("h2d_gbs_concurrent"))
    except Exception as ex:
        print(n, "FAILED", ex); print(open(f"gpurun_out/{n}.err").read()[-800:])
PY

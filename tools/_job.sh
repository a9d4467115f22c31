for cfg in "16 1024" "8 1024" "4 1024" "8 512" "12 512" "16 512" "6 512"; do set -- $cfg
B200VQA_HOST_THREADS=$1 python bench.py --steps 20 --blocks 1 --no-cpu-baseline --e2e-chunk $2 > gpurun_out/up_$1_$2.json 2>/dev/null
python - "$1" "$2" <<PY
import json, sys
j=json.loads([l for l in open(f"gpurun_out/up_{sys.argv[1]}_{sys.argv[2]}.json") if l.startswith("{")][-1])
print("threads", sys.argv[1], "chunk", sys.argv[2], "e2e", round(j["e2e"]["value"]), "ms", round(j["e2e"]["ms_per_step"],2), j["e2e"]["upload"])
PY
done
python bench.py --steps 20 --blocks 1 --no-cpu-baseline --e2e-upload fp32 > gpurun_out/up_fp32.json 2>/dev/null
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/up_fp32.json") if l.startswith("{")][-1])
print("fp32", "e2e", round(j["e2e"]["value"]), "ms", round(j["e2e"]["ms_per_step"],2))
PY

timeout 300 python -m pytest tests/test_iqap_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/e2e_trace.py fp16 512 5 2 bg 2>&1 | grep upload | tail -1
python tools/e2e_trace.py fp16 512 5 2 fg 2>&1 | grep upload | tail -1
python tools/e2e_trace.py fp16 512 5 3 bg 2>&1 | grep upload | tail -1
python tools/e2e_trace.py fp16 1024 5 2 bg 2>&1 | grep upload | tail -1
python tools/e2e_trace.py fp16 512 10 2 bg 2>&1 | grep upload | tail -1
for i in 1 2 3; do python bench.py --no-cpu-baseline --blocks 1 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(j['e2e']['ms_per_step'],2), round(j['e2e']['value']))"; done

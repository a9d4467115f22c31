timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
A="--blocks 3 --skip-host-e2e"
tools/ab_bench.sh tma "A=1" "$A" tma_serial "A=1" "--blocks 2 --skip-host-e2e --pipeline-depth 1" fa_tma "A=1" "--workload fa --steps 6 --blocks 3 --skip-host-e2e" | cut -c1-330
python tools/microbench_mem_attn.py 128 256 512 1024 4096 2>&1 | grep "len 243"

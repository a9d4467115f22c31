timeout 600 python -m pytest tests/test_fa_gpu.py tests/test_kernels_gpu.py tests/test_parity_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
A="--workload fa --steps 6 --blocks 2 --skip-host-e2e"
tools/ab_bench.sh fa_attn "A=1" "$A" | cut -c1-330

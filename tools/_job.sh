python tools/_mb.py 2>&1 | grep -v Warn | grep "epi=2"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fa_gpu.py tests/test_iqap_gpu.py -m gpu -x -q 2>&1 | tail -2
tools/ab_bench.sh lntab "A=1" "--blocks 2 --skip-host-e2e" fa_lntab "A=1" "--workload fa --steps 6 --blocks 2 --skip-host-e2e" | cut -c1-400

import sys
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo")
import microbench_gemm as m
for cfg in [(128, 256, 256, 2, 256), (128, 768, 256, 0, 256), (262144, 256, 256, 2, 256), (262144, 768, 256, 0, 256)]:
    m.run(*cfg)

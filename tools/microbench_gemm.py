"""GPU micro-benchmark of the tcgen05 GEMM family through the C ABI's debug entry point (dev tool)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from explainable_spatial_vqa_b200 import _native as nat

def run(M, N, K, epi, bn, iters=200):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / 16).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    gamma = torch.ones(N, device="cuda"); beta = torch.zeros(N, device="cuda")
    out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    a = nat.DbgGemmArgs()
    a.epilogue, a.tf32, a.block_n, a.M, a.N, a.K = epi, 0, bn, M, N, K
    a.A, a.W, a.bias, a.out, a.ldc = A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), N
    a.residual, a.gamma, a.beta = res.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.rows_in = a.rows_out = 1
    lib = nat.lib()
    s = nat.stream_ptr()
    for _ in range(10):
        nat.check(lib.b200vqa_dbg_gemm(C.byref(a), s), "g")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.b200vqa_dbg_gemm(C.byref(a), s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"M={M:7d} N={N:5d} K={K:5d} epi={epi} bn={bn:3d}: {us:8.2f} us  {2*M*N*K/us/1e6:8.1f} TFLOP/s")
    if M <= 1024:
        clk = torch.zeros(32, dtype=torch.int64, device="cuda")
        a.clk = clk.data_ptr()
        for _ in range(3):
            lib.b200vqa_dbg_gemm(C.byref(a), s)
        torch.cuda.synchronize()
        c = clk.cpu().tolist()
        names = ["entry", "setup done", "pdl_wait done", "W landed", "A[0] landed", "A[last] landed", "acc ready (epi)",
                 "epi stores issued", "stores drained", "exit"]
        print("   stage cycles since entry:", ", ".join(f"{n}={c[i] - c[0]}" for i, n in enumerate(names) if c[i]))
        extra = [f"#{i}={c[i] - c[0]}" for i in range(len(names), 32) if c[i]]
        if extra:
            print("   extra stamps:", ", ".join(extra))

if __name__ == "__main__":
    cfgs = [(1024, 256, 256, 2, 256), (1024, 256, 256, 2, 64), (128, 256, 256, 2, 64), (1024, 256, 256, 0, 256), (1024, 256, 256, 0, 64), (1024, 768, 256, 0, 64),
            (128, 256, 256, 2, 256), (128, 256, 256, 0, 64), (1024, 256, 2048, 2, 256),
            (262144, 2048, 256, 1, 256), (262144, 256, 2048, 2, 256), (262144, 768, 256, 0, 256), (262144, 256, 256, 2, 256)]
    if len(sys.argv) > 1:
        cfgs = cfgs[: int(sys.argv[1])]
    for cfg in cfgs:
        run(*cfg)

"""GPU: the tensor-core building blocks, called through the C ABI's debug entry points, against plain
PyTorch fp32 math on the same (bf16- or tf32-rounded) operands."""
import ctypes as C

import pytest
import torch

import common  # noqa: F401
from explainable_spatial_vqa_b200 import _native as nat

pytestmark = pytest.mark.gpu


def dbg_gemm(A, W, bias=None, epilogue=0, block_n=256, residual=None, gamma=None, beta=None, want_f32=False,
             rows_in=1, rows_out=1, row_off=0, pe=None, pe_off=0, out_rows=None):
    tf32 = A.dtype == torch.float32
    M, K = A.shape
    N = W.shape[0]
    out = torch.zeros(out_rows or M, N, dtype=torch.bfloat16, device=A.device)
    out_f32 = torch.zeros(M, N, dtype=torch.float32, device=A.device) if want_f32 else None
    a = nat.DbgGemmArgs()
    a.epilogue, a.tf32, a.block_n, a.M, a.N, a.K = epilogue, int(tf32), block_n, M, N, K
    a.A, a.W, a.bias, a.out, a.ldc = A.data_ptr(), W.data_ptr(), None if bias is None else bias.data_ptr(), out.data_ptr(), N
    a.residual = None if residual is None else residual.data_ptr()
    a.gamma = None if gamma is None else gamma.data_ptr()
    a.beta = None if beta is None else beta.data_ptr()
    a.out_f32 = None if out_f32 is None else out_f32.data_ptr()
    a.rows_in, a.rows_out, a.row_off, a.pe_off = rows_in, rows_out, row_off, pe_off
    a.pe = None if pe is None else pe.data_ptr()
    nat.check(nat.lib().b200vqa_dbg_gemm(C.byref(a), nat.stream_ptr()), "dbg_gemm")
    torch.cuda.synchronize()
    return (out, out_f32) if want_f32 else out


def ref_mm(A, W, bias):
    return A.float() @ W.float().t() + (0 if bias is None else bias)


@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (256, 768, 256, 256), (1000, 512, 256, 128),
                                     (4096 + 77, 2048, 256, 256), (300, 256, 2048, 256), (64, 256, 256, 256)])
@pytest.mark.parametrize("relu", [False, True])
def test_gemm_bf16_bias(M, N, K, bn, relu):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g)).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = dbg_gemm(A, W, bias, epilogue=1 if relu else 0, block_n=bn)
    ref = ref_mm(A, W, bias)
    if relu:
        ref = ref.relu()
    assert common.rel_err(out.float(), ref) < 6e-3  # bf16 output rounding (2^-9) of O(1) values


def test_gemm_matches_cuda_core_check_kernel():
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 777, 768, 256
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / 16).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    chk = torch.empty(M, N, device="cuda")
    nat.check(nat.lib().b200vqa_dbg_gemm_check(0, A.data_ptr(), W.data_ptr(), bias.data_ptr(), chk.data_ptr(), M, N, K,
                                               nat.stream_ptr()), "dbg_gemm_check")
    out = dbg_gemm(A, W, bias)
    torch.cuda.synchronize()
    assert common.rel_err(chk, ref_mm(A, W, bias)) < 1e-5
    assert common.rel_err(out.float(), chk) < 6e-3


@pytest.mark.parametrize("B", [5, 1, 3, 7, 200])   # 8 / 2 / 5 / 11 / 307 m-tiles: even and odd counts (pair mode's phantom tile)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_pe_remap_matches_image_projection_layout(B, dtype):
    """image_proj epilogue: row (item, pos) of the GEMM lands at item*256 + 1 + pos with bias + pe[1+pos].  The kernel
    serves two m-tiles per pass over the weight (pair mode); fp32 features run as tf32, 16-bit features as bf16."""
    g = torch.Generator(device="cuda").manual_seed(9 + B)
    P, K, N = 196, 1024, 256
    A = torch.randn(B * P, K, device="cuda", generator=g).relu().to(dtype)
    W = (torch.randn(N, K, device="cuda", generator=g) / 32).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    pe = torch.randn(243, N, device="cuda", generator=g)
    out = dbg_gemm(A, W, bias, epilogue=3, rows_in=P, rows_out=256, row_off=1, pe=pe, pe_off=1, out_rows=B * 256)
    ref = (A.double() @ W.double().t()).float() + bias
    ref = ref.view(B, P, N) + pe[1:1 + P][None]
    got = out.view(B, 256, N)
    assert common.rel_err(got[:, 1:1 + P].float(), ref) < 6e-3
    assert float(got[:, 0].abs().max()) == 0.0 and float(got[:, 1 + P:].abs().max()) == 0.0  # untouched rows


@pytest.mark.parametrize("M,K", [(128, 256), (1000, 256), (515, 2048), (3, 512)])
def test_gemm_residual_layernorm(M, K):
    g = torch.Generator(device="cuda").manual_seed(M + K)
    N = 256
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    beta = torch.randn(N, device="cuda", generator=g)
    out, out32 = dbg_gemm(A, W, bias, epilogue=2, residual=res, gamma=gamma, beta=beta, want_f32=True)
    ref = torch.nn.functional.layer_norm(ref_mm(A, W, bias) + res.float(), (N,), gamma, beta, 1e-5)
    assert common.rel_err(out32, ref) < 2e-4
    assert common.rel_err(out.float(), ref) < 6e-3


@pytest.mark.parametrize("M,K", [(1, 256), (128, 256), (130, 256), (1000, 256), (4096, 256), (128, 512), (300, 512),
                                 (1, 1024), (128, 1024), (1000, 1024)])
def test_gemm_residual_layernorm_cluster(M, K):
    """Decode-sized variant: four CTAs per 128-row tile, LayerNorm statistics exchanged through cluster shared memory
    (block_n = 64 selects it); residual rows carry a large common offset to exercise the variance merge.  K = 512 / 1024
    (the fused value + output projection of the absorbed cross-attention) stream k-blocks through an 8-stage ring."""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    N = 256
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = (torch.randn(M, N, device="cuda", generator=g) + 6.0).bfloat16()
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    beta = torch.randn(N, device="cuda", generator=g)
    out, out32 = dbg_gemm(A, W, bias, epilogue=2, block_n=64, residual=res, gamma=gamma, beta=beta, want_f32=True)
    ref = torch.nn.functional.layer_norm(ref_mm(A, W, bias) + res.float(), (N,), gamma, beta, 1e-5)
    assert common.rel_err(out32, ref) < 2e-4
    assert common.rel_err(out.float(), ref) < 6e-3
    # same inputs through the persistent kernel: the two paths agree to fp32 rounding
    out_p, out32_p = dbg_gemm(A, W, bias, epilogue=2, block_n=256, residual=res, gamma=gamma, beta=beta, want_f32=True)
    assert common.rel_err(out32, out32_p) < 1e-5


def ref_attention(qkv, lens, nhead):
    B = qkv.shape[0] // 256
    x = qkv.float().view(B, 256, 3, nhead, 256 // nhead)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / (256 // nhead) ** 0.5
    dead = torch.arange(256, device=qkv.device)[None] >= lens[:, None]
    s = s.masked_fill(dead[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, 256, 256)


@pytest.mark.parametrize("nhead", [4, 2])
@pytest.mark.parametrize("v_mode", [0, 1])
def test_encoder_attention(nhead, v_mode):
    g = torch.Generator(device="cuda").manual_seed(nhead * 10 + v_mode)
    lens = torch.tensor([243, 197, 217, 237, 256, 200, 129, 128], dtype=torch.int32, device="cuda")
    B = len(lens)
    qkv = torch.randn(B * 256, 768, device="cuda", generator=g).bfloat16()
    out = torch.full((B * 256, 256), 7.0, dtype=torch.bfloat16, device="cuda")
    nat.check(nat.lib().b200vqa_dbg_enc_attention(qkv.data_ptr(), lens.data_ptr(), 0, B, nhead, v_mode, out.data_ptr(),
                                                  nat.stream_ptr()), "dbg_enc_attention")
    torch.cuda.synchronize()
    ref = ref_attention(qkv, lens, nhead)
    got = out.float().view(B, 256, 256)
    for b in range(B):
        n = int(lens[b])
        assert common.rel_err(got[b, :n], ref[b, :n]) < 1.5e-2, (b, n)  # P rounded to bf16 before the PV product
    assert torch.isfinite(got).all()


def test_encoder_attention_constant_len_matches_lens_array():
    g = torch.Generator(device="cuda").manual_seed(3)
    B = 3
    qkv = torch.randn(B * 256, 768, device="cuda", generator=g).bfloat16()
    o1 = torch.zeros(B * 256, 256, dtype=torch.bfloat16, device="cuda")
    o2 = torch.zeros_like(o1)
    lens = torch.full((B,), 243, dtype=torch.int32, device="cuda")
    nat.check(nat.lib().b200vqa_dbg_enc_attention(qkv.data_ptr(), None, 243, B, 4, 0, o1.data_ptr(), nat.stream_ptr()), "a")
    nat.check(nat.lib().b200vqa_dbg_enc_attention(qkv.data_ptr(), lens.data_ptr(), 0, B, 4, 0, o2.data_ptr(), nat.stream_ptr()), "b")
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)


def ref_mem_attn(qp, mem, lens, nhead):
    """u[b, h] = softmax_j(q'[b, h] . m[b, j] / sqrt(dh)) applied to the memory rows themselves (MemAttnParams)."""
    B = qp.shape[0]
    q = qp.float().view(B, nhead, 256)
    m = mem.float().view(B, 256, 256)
    s = torch.einsum("bhd,bjd->bhj", q, m) / (256 // nhead) ** 0.5
    dead = torch.arange(256, device=qp.device)[None] >= lens[:, None]
    s = s.masked_fill(dead[:, None, :], float("-inf"))
    return torch.einsum("bhj,bjd->bhd", torch.softmax(s, -1), m).reshape(B, nhead * 256)


def dbg_mem_attn(qp, mem, lens, const_len, nhead, impl):
    B = qp.shape[0]
    out = torch.full((B, nhead * 256), 7.0, dtype=torch.bfloat16, device="cuda")
    nat.check(nat.lib().b200vqa_dbg_mem_attn(qp.data_ptr(), mem.data_ptr(), None if lens is None else lens.data_ptr(),
                                             const_len, B, nhead, impl, out.data_ptr(), None, nat.stream_ptr()), "dbg_mem_attn")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("nhead", [4, 2])
def test_memory_attention_kernel(nhead, impl=0):
    """Absorbed decode cross-attention (warp-MMA ring kernel) against fp32 torch math on the same bf16 operands: ragged
    lengths incl. a single row, exactly / just over a tile boundary, the IQAP length and the maximum."""
    g = torch.Generator(device="cuda").manual_seed(100 + nhead)
    lens = torch.tensor([243, 1, 128, 129, 256, 197, 64, 200, 17, 255, 130, 243, 243], dtype=torch.int32, device="cuda")
    B = len(lens)
    mem = torch.randn(B * 256, 256, device="cuda", generator=g).bfloat16()
    qp = (torch.randn(B, nhead * 256, device="cuda", generator=g) * 0.5).bfloat16()  # logits of a few units: peaked rows
    out = dbg_mem_attn(qp, mem, lens, 0, nhead, impl)
    ref = ref_mem_attn(qp, mem, lens, nhead)
    assert torch.isfinite(out.float()).all()
    for b in range(B):
        assert common.rel_err(out[b].float(), ref[b]) < 1.5e-2, (b, int(lens[b]))  # P rounded to bf16 before the product


def test_memory_attention_full_batch_and_constant_len():
    """1024 questions at the IQAP length: persistent CTAs walk several questions each; const_len == lens array."""
    g = torch.Generator(device="cuda").manual_seed(7)
    B = 1024
    mem = torch.randn(B * 256, 256, device="cuda", generator=g).bfloat16()
    qp = (torch.randn(B, 4 * 256, device="cuda", generator=g) * 0.25).bfloat16()
    lens = torch.full((B,), 243, dtype=torch.int32, device="cuda")
    ref = ref_mem_attn(qp, mem, lens, 4)
    for impl in (0,):
        o1 = dbg_mem_attn(qp, mem, None, 243, 4, impl)
        o2 = dbg_mem_attn(qp, mem, lens, 0, 4, impl)
        assert torch.equal(o1, o2)
        err = (o1.float() - ref).abs().amax(1) / ref.abs().amax(1)
        assert float(err.max()) < 1.5e-2, (impl, int(err.argmax()))


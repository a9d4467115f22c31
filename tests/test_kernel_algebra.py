"""CPU: the algebraic identities the decode kernels rely on, restated in plain torch / numpy so that they are checked
without a GPU (the kernels themselves are checked against the oracle in the `-m gpu` tests).

* absorbed query-key and value-output projections of the cross-attention (csrc/kernels.h: launch_absorb_qk,
  launch_absorb_ov; reference arithmetic: torch functional.py multi_head_attention_forward as called from
  inference_transformer_iqap.py:132-133,223-227);
* the split softmax of the tcgen05 memory-attention kernels (csrc/decode_kernels.cu: mem_attn_tc_kernel,
  mem_attn_ring_tc_kernel): per-half maxima / sums / unnormalised products combined at the end;
* the transposed warp reduction of the fused vocabulary head (csrc/ffn_small.cu: ffn_reduce_ln_kernel).
"""
import numpy as np
import torch

import common  # noqa: F401


def _mha(d=256, nhead=4, seed=0):
    torch.manual_seed(seed)
    return torch.nn.MultiheadAttention(d, nhead, bias=True).eval()


def test_absorbed_projections_reproduce_multihead_cross_attention():
    d, nhead, S, B = 256, 4, 37, 3
    dh = d // nhead
    mha = _mha(d, nhead)
    with torch.no_grad():
        mha.in_proj_bias.normal_()
        mha.out_proj.bias.normal_()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, B, d, generator=g)       # one decode position per question
    mem = torch.randn(S, B, d, generator=g)     # encoder memory (seq-first like IQAP)
    with torch.no_grad():
        want, _ = mha(x, mem, mem, need_weights=False)
        W, bvec = mha.in_proj_weight, mha.in_proj_bias
        Wq, Wk, Wv = W[:d], W[d:2 * d], W[2 * d:]
        bq, bv = bvec[:d], bvec[2 * d:]
        Wo, bo = mha.out_proj.weight, mha.out_proj.bias
        # launch_absorb_qk: row h*d + i = sum_e W_k[h*dh+e][i] * W_q[h*dh+e][:], bias likewise with b_q
        w_qk = torch.stack([Wk[h * dh:(h + 1) * dh].t() @ Wq[h * dh:(h + 1) * dh] for h in range(nhead)])  # [h, d, d]
        b_qk = torch.stack([Wk[h * dh:(h + 1) * dh].t() @ bq[h * dh:(h + 1) * dh] for h in range(nhead)])  # [h, d]
        qp = torch.einsum("hid,bd->bhi", w_qk, x[0]) + b_qk                       # absorbed queries [B, h, d]
        # the key bias adds a per-(question, head) constant to every score: softmax-invariant, dropped
        scores = torch.einsum("bhi,sbi->bhs", qp, mem) / dh ** 0.5
        u = torch.einsum("bhs,sbi->bhi", torch.softmax(scores, -1), mem)          # attention-weighted memory
        # launch_absorb_ov: w_ov[n][h*d + i] = sum_e W_o[n][h*dh+e] * W_v[h*dh+e][i], b_ov = b_o + W_o b_v
        w_ov = torch.cat([Wo[:, h * dh:(h + 1) * dh] @ Wv[h * dh:(h + 1) * dh] for h in range(nhead)], 1)  # [d, h*d]
        b_ov = bo + Wo @ bv
        got = u.reshape(B, nhead * d) @ w_ov.t() + b_ov
        # ... and the two-step form the default path runs: grouped value projection, then out_proj
        attn = torch.cat([u[:, h] @ Wv[h * dh:(h + 1) * dh].t() + bv[h * dh:(h + 1) * dh] for h in range(nhead)], 1)
        two_step = attn @ Wo.t() + bo
    assert torch.allclose(got, want[0], rtol=1e-4, atol=1e-4)
    assert torch.allclose(two_step, want[0], rtol=1e-4, atol=1e-4)


def test_split_softmax_combine_equals_full_softmax():
    g = torch.Generator().manual_seed(2)
    for length in (1, 100, 128, 129, 243, 256):
        m = torch.randn(256, 64, generator=g, dtype=torch.float64)
        s = torch.randn(256, generator=g, dtype=torch.float64) * 4
        s[length:] = -float("inf")
        want = torch.softmax(s, 0) @ m
        parts = []
        for half in range(2):
            sh, mh = s[half * 128:(half + 1) * 128], m[half * 128:(half + 1) * 128]
            if half * 128 >= length:                      # a half with no valid row: (max -inf, sum 0, U 0)
                parts.append((torch.tensor(-float("inf"), dtype=torch.float64), torch.tensor(0., dtype=torch.float64),
                              torch.zeros(64, dtype=torch.float64)))
                continue
            mx = sh.max()
            p = torch.exp2((sh - mx) * 1.4426950408889634)
            parts.append((mx, p.sum(), p @ mh))
        (m0, l0, u0), (m1, l1, u1) = parts
        mx = torch.maximum(m0, m1)
        w0, w1 = torch.exp2((m0 - mx) * 1.4426950408889634), torch.exp2((m1 - mx) * 1.4426950408889634)
        got = (w0 * u0 + w1 * u1) / (w0 * l0 + w1 * l1)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-12), length


def test_transposed_warp_reduction_leaves_sum_l_in_lane_l():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((32, 32))       # x[lane][i]: lane's partial of logit i
    want = x.sum(0)
    off = 16
    while off >= 1:
        new = x.copy()
        for lane in range(32):
            upper = bool(lane & off)
            peer = lane ^ off
            peer_upper = bool(peer & off)
            for i in range(off):
                sent = x[peer][i] if peer_upper else x[peer][i + off]   # what the peer sends: the half it does not keep
                keep = x[lane][i + off] if upper else x[lane][i]
                new[lane][i] = keep + sent
        x = new
        off //= 2
    assert np.allclose(x[:, 0], want)

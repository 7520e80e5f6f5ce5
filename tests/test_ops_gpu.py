"""Per-op parity of the CUDA kernels (through the C ABI) against fp32 math on identical bf16 inputs.

Tolerance rule: each kernel's output, given bit-identical bf16 inputs, must be within
1e-3 relative (Frobenius) of the fp32-math result rounded once to bf16 -- stated per test.  Index / copy paths
are bit-exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def rel(a, b):
    """Relative Frobenius error against the fp32-math reference ROUNDED ONCE to the output dtype of `a`
    (bf16 rounding noise alone is ~1.6e-3 relative, so the un-rounded reference cannot be the yardstick)."""
    if a.dtype == BF and b.dtype != BF:
        b = b.to(BF)
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, scale=1.0, seed=0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


@pytest.fixture(scope="module")
def ops():
    from ultravox_b200 import ops as o
    return o


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (201, 256, 512), (1, 128, 256), (300, 192, 200), (1500, 1280, 1280),
                                   (201, 4096, 4096), (129, 64, 64), (201, 28672, 512)])
def test_gemm_plain(ops, M, N, K):
    x, w = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2)
    y = ops.linear(x, w)
    ref = (x.float() @ w.float().T)
    assert y.shape == (M, N) and y.dtype == BF
    assert rel(y, ref.to(BF)) < 1e-3, rel(y, ref)


def test_gemm_exact_bench_shapes(ops):
    """The shapes only bench.py exercised in round 1: Llama-3.1-8B prefill at S=201 - fused qkv
    201x6144x4096, o_proj 201x4096x4096 (+residual, fused RMSNorm), gate|up 201x28672x4096 (ragged 208-wide tiles at full K),
    down 201x4096x14336 (split-K + residual + fused RMSNorm) - and the encoder GEMMs at T=1500 (qkv+bias, fc1+bias+GELU,
    fc2+bias+residual).  fp32 math on the same bf16 inputs (torch fp32 matmul on the GPU, TF32 off), rounded once."""
    torch.backends.cuda.matmul.allow_tf32 = False
    M = 201
    x = rnd(M, 4096, seed=1)
    for N, K, seed in ((6144, 4096, 2), (28672, 4096, 3)):
        w = rnd(N, K, scale=0.02, seed=seed)
        y = ops.linear(x, w)
        assert rel(y, x.float() @ w.float().T) < 1e-3, (N, K)
    for N, K, seed in ((4096, 4096, 4), (4096, 14336, 5)):
        xa, w = rnd(M, K, seed=seed), rnd(N, K, scale=0.02, seed=seed + 10)
        r, nw = rnd(M, N, seed=seed + 20), rnd(N, seed=seed + 30)
        h, xn = r.clone(), torch.empty(M, N, dtype=BF, device="cuda")
        ops.linear(xa, w, residual=h, out=h, norm=(nw, 1e-5, xn))
        assert rel(h, xa.float() @ w.float().T + r.float()) < 1e-3, (N, K)
        assert torch.equal(xn, ops.rmsnorm(h, nw, 1e-5)), (N, K)
    T = 1500
    xe = rnd(T, 1280, seed=6)
    wq, bq = rnd(3840, 1280, scale=0.02, seed=7), rnd(3840, seed=8)
    assert rel(ops.linear(xe, wq, bias=bq), xe.float() @ wq.float().T + bq.float()) < 1e-3
    w1, b1 = rnd(5120, 1280, scale=0.02, seed=9), rnd(5120, seed=10)
    f1 = ops.linear(xe, w1, bias=b1, act=ops.ACT_GELU)
    assert rel(f1, F.gelu(xe.float() @ w1.float().T + b1.float())) < 1e-3
    w2, b2, r2 = rnd(1280, 5120, scale=0.02, seed=11), rnd(1280, seed=12), rnd(T, 1280, seed=13)
    assert rel(ops.linear(f1, w2, bias=b2, residual=r2), f1.float() @ w2.float().T + b2.float() + r2.float()) < 1e-3


@pytest.mark.parametrize("cfg,splits", [(2128, 4), (2256, 3), (1128, 5), (1064, 2), (2064, 7), (1256, 2), (2208, 1), (1208, 1)])
def test_gemm_forced_configs_and_split_k(ops, cfg, splits):
    from ultravox_b200 import _lib
    M, N, K = 201, 512, 2048
    x, w, b, r = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    ref = x.float() @ w.float().T + b.float() + r.float()
    _lib.lib().uvx_debug_gemm_override(cfg, splits)
    try:
        y1 = ops.linear(x, w, bias=b, residual=r)
        y2 = ops.linear(x, w, bias=b, residual=r)
    finally:
        _lib.lib().uvx_debug_gemm_override(0, 0)
    assert rel(y1, ref) < 1e-3
    assert torch.equal(y1, y2)  # split-K reduction order is fixed -> bitwise reproducible


@pytest.mark.parametrize("cfg", [4128, 4256, 5416, 5512, 9128])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1500, 1280, 1280), (201, 512, 4096), (700, 768, 200)])
def test_gemm_two_sm_pairs(ops, cfg, M, N, K):
    """cta_group::2 kernel (cluster of 2 CTAs, 256-row pair tiles; 5xxx = two accumulators per pair tile, ragged N),
    all epilogue features, M / K tails."""
    from ultravox_b200 import _lib
    x, w, b, r = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    ref = F.gelu(x.float() @ w.float().T + b.float()) + r.float()
    _lib.lib().uvx_debug_gemm_override(cfg, 1)
    try:
        y = ops.linear(x, w, bias=b, act=ops.ACT_GELU, residual=r)
        y32 = ops.linear(x, w, out_dtype=torch.float32)
    finally:
        _lib.lib().uvx_debug_gemm_override(0, 0)
    assert rel(y, ref) < 1e-3
    assert rel(y32, x.float() @ w.float().T) < 1e-5


def test_gemm_deep_k_one_wave_takes_the_pair_kernel(ops):
    """The encoder fc2 shape (1500 x 1280 x 5120, bias + in-place residual) is routed to the 2-SM pair kernel with 8 epilogue warps:
    fp32 math on the same inputs, and the same call on the 1-SM kernel (forced) within bf16 rounding."""
    from ultravox_b200 import _lib
    M, N, K = 1500, 1280, 5120
    x, w, b, r = rnd(M, K, seed=1), rnd(N, K, scale=0.03, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    ref = x.float() @ w.float().T + b.float() + r.float()
    h = r.clone()
    ops.linear(x, w, bias=b, residual=h, out=h)
    assert rel(h, ref) < 1e-3
    _lib.lib().uvx_debug_gemm_override(1128, 1)
    try:
        h1 = r.clone()
        ops.linear(x, w, bias=b, residual=h1, out=h1)
    finally:
        _lib.lib().uvx_debug_gemm_override(0, 0)
    assert rel(h, h1.float()) < 2e-3
    y = ops.linear(x, w, bias=b, act=ops.ACT_GELU)
    assert rel(y, F.gelu(x.float() @ w.float().T + b.float())) < 1.5e-3


def test_gemm_epilogues(ops):
    M, N, K = 333, 384, 320
    x, w, b, r = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    ref = F.gelu(x.float() @ w.float().T + b.float()) + r.float()
    y = ops.linear(x, w, bias=b, act=ops.ACT_GELU, residual=r)
    assert rel(y, ref) < 1e-3
    # in-place residual (out aliases residual), fp32 output, alpha
    y32 = ops.linear(x, w, out_dtype=torch.float32, alpha=0.5)
    assert y32.dtype == torch.float32 and rel(y32, 0.5 * (x.float() @ w.float().T)) < 1e-5
    r2 = r.clone()
    ops.linear(x, w, bias=b, residual=r2, out=r2)
    assert rel(r2, x.float() @ w.float().T + b.float() + r.float()) < 1e-3


@pytest.mark.parametrize("cfg,splits", [(0, 0), (2128, 4), (1128, 1), (4128, 1), (1064, 3)])
def test_gemm_fused_rmsnorm(ops, cfg, splits):
    """o_proj / down_proj shape: C = x@w.T + residual (in place) and norm_out = RMSNorm(C); the fused split-K pass must be
    bit-identical to running uvx_rmsnorm on the finished C."""
    from ultravox_b200 import _lib
    M, N, K = 201, 1024, 4096
    x, w, r, nw = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(M, N, seed=4), rnd(N, seed=5)
    h, xn = r.clone(), torch.empty(M, N, dtype=BF, device="cuda")
    _lib.lib().uvx_debug_gemm_override(cfg, splits)
    try:
        ops.linear(x, w, residual=h, out=h, norm=(nw, 1e-5, xn))
    finally:
        _lib.lib().uvx_debug_gemm_override(0, 0)
    assert rel(h, x.float() @ w.float().T + r.float()) < 1e-3
    assert torch.equal(xn, ops.rmsnorm(h, nw, 1e-5))


@pytest.mark.parametrize("cm,cn", [(1, 2), (2, 1), (2, 2), (1, 4)])
@pytest.mark.parametrize("cfg,splits,M,N,K", [(2208, 1, 201, 1024, 1024), (2128, 3, 201, 768, 2048), (1128, 1, 1500, 1280, 640),
                                              (1256, 1, 700, 1280, 200), (1064, 2, 130, 192, 1024)])
def test_gemm_multicast_clusters(ops, cm, cn, cfg, splits, M, N, K):
    """Thread-block clusters sharing operand tiles by TMA multicast (A across cn column tiles, W across cm row tiles),
    including cluster padding tiles (tile counts not divisible by the cluster shape) and split-K."""
    from ultravox_b200 import _lib
    x, w, b, r = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    ref = x.float() @ w.float().T + b.float() + r.float()
    _lib.lib().uvx_debug_gemm_override(cfg, splits)
    _lib.lib().uvx_debug_gemm_cluster(cm, cn)
    try:
        y = ops.linear(x, w, bias=b, residual=r)
        y2 = ops.linear(x, w, bias=b, residual=r)
    finally:
        _lib.lib().uvx_debug_gemm_override(0, 0)
        _lib.lib().uvx_debug_gemm_cluster(0, 0)
    assert rel(y, ref) < 1e-3
    assert torch.equal(y, y2)


def test_gemm_row_map(ops):
    M, N, K = 50, 128, 64
    x, w = rnd(M, K, seed=1), rnd(N, K, seed=2)
    out = torch.zeros(80, N, dtype=BF, device="cuda")
    perm = torch.randperm(80)[:M].to(torch.int32)
    perm[7] = -1
    ops.linear(x, w, out=out, row_map=perm.cuda())
    ref = (x.float() @ w.float().T).to(BF)
    for m in range(M):
        if perm[m] >= 0:
            assert torch.equal(out[perm[m]], ref[m].to(out.dtype)) or rel(out[perm[m]], ref[m]) < 1e-3
    untouched = sorted(set(range(80)) - set(int(v) for v in perm if v >= 0))
    assert torch.count_nonzero(out[untouched]) == 0


@pytest.mark.parametrize("stride,Cin,Cout,T,N", [(1, 80, 128, 100, 2), (2, 128, 128, 100, 2), (2, 128, 256, 37, 3),
                                                 (1, 128, 1280, 3000, 1), (2, 1280, 1280, 3000, 1)])
def test_conv_implicit_gemm(ops, stride, Cin, Cout, T, N):
    x = rnd(N, Cin, T, seed=5).float()           # [N, C, T] "mel"
    w = rnd(Cout, Cin, 3, scale=0.05, seed=6)
    b = rnd(Cout, seed=7)
    x_tm = torch.zeros(N, T + 2, Cin, dtype=BF, device="cuda")
    x_tm[:, 1:T + 1] = x.transpose(1, 2).to(BF)
    w_r = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()
    Tout = (T + stride - 1) // stride if stride > 1 else T
    ref = F.gelu(F.conv1d(x.to(BF).float(), w.float(), b.float(), stride=stride, padding=1)).transpose(1, 2)
    assert ref.shape[1] == Tout
    out = torch.zeros(N, Tout + 2, Cout, dtype=BF, device="cuda")
    ops.conv1d_k3(x_tm, w_r, b, stride, out, out_guard=True)
    assert rel(out[:, 1:Tout + 1], ref) < 1e-3
    assert torch.count_nonzero(out[:, 0]) == 0 and torch.count_nonzero(out[:, Tout + 1]) == 0
    pos = rnd(Tout, Cout, seed=8)
    out2 = torch.empty(N, Tout, Cout, dtype=BF, device="cuda")
    ops.conv1d_k3(x_tm, w_r, b, stride, out2, out_guard=False, pos=pos)
    assert rel(out2, ref + pos.float()) < 1e-3


@pytest.mark.parametrize("rows,cols", [(7, 128), (1500, 1280), (201, 4096), (33, 10240)])
def test_norms(ops, rows, cols):
    x, w, b = rnd(rows, cols, seed=1), rnd(cols, seed=2), rnd(cols, seed=3)
    y = ops.layernorm(x, w, b, 1e-5)
    assert rel(y, F.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-5)) < 1e-3
    y = ops.rmsnorm(x, w, 1e-6)
    xf = x.float()
    # LlamaRMSNorm order: the normalised value is rounded to bf16 BEFORE the weight multiply
    ref = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(BF).float()
    assert rel(y, ref) < 1e-3


@pytest.mark.parametrize("T,C", [(50, 384), (1500, 1280), (8, 128), (3, 128)])
def test_stack_rmsnorm(ops, T, C):
    enc, w = rnd(2, T, C, seed=1), rnd(8 * C, seed=2)
    y = ops.stack_rmsnorm(enc, w, 8)
    Tp = (T + 7) // 8 * 8
    st = F.pad(enc.float(), (0, 0, 0, Tp - T)).reshape(2, Tp // 8, 8 * C)
    ref = w.float() * (st * torch.rsqrt(st.pow(2).mean(-1, keepdim=True) + 1e-6)).to(BF).float()
    assert y.shape == ref.shape and rel(y, ref) < 1e-3


def _attn_ref(q, k, v, scale, causal=False, kv_len=None, block=0):
    B, Hq, S, D = q.shape
    Hkv = k.shape[1]
    k = k.repeat_interleave(Hq // Hkv, 1)
    v = v.repeat_interleave(Hq // Hkv, 1)
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    i = torch.arange(S, device=q.device)[:, None]
    j = torch.arange(k.shape[2], device=q.device)[None, :]
    ok = torch.ones(S, k.shape[2], dtype=torch.bool, device=q.device)
    if causal:
        ok &= j <= i
    if block:
        ok &= (j // block) <= (i // block)
    ok = ok[None, None].expand(B, 1, -1, -1).clone()
    if kv_len is not None:
        ok &= (j[None, None] < kv_len.view(B, 1, 1, 1))
    s = s.masked_fill(~ok, float("-inf"))
    return torch.softmax(s, -1) @ v.float()


@pytest.mark.parametrize("B,Hq,Hkv,S,D,causal,block,ragged", [
    (2, 6, 6, 50, 64, False, 0, True), (1, 20, 20, 1500, 64, False, 0, False), (2, 4, 4, 300, 64, False, 100, True),
    (1, 32, 8, 201, 128, True, 0, False), (2, 8, 2, 77, 64, True, 0, False), (3, 4, 2, 130, 128, True, 0, False)])
def test_attention(ops, B, Hq, Hkv, S, D, causal, block, ragged):
    W = (Hq + 2 * Hkv) * D
    qkv = rnd(B * S, W, seed=3)
    kv_len = torch.tensor([S, max(1, S // 3), S - 1][:B], dtype=torch.int32).cuda() if ragged else None
    out = ops.attention_fused_qkv(qkv, B, S, Hq, Hkv, D, D ** -0.5, causal, kv_len, block)
    t = qkv.view(B, S, Hq + 2 * Hkv, D).permute(0, 2, 1, 3)
    ref = _attn_ref(t[:, :Hq], t[:, Hq:Hq + Hkv], t[:, Hq + Hkv:], D ** -0.5, causal, kv_len, block)
    ref = ref.permute(0, 2, 1, 3).reshape(B * S, Hq * D)
    assert rel(out, ref) < 3e-3  # P is rounded to bf16 before PV (as in every flash kernel)


@pytest.mark.parametrize("B,Hq,Hkv,S,D", [(3, 8, 2, 201, 128), (2, 4, 4, 77, 64)])
def test_attention_left_padding(ops, B, Hq, Hkv, S, D):
    """kv_start: keys in the left padding are masked; rows in the padding see nothing and come out as zeros."""
    W = (Hq + 2 * Hkv) * D
    qkv = rnd(B * S, W, seed=5)
    start = torch.tensor([0, 70, S - 3][:B], dtype=torch.int32).cuda()
    out = ops.attention_fused_qkv(qkv, B, S, Hq, Hkv, D, D ** -0.5, True, None, 0, kv_start=start).view(B, S, Hq * D)
    t = qkv.view(B, S, Hq + 2 * Hkv, D).permute(0, 2, 1, 3)
    for b in range(B):
        p0 = int(start[b])
        sub = t[b:b + 1, :, p0:]                       # the unpadded sequence on its own
        ref = _attn_ref(sub[:, :Hq], sub[:, Hq:Hq + Hkv], sub[:, Hq + Hkv:], D ** -0.5, True)
        ref = ref.permute(0, 2, 1, 3).reshape(S - p0, Hq * D)
        assert rel(out[b, p0:], ref) < 3e-3
        assert torch.count_nonzero(out[b, :p0]) == 0


@pytest.mark.parametrize("B,Hq,Hkv,S,causal,ragged", [(1, 32, 8, 201, True, False), (2, 8, 2, 128, True, False), (1, 4, 4, 129, True, False),
                                                       (2, 8, 2, 640, True, True), (1, 4, 1, 300, False, True), (3, 4, 2, 77, True, True)])
def test_attention_llm_tcgen05(ops, B, Hq, Hkv, S, causal, ragged):
    """tcgen05 / TMEM causal GQA attention (head_dim 128) against fp32 math, against the mma.sync kernel, with the log-sum-exp the
    training backward reads, and in the KV-cache form (separate K / V buffers longer than the query block)."""
    from ultravox_b200 import _lib
    D = 128
    W = (Hq + 2 * Hkv) * D
    qkv = rnd(B * S, W, seed=3)
    kv_len = torch.tensor([S, max(1, S // 3), S - 1][:B], dtype=torch.int32).cuda() if ragged else None
    t = qkv.view(B, S, Hq + 2 * Hkv, D).permute(0, 2, 1, 3)
    ref = _attn_ref(t[:, :Hq], t[:, Hq:Hq + Hkv], t[:, Hq + Hkv:], D ** -0.5, causal, kv_len).permute(0, 2, 1, 3).reshape(B * S, Hq * D)
    out = torch.empty(B * S, Hq * D, dtype=BF, device="cuda")
    lse = torch.empty(B * Hq * S, dtype=torch.float32, device="cuda")
    ops.attention_fused_qkv_train(qkv, B, S, Hq, Hkv, D, D ** -0.5, causal, out, lse, kv_len)
    _lib.lib().uvx_debug_attn_tc(0)
    try:
        out_mma = torch.empty_like(out)
        lse_mma = torch.empty_like(lse)
        ops.attention_fused_qkv_train(qkv, B, S, Hq, Hkv, D, D ** -0.5, causal, out_mma, lse_mma, kv_len)
    finally:
        _lib.lib().uvx_debug_attn_tc(-1)
    valid = torch.ones(B, S, dtype=torch.bool, device="cuda")
    assert rel(out, ref) < 3e-3, rel(out, ref)
    assert rel(out, out_mma.float()) < 3e-3
    assert float((lse - lse_mma).abs().max()) < 2e-3
    # KV-cache form: query block of 70 rows at the end of a longer key sequence (shift = Skv - Sq > 0), strided cache buffers
    if causal and S > 80 and not ragged:
        Sq = 70
        kc = torch.zeros(B, S + 9, Hkv, D, dtype=BF, device="cuda")
        vc = torch.zeros_like(kc)
        t4 = qkv.view(B, S, Hq + 2 * Hkv, D)
        kc[:, :S], vc[:, :S] = t4[:, :, Hq:Hq + Hkv], t4[:, :, Hq + Hkv:]
        qb = t4[:, S - Sq:, :Hq].contiguous()                           # [B, Sq, Hq, D]
        o2 = torch.empty(B * Sq, Hq * D, dtype=BF, device="cuda")
        smax = S + 9
        ops.attention(qb.data_ptr(), kc.data_ptr(), vc.data_ptr(), o2, B, Hq, Hkv, Sq, S, D,
                      (Hq * D, Sq * Hq * D, Hkv * D, smax * Hkv * D, Hkv * D, smax * Hkv * D, Hq * D, Sq * Hq * D), D ** -0.5, True)
        assert rel(o2.view(B, Sq, -1), ref.view(B, S, -1)[:, S - Sq:]) < 3e-3


@pytest.mark.parametrize("B,H,S,block,ragged", [(1, 2, 128, 0, False), (1, 3, 256, 0, False), (2, 6, 50, 0, True),
                                                 (1, 20, 1500, 0, False), (2, 4, 300, 100, True), (3, 2, 700, 0, True)])
def test_attention_encoder_tcgen05(ops, B, H, S, block, ragged):
    """tcgen05 / TMEM encoder attention against fp32 math and against the mma.sync kernel."""
    D = 64
    qkv = rnd(B * S, 3 * H * D, seed=3)
    kv_len = torch.tensor([S, max(1, S // 3), S - 1][:B], dtype=torch.int32).cuda() if ragged else None
    out = ops.attention_encoder_tc(qkv, B, S, H, D ** -0.5, kv_len, block)
    torch.cuda.synchronize()
    t = qkv.view(B, S, 3 * H, D).permute(0, 2, 1, 3)
    ref = _attn_ref(t[:, :H], t[:, H:2 * H], t[:, 2 * H:], D ** -0.5, False, kv_len, block)
    ref = ref.permute(0, 2, 1, 3).reshape(B * S, H * D)
    assert rel(out, ref) < 3e-3, rel(out, ref)
    other = ops.attention_fused_qkv(qkv, B, S, H, H, D, D ** -0.5, False, kv_len, block)
    assert rel(out, other.float()) < 3e-3


def test_rope(ops):
    Hq, Hkv, D, S, B = 8, 2, 128, 40, 2
    qkv = rnd(B * S, (Hq + 2 * Hkv) * D, seed=1)
    inv = ops.llama3_inv_freq(D, 500000.0, dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0,
                                                high_freq_factor=4.0, original_max_position_embeddings=8192))
    cos, sin = ops.rope_tables(inv, 64, "cuda")
    ref = qkv.float().view(B, S, Hq + 2 * Hkv, D).clone()
    pos = torch.arange(S, device="cuda").float()
    fr = pos[:, None] * inv.cuda()[None]
    c, s_ = torch.cat([fr, fr], -1).cos()[None, :, None], torch.cat([fr, fr], -1).sin()[None, :, None]
    x = ref[:, :, :Hq + Hkv]
    rot = torch.cat([-x[..., D // 2:], x[..., :D // 2]], -1)
    ref[:, :, :Hq + Hkv] = x * c + rot * s_
    got = ops.rope_(qkv.clone(), Hq, Hkv, D, cos, sin, rows_per_seq=S)
    assert rel(got.view(B, S, -1, D), ref) < 1e-3
    assert torch.equal(got.view(B, S, -1, D)[:, :, Hq + Hkv:], qkv.view(B, S, -1, D)[:, :, Hq + Hkv:])  # v untouched


def test_swiglu(ops):
    x = rnd(77, 512, seed=1)
    a, g = x.float().chunk(2, -1)
    # torch order: silu(gate) is rounded to bf16 before the multiply
    assert rel(ops.swiglu(x, gate_first=False), F.silu(g).to(BF).float() * a) < 1e-3
    assert rel(ops.swiglu(x, gate_first=True), F.silu(a).to(BF).float() * g) < 1e-3


def test_embed_splice_bit_exact(ops):
    B, S, d, V = 3, 40, 256, 1000
    table, audio = rnd(V, d, seed=1), rnd(4, 12, d, seed=2)
    ids = torch.randint(0, V, (B, S), generator=torch.Generator().manual_seed(3)).cuda()
    start = torch.tensor([5, 20, 0, 28], dtype=torch.int64).cuda()
    tlen = torch.tensor([7, 12, 3, 12], dtype=torch.int32).cuda()
    abs_ = torch.tensor([2, 0, 2], dtype=torch.int64).cuda()
    ref = table[ids].clone()
    a = 0
    for b, cnt in enumerate(abs_.tolist()):
        for _ in range(cnt):
            s, n = int(start[a]), int(tlen[a])
            ref[b, s:s + n] = audio[a, :n]
            a += 1
    src = ops.splice_plan(start, tlen, abs_, B, S, audio.shape[1])
    out = ops.embed_splice(ids, table, audio, src)
    assert torch.equal(out, ref)
    assert torch.equal(ops.embed_splice(ids, table, None, None), table[ids])


def test_lm_head_argmax(ops):
    V, d = 5000, 512
    h, w = rnd(3, d, seed=1), rnd(V, d, scale=0.05, seed=2)
    lg = ops.lm_head(h, w)
    ref = h.float() @ w.float().T
    assert rel(lg, ref) < 1e-5
    assert torch.equal(ops.argmax(lg), lg.argmax(-1))
    t = torch.zeros(2, 777, device="cuda")
    t[0, 5] = t[0, 300] = 2.0
    t[1, 776] = 1.0
    assert ops.argmax(t).tolist() == [5, 776]


@pytest.mark.parametrize("n_mels,secs,B", [(80, 1.0, 1), (128, 30.0, 1), (128, 2.5, 3)])
def test_logmel(ops, n_mels, secs, B):
    from oracle import logmel as om
    n = int(16000 * secs)
    waves = [np.random.default_rng(1000 + i).standard_normal(n - 137 * i).astype(np.float32) for i in range(B)]
    t = np.arange(n) / 16000.0
    if B > 1:
        waves[1] = (0.1 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
        waves[1][int(0.75 * n):] = 0
    padded, _ = om.pad_batch(waves)
    ref = om.log_mel(padded, n_mels)                       # float64 oracle
    got, tm = ops.logmel(torch.from_numpy(padded).cuda(), n_mels, want_f32=True, want_tm=True)
    got = got.cpu().numpy()
    assert got.shape == ref.shape
    # fp32 DFT vs float64: values live in [-1.5, 2]; 2e-3 abs covers near-floor bins of tonal input
    assert np.abs(got - ref).max() < 2e-3, np.abs(got - ref).max()
    assert np.sqrt(((got - ref) ** 2).mean()) < 1e-4
    T = padded.shape[1] // 160
    assert torch.equal(tm[:, 1:T + 1].float().cpu(), torch.from_numpy(got).transpose(1, 2).to(BF).float())
    assert torch.count_nonzero(tm[:, 0]) == 0 and torch.count_nonzero(tm[:, T + 1]) == 0
    av = ops.mel_to_timemajor(torch.from_numpy(got).cuda())
    assert torch.equal(av, tm)


# ------------------------------------------------------------------------------------------ decode-loop helpers (generate.cu)
def test_repetition_penalty_matches_hf_processor(ops):
    from transformers.generation.logits_process import RepetitionPenaltyLogitsProcessor
    g = torch.Generator().manual_seed(3)
    B, V, cap = 3, 1000, 40
    logits = torch.randn(B, V, generator=g)
    seq = torch.randint(0, V, (B, cap), generator=g)
    seq[0, 5] = seq[0, 2]                                   # duplicates are penalised once (gather - rescale - scatter)
    seq[1, :7] = 11
    n = 23
    ref = RepetitionPenaltyLogitsProcessor(1.1)(seq[:, :n], logits.clone())
    got = logits.clone().cuda()
    ops.repetition_penalty_(got, seq.cuda(), torch.tensor([n], dtype=torch.int32).cuda(), 1.1,
                            torch.empty(B, cap, dtype=torch.float32, device="cuda"))
    assert torch.equal(got.cpu(), ref)


def test_sample_matches_softmax_distribution_and_top_k(ops):
    g = torch.Generator().manual_seed(5)
    V, n = 64, 40000
    logits = (torch.randn(1, V, generator=g) * 2).cuda()
    big = logits.expand(n, V).contiguous()
    u = torch.rand(n, generator=g).cuda()
    T = 0.7
    picks = ops.sample(big, T, 0, u)
    p = torch.softmax(logits[0].double() / T, -1).cpu()
    hist = torch.bincount(picks.cpu(), minlength=V).double() / n
    assert float((hist - p).abs().max()) < 0.01
    assert torch.equal(picks, ops.sample(big, T, 0, u))                 # deterministic given u
    k = 5
    picks_k = ops.sample(big, T, k, u)
    topk = set(logits[0].topk(k).indices.tolist())
    assert set(picks_k.unique().tolist()) <= topk
    pk = torch.zeros(V, dtype=torch.double)
    idx = logits[0].topk(k).indices.cpu()
    pk[idx] = torch.softmax(logits[0, idx.cuda()].double().cpu() / T, -1)
    hist_k = torch.bincount(picks_k.cpu(), minlength=V).double() / n
    assert float((hist_k - pk).abs().max()) < 0.01
    # inverse CDF ends: u = 0 -> the first entry with mass, u -> 1 -> the last
    ends = ops.sample(logits.expand(2, V).contiguous(), T, 0, torch.tensor([0.0, 0.9999999], device="cuda"))
    assert ends.tolist() == [0, V - 1]
    # vocabulary-sized rows: valid ids inside the top-k set, step-indexed uniforms
    Vb = 128256
    lg = torch.randn(4, Vb, generator=g).cuda()
    uu = torch.rand(3, 4, generator=g).cuda()
    step = torch.tensor([2], dtype=torch.int32).cuda()
    out = ops.sample(lg, 1.0, 50, uu, step)
    top = lg.topk(50, -1).indices
    assert all(int(out[b]) in top[b].tolist() for b in range(4))
    assert torch.equal(out, ops.sample(lg, 1.0, 50, uu[2].contiguous()))


def test_kv_write_and_token_finish(ops):
    B, S, Hq, Hkv, D, smax, past = 2, 5, 4, 2, 64, 16, 3
    qkv = rnd(B * S, (Hq + 2 * Hkv) * D, seed=1)
    kc = torch.zeros(B, smax, Hkv, D, dtype=BF, device="cuda")
    vc = torch.zeros_like(kc)
    ops.kv_write(qkv, kc, vc, B, S, past, Hq, Hkv, D)
    q3 = qkv.view(B, S, -1)
    assert torch.equal(kc[:, past:past + S].reshape(B, S, -1), q3[:, :, Hq * D:(Hq + Hkv) * D])
    assert torch.equal(vc[:, past:past + S].reshape(B, S, -1), q3[:, :, (Hq + Hkv) * D:])
    assert torch.count_nonzero(kc[:, :past]) == 0 and torch.count_nonzero(kc[:, past + S:]) == 0
    tok = torch.tensor([7, 9, 4], dtype=torch.int64, device="cuda")
    done = torch.tensor([0, 1, 0], dtype=torch.int32, device="cuda")
    eos = torch.tensor([4, 100], dtype=torch.int64, device="cuda")
    seq = torch.zeros(3, 10, dtype=torch.int64, device="cuda")
    cur = torch.tensor([6], dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    pos = torch.tensor([5, 5, 5], dtype=torch.int32, device="cuda")
    lens = pos + 1
    alld = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.token_finish(tok, done, eos, 55, seq, cur, step, (pos, lens), alld)
    assert tok.tolist() == [7, 55, 4] and done.tolist() == [0, 1, 1] and seq[:, 6].tolist() == [7, 55, 4]
    assert int(cur) == 7 and int(step) == 1 and pos.tolist() == [6, 6, 6] and lens.tolist() == [7, 7, 7] and int(alld) == 0
    tok.copy_(torch.tensor([100, 1, 2]))
    ops.token_finish(tok, done, eos, 55, seq, cur, step, (pos, lens), alld)
    assert tok.tolist() == [100, 55, 55] and done.tolist() == [1, 1, 1] and int(alld) == 1 and int(cur) == 8


# ------------------------------------------------------------------------------------------ round 2: weight-stream GEMM
@pytest.mark.parametrize("B", [1, 2, 5])
def test_decode_step_fusions_bit_identical(ops, B):
    """uvx_gemv_fused_bf16 (RMSNorm / SwiGLU in the prologue) and uvx_rope_kv_append against the separate kernels they replace in the
    decode step: same roundings, same fp32 summation order -> identical bits."""
    K, N, Fh = 4096, 1024, 2048
    x, w, nw, r = rnd(B, K, seed=1), rnd(N, K, scale=0.03, seed=2), rnd(K, seed=3), rnd(B, N, seed=4)
    want = ops.gemv(ops.rmsnorm(x, nw, 1e-5), w, residual=r)
    assert torch.equal(ops.gemv(x, w, residual=r, norm=(nw, 1e-5)), want)
    gu, w2 = rnd(B, 2 * Fh, seed=5), rnd(N, Fh, scale=0.03, seed=6)
    want = ops.gemv(ops.swiglu(gu, gate_first=True), w2, residual=r)
    assert torch.equal(ops.gemv(gu, w2, residual=r, swiglu=True), want)
    Hq, Hkv, D, smax = 8, 2, 128, 40
    qkv = rnd(B, (Hq + 2 * Hkv) * D, seed=7)
    inv = ops.llama3_inv_freq(D, 500000.0, None)
    cos, sin = ops.rope_tables(inv, 64, "cuda")
    rope_pos = torch.randint(0, 60, (B,), dtype=torch.int32, device="cuda")
    slot = torch.randint(0, smax, (B,), dtype=torch.int32, device="cuda")
    a, b = qkv.clone(), qkv.clone()
    kc1, vc1 = torch.zeros(B, smax, Hkv, D, dtype=BF, device="cuda"), torch.zeros(B, smax, Hkv, D, dtype=BF, device="cuda")
    kc2, vc2 = kc1.clone(), vc1.clone()
    ops.rope_(a, Hq, Hkv, D, cos, sin, rows_per_seq=1, positions=rope_pos)
    ops.kv_append(a, kc1, vc1, slot, Hq, Hkv, D)
    ops.rope_kv_append_(b, Hq, Hkv, D, cos, sin, rope_pos, kc2, vc2, slot)
    assert torch.equal(a, b) and torch.equal(kc1, kc2) and torch.equal(vc1, vc2)


def _tile_ref(w, R, interleave):
    N, K = w.shape
    n_tiles = -(-N // R)
    if interleave == 16:
        F = N // 2
        rows = []
        for t in range(n_tiles):
            for r in range(R):
                qq, j = divmod(r, 32)
                f = t * (R // 2) + qq * 16 + (j & 15)
                rows.append((f if j < 16 else F + f) if f < F else -1)
    elif interleave == 1:
        rows = []
        for t in range(n_tiles):
            for r in range(R):
                qq, j = divmod(r, 32)
                rows.append(t * R + (qq * 16 + j if j < 16 else 64 + qq * 16 + j - 16))
    elif interleave:
        F = N // 2
        rows = []
        for t in range(n_tiles):
            for r in range(R):
                g, j = divmod(r, 16)
                f = t * (R // 2) + g * 8 + (j & 7)
                rows.append((f if j < 8 else F + f) if f < F else -1)
    else:
        rows = [t * R + r if t * R + r < N else -1 for t in range(n_tiles) for r in range(R)]
    idx = torch.tensor(rows, device=w.device)
    src = torch.cat([w, torch.zeros(1, K, dtype=w.dtype, device=w.device)])[idx.clamp_min(-1)]        # -1 -> the zero row
    return src.view(n_tiles, R, K // 64, 64).permute(0, 2, 1, 3).contiguous().view(-1)


@pytest.mark.parametrize("N,K,R,inter", [(512, 256, 128, 0), (1000, 128, 208, 0), (1024, 192, 208, 8), (28672, 256, 208, 8), (6144, 128, 128, 0),
                                         (1024, 192, 128, 16), (28672, 128, 128, 16), (6144, 128, 128, 1)])
def test_tile_weight_layout(ops, N, K, R, inter):
    w = rnd(N, K, seed=3)
    t = ops.TiledWeight(w, R, swiglu=inter in (8, 16), rope_pairs=inter == 1)
    assert torch.equal(t.image, _tile_ref(w, R, inter))


@pytest.mark.parametrize("M", [201, 77, 402])
def test_gemm_tiled_weights_match_row_major(ops, M):
    """Same tile width, same k order -> the tiled image must reproduce the row-major result bit for bit (plain, residual +
    fused RMSNorm through split-K, ragged 208-wide tiles), for every L2 prefetch distance."""
    from ultravox_b200 import _lib
    x = rnd(M, 1024, seed=1)
    for N, R, splits in ((1024, 128, 0), (4096, 128, 0), (1984, 208, 0), (512, 64, 0), (768, 256, 0)):
        w = rnd(N, 1024, scale=0.05, seed=N)
        tw = ops.TiledWeight(w, R)
        ref32 = x.float() @ w.float().T
        outs = []
        for pf in (0, 3, 12):
            _lib.lib().uvx_debug_gemm_pf(pf)
            outs.append(ops.linear_tiled(x, tw))
        _lib.lib().uvx_debug_gemm_pf(-1)
        assert rel(outs[0], ref32) < 1e-3, (N, R)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # residual + fused norm (o_proj / down_proj form) against the row-major path with the same configuration
    N, K = 1024, 4096
    xa, w, r, nw = rnd(M, K, seed=5), rnd(N, K, scale=0.03, seed=6), rnd(M, N, seed=7), rnd(N, seed=8)
    tw = ops.TiledWeight(w, 128)
    h1, n1 = r.clone(), torch.empty(M, N, dtype=BF, device="cuda")
    ops.linear_tiled(xa, tw, residual=h1, out=h1, norm=(nw, 1e-5, n1))
    assert rel(h1, xa.float() @ w.float().T + r.float()) < 1e-3
    assert torch.equal(n1, ops.rmsnorm(h1, nw, 1e-5))


@pytest.mark.parametrize("M,F,K", [(201, 14336, 4096), (201, 992, 512), (64, 512, 256), (300, 2048, 1024)])
def test_gemm_fused_swiglu(ops, M, F, K):
    """act(gate) * up finished in the gate|up GEMM epilogue against GEMM -> bf16 [M, 2F] -> uvx_swiglu and against fp32 math.
    Not bit-identical by construction: the interleaved image puts a feature's gate and up rows into another tile than the
    row-major weight does, the per-tile rotated K start changes the fp32 summation order, and the fused epilogue uses the
    ex2 / rcp approximations for the sigmoid - all far below the two bf16 roundings of the reference's own op order."""
    x, w = rnd(M, K, seed=1), rnd(2 * F, K, scale=0.03, seed=2)
    want = ops.swiglu(ops.linear(x, w), gate_first=True)
    got = ops.linear_tiled(x, ops.TiledWeight(w, 208, swiglu=True), act=ops.ACT_SWIGLU)
    assert got.shape == (M, F)
    assert rel(got, want.float()) < 3e-3
    assert float((got.float() - want.float()).abs().max()) <= 2 ** -6 * float(want.float().abs().max())   # a few bf16 ulps at most
    xf, wf = x.float(), w.float()
    ref = F_silu_mul(xf @ wf[:F].T, xf @ wf[F:].T)
    assert rel(got, ref) < 4e-3          # two bf16 roundings inside, like the reference's op order


def F_silu_mul(g, u):
    return F.silu(g.to(BF).float()).to(BF).float() * u.to(BF).float()


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("M,S,past,with_pos", [(201, 201, 0, False), (402, 201, 3, False), (77, 77, 0, True)])
def test_gemm_fused_rope_bit_exact(ops, tiled, M, S, past, with_pos):
    _rope_case(ops, tiled, M, S, past, with_pos, ws=False)


@pytest.mark.parametrize("M,S,past,with_pos", [(201, 201, 0, False), (77, 77, 0, True), (8, 1, 37, False)])
def test_gemm_ws_fused_rope_bit_exact(ops, ws_on, M, S, past, with_pos):
    _rope_case(ops, True, M, S, past, with_pos, ws=True)


def _rope_case(ops, tiled, M, S, past, with_pos, ws):
    """RoPE in the q|k|v GEMM epilogue (head_dim 128) == GEMM then uvx_rope, bit for bit; v heads untouched."""
    Hq, Hkv, D, K = 8, 2, 128, 1024
    N = (Hq + 2 * Hkv) * D
    x, w = rnd(M, K, seed=1), rnd(N, K, scale=0.03, seed=2)
    inv = ops.llama3_inv_freq(D, 500000.0, dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                                 original_max_position_embeddings=8192))
    cos, sin = ops.rope_tables(inv, 512, "cuda")
    from ultravox_b200 import _lib
    positions = torch.randint(0, 500, (M,), dtype=torch.int32, device="cuda") if with_pos else None
    if ws:
        want = ops.linear(x, w)                              # weight-streaming form on both sides: same units, same summation order
    else:
        _lib.lib().uvx_debug_gemm_override(1128, 1)          # the fused form runs one 128-wide head per tile: same tiling for the reference
        try:
            want = ops.linear(x, w)
        finally:
            _lib.lib().uvx_debug_gemm_override(0, 0)
    ops.rope_(want, Hq, Hkv, D, cos, sin, rows_per_seq=S, pos_offset=past, positions=positions)
    rope = (cos, sin, positions, S, past, (Hq + Hkv) * D)
    if ws:
        got = ops.linear_tiled(x, ops.TiledWeight(w, 128, rope_pairs=True), rope=rope)      # weight-streaming form: pair-permuted image
        assert torch.equal(ops.linear_tiled(x, ops.TiledWeight(w, 128, rope_pairs=True)), ops.linear(x, w))   # the permutation alone moves no bits
    else:
        got = ops.linear_tiled(x, ops.TiledWeight(w, 128), rope=rope) if tiled else ops.linear(x, w, rope=rope)
    assert torch.equal(got, want)


@pytest.mark.parametrize("M,N,K", [(201, 6144, 1024), (201, 28672, 512), (1500, 5120, 640), (77, 192, 200), (300, 1984, 256), (128, 64, 64)])
def test_gemm_tma_store_epilogue_bit_identical(ops, M, N, K):
    """bf16 outputs in plain row order leave through TMA stores (32 x 32 boxes, rows / columns past the matrix clipped by the
    tensor map); same arithmetic as the transposing epilogue -> identical bits, with bias + GELU, ragged 208-wide tiles and
    M / N tails, and nothing written outside the output."""
    from ultravox_b200 import _lib
    x, w, b = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
    res = {}
    for on in (7, 0):                                     # 7: every TMA-store variant (plain, residual, fp32 split-K partials)
        _lib.lib().uvx_debug_gemm_tma_store(on)
        try:
            guard = torch.full((M + 2, N + 64), 7.0, dtype=BF, device="cuda")     # output is a window of a larger buffer
            out = guard[1:M + 1, :N]
            ops.linear(x, w, out=out)
            r = rnd(M, N, seed=4)
            _lib.lib().uvx_debug_gemm_override(0, 3 if K >= 512 else 0)          # split-K partials as well where K allows
            sk = ops.linear(x, w, residual=r)
            _lib.lib().uvx_debug_gemm_override(0, 0)
            res[on] = (out.clone(), ops.linear(x, w, bias=b, act=ops.ACT_GELU), guard, ops.linear(x, w, bias=b, residual=r), sk)
        finally:
            _lib.lib().uvx_debug_gemm_tma_store(-1)
            _lib.lib().uvx_debug_gemm_override(0, 0)
    for i in (0, 1, 3, 4):
        assert torch.equal(res[7][i], res[0][i]), i
    assert rel(res[7][0], x.float() @ w.float().T) < 1e-3
    g = res[7][2]
    assert bool((g[0] == 7).all()) and bool((g[M + 1] == 7).all()) and bool((g[:, N:] == 7).all())


# ------------------------------------------------------------------------------------------ weight-streaming form (gemm_ws.cu)
def _old_kernel(ops, *a, **k):
    return ops.linear(*a, flags=2, **k)


@pytest.fixture
def ws_on():
    """The weight-streaming form is opt-in (UVX_GEMM_WS=1): these tests switch it on through the tuning hook."""
    from ultravox_b200 import _lib
    _lib.lib().uvx_debug_gemm_ws(1, 0, 0)
    yield
    _lib.lib().uvx_debug_gemm_ws(-1, 0, 0)


@pytest.mark.parametrize("M,N,K", [(201, 4096, 4096), (201, 6144, 4096), (201, 28672, 512), (201, 4096, 14336), (1, 4096, 4096), (16, 1024, 256),
                                   (37, 192, 200), (256, 128, 64), (129, 320, 72), (77, 2048, 8192), (5, 16064, 512),
                                   (8, 28672, 512), (2, 19200, 256), (8, 4096, 14336), (32, 28672, 4096)])
def test_gemm_ws_plain_and_epilogues(ops, ws_on, M, N, K):
    """Rows <= 256 run the weight-streaming form (tokens on the UMMA N dimension, stream-K with in-kernel fix-up): fp32 math on
    the same bf16 inputs rounded once (1e-3), agreement with gemm_tc_kernel on the same call, run-to-run identical bits
    (the owner adds the contributors' partial accumulators in CTA order), nothing written outside the output window;
    bias / GELU / residual / in-place residual + RMSNorm epilogues; N and K tails."""
    torch.backends.cuda.matmul.allow_tf32 = False
    x, w, b, r = rnd(M, K, seed=1), rnd(N, K, scale=0.03, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    ref = x.float() @ w.float().T
    guard = torch.full((M + 2, N + 64), 7.0, dtype=BF, device="cuda")
    out = guard[1:M + 1, :N]
    ops.linear(x, w, out=out)
    assert rel(out, ref) < 1e-3
    assert bool((guard[0] == 7).all()) and bool((guard[M + 1] == 7).all()) and bool((guard[:, N:] == 7).all())
    assert torch.equal(out, ops.linear(x, w))                                        # deterministic
    assert rel(out, _old_kernel(ops, x, w).float()) < 2e-3                           # both rounded: two bf16 roundings apart at most
    y = ops.linear(x, w, bias=b, act=ops.ACT_GELU)
    assert rel(y, F.gelu(ref + b.float())) < 1.5e-3
    y = ops.linear(x, w, bias=b, residual=r)
    assert rel(y, ref + b.float() + r.float()) < 1e-3
    if N <= 16384:                                                                   # (uvx_rmsnorm's row limit)
        nw, h, n = rnd(N, seed=5), r.clone(), torch.empty(M, N, dtype=BF, device="cuda")
        ops.linear(x, w, residual=h, out=h, norm=(nw, 1e-5, n))                      # o_proj / down_proj form (in-place residual stream)
        assert rel(h, ref + r.float()) < 1e-3
        assert torch.equal(n, ops.rmsnorm(h, nw, 1e-5))
    if K % 64 == 0:
        assert torch.equal(ops.linear_tiled(x, ops.TiledWeight(w, 128)), out)        # same units, same order: the image only moves bytes


@pytest.mark.parametrize("grid", [148, 97, 32, 5])
def test_gemm_ws_stream_k_any_grid(ops, grid):
    """The unit ranges are cut for whatever grid runs: tiles split over 1..many CTAs, contributors of a tile in the middle of
    another CTA's range - every cut gives the fp32-math result."""
    from ultravox_b200 import _lib
    M, N, K = 201, 1536, 2048
    x, w = rnd(M, K, seed=1), rnd(N, K, scale=0.03, seed=2)
    _lib.lib().uvx_debug_gemm_ws(1, 0, grid)
    try:
        y = ops.linear(x, w)
        y2 = ops.linear(x, w)
    finally:
        _lib.lib().uvx_debug_gemm_ws(-1, 0, 0)
    assert rel(y, x.float() @ w.float().T) < 1e-3 and torch.equal(y, y2)


@pytest.mark.parametrize("M,F,K", [(201, 14336, 4096), (201, 1024, 512), (64, 512, 256), (1, 256, 128), (256, 128, 4096)])
def test_gemm_ws_fused_swiglu(ops, ws_on, M, F, K):
    """act(gate) * up in the weight-streaming epilogue (64 gate | 64 up rows per tile, partner values swapped through shared
    memory) against GEMM -> bf16 [M, 2F] -> uvx_swiglu (a few bf16 ulps: ex2 / rcp sigmoid) and against fp32 math."""
    x, w = rnd(M, K, seed=1), rnd(2 * F, K, scale=0.03, seed=2)
    want = ops.swiglu(ops.linear(x, w), gate_first=True)
    got = ops.linear_tiled(x, ops.TiledWeight(w, 128, swiglu=True), act=ops.ACT_SWIGLU)
    assert got.shape == (M, F)
    assert rel(got, want.float()) < 3e-3
    assert float((got.float() - want.float()).abs().max()) <= 2 ** -6 * float(want.float().abs().max())
    xf, wf = x.float(), w.float()
    assert rel(got, F_silu_mul(xf @ wf[:F].T, xf @ wf[F:].T)) < 4e-3

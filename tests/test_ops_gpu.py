"""Per-op parity of the hot-path kernels THROUGH THE C ABI (mv_gemm, mv_attention_fwd/bwd, mv_layernorm_*, mv_mlm_ce,
mv_adamw) against plain fp32 restatements of the same operator, at the pre-training step's real shapes
(BASELINE.json configs[1] and the joint-length-512 mask sweep of configs[2]).

Tolerances: the fp32 check-mode twins (SIMT) must agree to 1e-4 relative (north_star "fp32 check mode"); the bf16
tcgen05 kernels to 1e-2 relative Frobenius error (north_star "within 1e-2 relative for bf16"); integer outputs
(argmax, correct counts, mask cells) bit-exact.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _L():
    from medvill_b200 import _lib

    return _lib


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_gemm(prec, M, N, K, A, B, Cout, a_mn=0, b_mn=0, c_f32=0, accumulate=0, epi=0, bias=None, resid=None, aux=None, C2=None):
    L = _L()
    d = L.mv_gemm_desc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn = A.data_ptr(), A.stride(0), a_mn
    d.B, d.ldb, d.b_mn = B.data_ptr(), B.stride(0), b_mn
    d.C, d.ldc, d.c_f32, d.accumulate = Cout.data_ptr(), Cout.stride(0), c_f32, accumulate
    if C2 is not None:
        d.C2, d.ldc2 = C2.data_ptr(), C2.stride(0)
    d.epi = epi
    d.bias = bias.data_ptr() if bias is not None else None
    if resid is not None:
        d.resid, d.ldr = resid.data_ptr(), resid.stride(0)
    if aux is not None:
        d.aux, d.ldaux = aux.data_ptr(), aux.stride(0)
    L.check(L.lib().mv_gemm(C.byref(d), prec, L.stream_ptr()), "mv_gemm")
    torch.cuda.synchronize()


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


# (M, N, K): QKV projection, FFN-1, FFN-2 of BERT-base at B=64 / L=436, a ragged MLM-head shape, a tiny one
FWD_SHAPES = [(27904, 2304, 768), (27904, 3072, 768), (27904, 768, 3072), (1283, 30522, 768), (7, 768, 768)]


@pytest.mark.parametrize("prec,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
@pytest.mark.parametrize("M,N,K", FWD_SHAPES)
def test_gemm_forward_bias_gelu(prec, tol, M, N, K):
    """nn.Linear + erf-GELU (models/cxrbert_origin.py:176-181): Y = gelu(X W^T + b), pre-activation kept for backward"""
    L = _L()
    if prec == "fp32" and M * N * K > 3e11 / 4:
        M = 4096          # the SIMT check GEMM is a parity aid, not a throughput kernel
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    X = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(dt)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dt)
    b = torch.randn(N, device=DEV, generator=g)
    ld = (N + 63) // 64 * 64
    prec_id = L.MV_PREC_BF16 if prec == "bf16" else L.MV_PREC_FP32
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        pre = X.float() @ W.float().t() + b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    if N % 32:      # the tied MLM decoder (N = vocab): fp32 logits with a padded row stride, bias only (engine.cu forward)
        Y = torch.zeros(M, ld, device=DEV, dtype=torch.float32)
        run_gemm(prec_id, M, N, K, X, W, Y, c_f32=1, epi=L.EPI_BIAS, bias=b)
        assert rel_err(Y[:, :N], pre) < tol
        return
    Y = torch.zeros(M, ld, device=DEV, dtype=dt)
    P = torch.zeros(M, ld, device=DEV, dtype=dt)
    run_gemm(prec_id, M, N, K, X, W, Y, epi=L.EPI_BIAS_GELU, bias=b, C2=P)
    assert rel_err(P[:, :N].float(), pre) < tol
    assert rel_err(Y[:, :N].float(), gelu_erf(pre)) < tol
    # the pair the engine uses: forward saves gelu'(pre) instead of pre, the backward epilogue multiplies by it
    G = torch.zeros(M, ld, device=DEV, dtype=dt)
    run_gemm(prec_id, M, N, K, X, W, Y, epi=L.EPI_BIAS_GELU_GRAD, bias=b, C2=G)
    dgelu = 0.5 * (1.0 + torch.erf(pre * 0.7071067811865476)) + pre * torch.exp(-0.5 * pre * pre) * 0.3989422804014327
    assert rel_err(Y[:, :N].float(), gelu_erf(pre)) < tol
    assert rel_err(G[:, :N].float(), dgelu) < tol
    dY = (torch.randn(M, K, device=DEV, generator=g) * 0.1).to(dt)          # dX[M,N] = (dY[M,K] . W2[K,N]) * G, W2 read MN-major
    W2 = (torch.randn(K, N, device=DEV, generator=g) * 0.05).to(dt)
    dX = torch.zeros(M, ld, device=DEV, dtype=dt)
    run_gemm(prec_id, M, N, K, dY, W2, dX, a_mn=0, b_mn=1, epi=L.EPI_MUL, aux=G)
    assert rel_err(dX[:, :N].float(), (dY.float() @ W2.float()) * G[:, :N].float()) < tol


@pytest.mark.parametrize("prec,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_gemm_dgrad_and_wgrad_majors(prec, tol):
    """dX = dY . W (W read in place, MN-major B) with the GELU-derivative epilogue; dW += dY^T . X (both MN-major, fp32
    reduce-add into a non-zero gradient arena) — the two backward contractions of every nn.Linear."""
    L = _L()
    P = L.MV_PREC_BF16 if prec == "bf16" else L.MV_PREC_FP32
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    M, N, K = (27904, 768, 3072) if prec == "bf16" else (2048, 768, 3072)   # dY[M,N], W[N,K] -> dX[M,K]
    g = torch.Generator(device=DEV).manual_seed(5)
    dY = (torch.randn(M, N, device=DEV, generator=g) * 0.1).to(dt)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dt)
    X = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(dt)
    pre = torch.randn(M, K, device=DEV, generator=g).to(dt)
    dX = torch.zeros(M, K, device=DEV, dtype=dt)
    run_gemm(P, M, K, N, dY, W, dX, a_mn=0, b_mn=1, epi=L.EPI_DGELU, aux=pre)
    torch.backends.cuda.matmul.allow_tf32 = False
    x = pre.float()
    dgelu = 0.5 * (1.0 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327
    assert rel_err(dX.float(), (dY.float() @ W.float()) * dgelu) < tol
    dW = torch.full((N, K), 0.25, device=DEV, dtype=torch.float32)
    run_gemm(P, N, K, M, dY, X, dW, a_mn=1, b_mn=1, c_f32=1, accumulate=1)
    assert rel_err(dW, 0.25 + dY.float().t() @ X.float()) < tol


def dense_mask(mode, t_len, B, A, Lq):
    L = _L()
    out = torch.empty(B, Lq, Lq, dtype=torch.uint8, device=DEV)
    L.check(L.lib().mv_attn_mask_dump(L.ptr(mode), L.ptr(t_len), B, A, Lq, L.ptr(out), L.stream_ptr()), "mask_dump")
    torch.cuda.synchronize()
    return out.bool()


def attention_reference(qkv, mask, nh):
    """upstream BertSelfAttention with the reference's additive -10000 mask (models/cxrbert_origin.py:82-83), fp32"""
    B, Lq, H3 = qkv.shape
    H = H3 // 3
    q, k, v = [t.view(B, Lq, nh, 64).transpose(1, 2) for t in qkv.float().split(H, dim=2)]
    s = q @ k.transpose(-1, -2) / 8.0 + (1.0 - mask[:, None].float()) * -10000.0
    p = torch.softmax(s, dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(B, Lq, H)
    return ctx, torch.logsumexp(s, dim=-1)


SWEEP = [  # (joint length, A = regions + 2, per-sample modes): configs[1] and the four masks of configs[2] at L = 512
    (436, 182, "bar"), (512, 258, "bidir"), (512, 258, "s2s"), (512, 258, "mixed"), (512, 258, "noncross"), (512, 258, "bar"),
    (77, 11, "mixed"),
]


@pytest.mark.parametrize("prec,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
@pytest.mark.parametrize("Lq,A,kind", SWEEP)
def test_attention_fwd_bwd_all_masks(prec, tol, Lq, A, kind):
    L = _L()
    P = L.MV_PREC_BF16 if prec == "bf16" else L.MV_PREC_FP32
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    B, nh = 6, 12
    H = nh * 64
    rng = np.random.RandomState(Lq + A)
    modes = {"bar": [orc.MODE_BAR] * B, "bidir": [orc.MODE_BIDIR] * B, "s2s": [orc.MODE_S2S] * B, "noncross": [orc.MODE_NONCROSS] * B,
             "mixed": [orc.MODE_S2S if rng.rand() < 0.75 else orc.MODE_BIDIR for _ in range(B)]}[kind]
    T = Lq - A
    t_len = np.concatenate([[1, T], rng.randint(2, T + 1, size=B - 2)]).astype(np.int32)     # shortest and longest text
    mode = torch.tensor(modes, dtype=torch.uint8, device=DEV)
    tl = torch.tensor(t_len, dtype=torch.int32, device=DEV)
    mask = dense_mask(mode, tl, B, A, Lq)
    g = torch.Generator(device=DEV).manual_seed(Lq)
    qkv = torch.randn(B, Lq, 3 * H, device=DEV, generator=g).to(dt)
    dctx = (torch.randn(B, Lq, H, device=DEV, generator=g) * 0.1).to(dt)
    ctx = torch.zeros(B, Lq, H, device=DEV, dtype=dt)
    lse = torch.zeros(B, nh, Lq, device=DEV)
    L.check(L.lib().mv_attention_fwd(B, Lq, nh, A, L.ptr(mode), L.ptr(tl), L.ptr(qkv), L.ptr(ctx), L.ptr(lse), 0.0, 0, 0, P,
                                     L.stream_ptr()), "attention_fwd")
    dqkv = torch.zeros(B, Lq, 3 * H, device=DEV, dtype=dt)
    dq_acc = torch.zeros(B * Lq, H, device=DEV)
    delta = torch.zeros(B, nh, Lq, device=DEV)
    L.check(L.lib().mv_attention_bwd(B, Lq, nh, A, L.ptr(mode), L.ptr(tl), L.ptr(qkv), L.ptr(ctx), L.ptr(lse), L.ptr(dctx), L.ptr(dqkv),
                                     L.ptr(dq_acc), L.ptr(delta), 0.0, 0, 0, P, L.stream_ptr()), "attention_bwd")
    torch.cuda.synchronize()
    qkv_ref = qkv.float().requires_grad_(True)
    ctx_ref, lse_ref = attention_reference(qkv_ref, mask, nh)
    ctx_ref.backward(dctx.float())
    assert rel_err(ctx.float(), ctx_ref.detach()) < tol
    assert float((lse - lse_ref.detach()).abs().max()) < (2e-2 if prec == "bf16" else 1e-4)
    H_ = H
    for name, sl in (("dQ", slice(0, H_)), ("dK", slice(H_, 2 * H_)), ("dV", slice(2 * H_, 3 * H_))):
        assert rel_err(dqkv[..., sl].float(), qkv_ref.grad[..., sl]) < (1.5e-2 if prec == "bf16" else 1e-4), name


def test_attention_dropout_matches_between_forward_and_backward():
    """attention-probability dropout is regenerated from (seed, site, element) in both passes: with the same seed the
    bf16 tcgen05 kernel and the fp32 SIMT twin drop the SAME probabilities, and a different seed drops others."""
    L = _L()
    B, nh, Lq, A = 3, 12, 436, 182
    H = nh * 64
    mode = torch.full((B,), orc.MODE_BAR, dtype=torch.uint8, device=DEV)
    tl = torch.tensor([5, 254, 100], dtype=torch.int32, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv32 = torch.randn(B, Lq, 3 * H, device=DEV, generator=g).to(torch.bfloat16).float()
    outs = {}
    for prec, dt, seed in ((L.MV_PREC_BF16, torch.bfloat16, 77), (L.MV_PREC_FP32, torch.float32, 77), (L.MV_PREC_BF16, torch.bfloat16, 78)):
        qkv = qkv32.to(dt)
        ctx = torch.zeros(B, Lq, H, device=DEV, dtype=dt)
        lse = torch.zeros(B, nh, Lq, device=DEV)
        L.check(L.lib().mv_attention_fwd(B, Lq, nh, A, L.ptr(mode), L.ptr(tl), L.ptr(qkv), L.ptr(ctx), L.ptr(lse), 0.1, seed, 16, prec,
                                         L.stream_ptr()), "attention_fwd")
        torch.cuda.synchronize()
        outs[(prec, seed)] = ctx.float()
    assert rel_err(outs[(L.MV_PREC_BF16, 77)], outs[(L.MV_PREC_FP32, 77)]) < 1e-2
    assert rel_err(outs[(L.MV_PREC_BF16, 78)], outs[(L.MV_PREC_FP32, 77)]) > 5e-2


@pytest.mark.parametrize("prec,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
@pytest.mark.parametrize("rows,eps", [(27904, 1e-12), (1283, 1e-5), (3, 1e-12)])
def test_layernorm_fwd_bwd(prec, tol, rows, eps):
    """nn.LayerNorm(eps=1e-12) of the encoder and the TF-style eps=1e-5 LayerNorm of the MLM head
    (models/cxrbert_origin.py:189-202), forward and backward with fp32 dgamma / dbeta accumulation"""
    L = _L()
    P = L.MV_PREC_BF16 if prec == "bf16" else L.MV_PREC_FP32
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    H = 768
    g = torch.Generator(device=DEV).manual_seed(rows)
    x = (torch.randn(rows, H, device=DEV, generator=g) * 1.5 + 0.3).to(dt)
    dy = torch.randn(rows, H, device=DEV, generator=g).to(dt)
    gamma = torch.rand(H, device=DEV, generator=g) + 0.5
    beta = torch.randn(H, device=DEV, generator=g)
    y = torch.zeros_like(x)
    L.check(L.lib().mv_layernorm_fwd(L.ptr(x), L.ptr(y), L.ptr(gamma), L.ptr(beta), rows, H, eps, P, L.stream_ptr()), "ln_fwd")
    dx = torch.zeros_like(x)
    dgamma = torch.zeros(H, device=DEV)
    dbeta = torch.zeros(H, device=DEV)
    L.check(L.lib().mv_layernorm_bwd(L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), rows, H, eps, P,
                                     L.stream_ptr()), "ln_bwd")
    torch.cuda.synchronize()
    xr = x.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (H,), gr, br, eps)
    yr.backward(dy.double())
    assert rel_err(y.float(), yr.detach()) < tol
    assert rel_err(dx.float(), xr.grad) < tol
    assert rel_err(dgamma, gr.grad) < tol and rel_err(dbeta, br.grad) < tol


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_mlm_cross_entropy_rows(prec):
    """CE(ignore_index=-100) on the labelled rows + argmax accuracy (models/train_origin.py:62,120,138-146):
    loss sum, per-row LSE, FIRST argmax (torch.max semantics) and the correct counter are checked; gradient to tolerance"""
    L = _L()
    P = L.MV_PREC_BF16 if prec == "bf16" else L.MV_PREC_FP32
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    n, V = 1283, 30522
    ld = (V + 63) // 64 * 64
    g = torch.Generator(device=DEV).manual_seed(3)
    logits = torch.zeros(n, ld, device=DEV)
    logits[:, :V] = torch.randn(n, V, device=DEV, generator=g) * 3.0
    labels = torch.randint(0, V, (n,), device=DEV, generator=g)
    logits[5, 100] = logits[5, 200] = 50.0          # tie: the first index wins
    logits[torch.arange(0, n, 7, device=DEV), labels[::7]] = 60.0    # some rows predicted correctly
    dlog = torch.zeros(n, ld, device=DEV, dtype=dt)
    loss = torch.zeros(1, device=DEV)
    correct = torch.zeros(1, device=DEV, dtype=torch.int32)
    row_lse = torch.zeros(n, device=DEV)
    row_arg = torch.zeros(n, device=DEV, dtype=torch.int32)
    gscale = 1.0 / n
    L.check(L.lib().mv_mlm_ce(L.ptr(logits), ld, L.ptr(labels), n, V, L.ptr(dlog), gscale, L.ptr(loss), L.ptr(correct), L.ptr(row_lse),
                              L.ptr(row_arg), P, L.stream_ptr()), "mlm_ce")
    torch.cuda.synchronize()
    lr = logits[:, :V].double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, labels, reduction="sum")
    (ref * gscale).backward()
    assert abs(float(loss) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert torch.equal(row_arg.long(), logits[:, :V].argmax(dim=1))
    assert int(row_arg[5]) == 100
    assert int(correct) == int((logits[:, :V].argmax(dim=1) == labels).sum())
    assert float((row_lse.double() - torch.logsumexp(lr.detach(), dim=1)).abs().max()) < 1e-4
    assert rel_err(dlog[:, :V].float(), lr.grad) < (1e-2 if prec == "bf16" else 1e-5)
    assert float(dlog[:, V:].float().abs().max()) == 0.0


def test_adamw_kernel_matches_hf_formula():
    """HF-3.x AdamW, correct_bias=True, wd=0, eps=1e-6 (models/train_origin.py:60,129-131): three steps vs the oracle's
    restatement; zero_grad fused; the bf16 shadow equals the rounded master weights"""
    L = _L()
    n = 1_000_004          # arena lengths are multiples of 4 (float4 lanes)
    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(n, generator=g) * 0.02
    p = p0.clone().to(DEV)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    shadow = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    ref_p, state = {"w": p0.clone()}, {}
    for step in (1, 2, 3):
        grad = torch.randn(n, generator=g) * 0.01
        gd = grad.clone().to(DEV)
        L.check(L.lib().mv_adamw(L.ptr(p), L.ptr(gd), L.ptr(m), L.ptr(v), L.ptr(shadow), n, 1e-3, 0.9, 0.999, 1e-6, 0.0, step, 1.0, 1,
                                 L.stream_ptr()), "adamw")
        torch.cuda.synchronize()
        assert float(gd.abs().max()) == 0.0
        ref_p.update(orc.adamw_step(ref_p, {"w": grad}, state, lr=1e-3, step=step))
    assert rel_err(p.cpu(), ref_p["w"]) < 1e-6
    assert torch.equal(shadow.cpu(), p.cpu().to(torch.bfloat16))


@pytest.mark.parametrize("p", [0.1, 0.25, 0.037])
def test_dropout_probability_and_scale_are_the_configured_ones(p):
    """Dropout keep decisions are byte compares against a per-group DITHERED threshold (floor(256 p) or floor(256 p) + 1, by
    an 8-bit hash of the group): the drop rate is p to ~1e-4 over 12.6 M elements (plain 8-bit rounding would give 26/256 =
    0.1016 for p = 0.1) and kept values are scaled by exactly 1 / (1 - p), as torch.nn.Dropout does."""
    L = _L()
    M, N, K = 16384, 768, 64
    g = torch.Generator(device=DEV).manual_seed(3)
    X = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
    b = torch.full((N,), 4.0, device=DEV)                       # keeps every pre-dropout value well away from zero
    R = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    Y = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    d = L.mv_gemm_desc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.B, d.ldb = X.data_ptr(), K, W.data_ptr(), K
    d.C, d.ldc = Y.data_ptr(), N
    d.epi, d.bias, d.resid, d.ldr = L.EPI_BIAS_RESID, b.data_ptr(), R.data_ptr(), N
    d.dropout_p, d.dropout_seed, d.dropout_site = p, 12345, 9
    L.check(L.lib().mv_gemm(C.byref(d), L.MV_PREC_BF16, L.stream_ptr()), "mv_gemm")
    torch.cuda.synchronize()
    pre = X.float() @ W.float().t() + b
    dropped = Y == 0
    rate = float(dropped.float().mean())
    sigma = (p * (1 - p) / (M * N)) ** 0.5
    print("dropout p=%.3f: measured drop rate %.6f (binomial sigma %.1e)" % (p, rate, sigma))
    assert abs(rate - p) <= 6 * sigma + 2e-5
    ratio = (Y.float()[~dropped] / pre[~dropped]).mean()
    assert abs(float(ratio) - 1.0 / (1.0 - p)) <= 3e-3 * (1.0 / (1.0 - p))       # bf16 output rounding averages out

"""GPU: the kernels behind the C-ABI at the SDXL shapes of BASELINE.json, checked through
size-independent properties and against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm_ref(A, B, bias=None):
    out = A.float() @ B.float().t()
    return out if bias is None else out + bias.float()


SHAPES = [
    (8192, 1280, 1280),   # attn projections, SDXL d=1280 (M = 8 * 32 * 32)
    (8192, 10240, 1280),  # GEGLU proj
    (8192, 1280, 5120),   # ff.net.2
    (32768, 640, 640),    # attn projections d=640
    (616, 1280, 2048),    # cross-attn k/v (8 * 77 tokens)
    (8, 1280, 1280),      # time_emb_proj
    (1000, 328, 200),     # ragged: nothing is a tile multiple
    (128, 64, 64),
    (1, 8, 8),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_forward_dgrad_wgrad(M, N, K):
    from lycoris_b200.engine import kernels as k

    torch.manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    dY = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    # forward: K-major x K-major
    y = k.gemm(X, W, bias=b)
    ref = _gemm_ref(X, W, b)
    assert float((y.float() - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max()) + 1e-3
    # dgrad: B consumed MN-major, no transposed copy
    dx = k.gemm(dY, W, b_mn=True)
    ref = dY.float() @ W.float()
    assert float((dx.float() - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max()) + 1e-3
    # wgrad: both operands MN-major, fp32 out, split over M
    dw = k.gemm(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32)
    ref = dY.float().t() @ X.float()
    assert float((dw - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-3
    # split count must not change the result beyond fp32 reassociation
    dw1 = k.gemm(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=1)
    assert float((dw - dw1).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-3


def test_gemm_linearity_and_zero():
    from lycoris_b200.engine import kernels as k

    torch.manual_seed(0)
    X = torch.randn(4096, 640, device="cuda", dtype=torch.bfloat16)
    W1 = (torch.randn(640, 640, device="cuda") * 0.03).to(torch.bfloat16)
    Z = torch.zeros_like(W1)
    assert float(k.gemm(X, Z).float().abs().max()) == 0.0
    y1 = k.gemm(X, W1).float()
    y2 = k.gemm(X, (W1.float() * 2).to(torch.bfloat16)).float()  # exact scaling by 2 in bf16
    assert torch.equal(y2, y1 * 2)


def test_gemm_fp16_operands():
    from lycoris_b200.engine import kernels as k

    X = torch.randn(1024, 512, device="cuda", dtype=torch.float16)
    W = (torch.randn(768, 512, device="cuda") * 0.04).to(torch.float16)
    y = k.gemm(X, W)
    ref = _gemm_ref(X, W)
    assert float((y.float() - ref).abs().max()) <= 2.0 ** -10 * float(ref.abs().max()) + 1e-3


def test_gemm_rejects_bad_alignment():
    from lycoris_b200.engine import kernels as k

    X = torch.randn(64, 36, device="cuda", dtype=torch.bfloat16)  # K = 36 is not a multiple of 8
    W = torch.randn(64, 36, device="cuda", dtype=torch.bfloat16)
    assert not k.gemm_supported(X, W)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        k.gemm(X, W)


MERGE_SHAPES = [(1280, 1280), (10240, 1280), (320, 2880), (1280, 11520)]


@pytest.mark.parametrize("N,K", MERGE_SHAPES)
def test_merge_lokr_bit_exact_vs_torch_kron(N, K):
    """LoKr factor 8, full-dim: the kernel's W' equals bf16(W + bf16(kron(w1, w2))) bit for bit."""
    from lycoris_b200.engine import kernels as k

    torch.manual_seed(N + K)
    W = (torch.rand(N, K, device="cuda") * 0.25 - 0.125).to(torch.bfloat16)
    w1 = torch.randn(8, 8, device="cuda") * 0.3
    w2 = torch.randn(N // 8, K // 8, device="cuda") * 0.02
    d = k.make_desc(k.ALGO_LOKR, N, K, factors=[w1, w2], w_dtype=torch.bfloat16, up=8, uq=8, vp=N // 8, vq=K // 8)
    out = k.merge_weight(d, W)
    ref = W + torch.kron(w1, w2).to(torch.bfloat16)
    assert torch.equal(out, ref)
    # zero factors -> identity (idempotence of the merge on an unadapted weight)
    d0 = k.make_desc(k.ALGO_LOKR, N, K, factors=[w1, torch.zeros_like(w2)], w_dtype=torch.bfloat16, up=8, uq=8,
                     vp=N // 8, vq=K // 8)
    assert torch.equal(k.merge_weight(d0, W), W)


@pytest.mark.parametrize("N,K", MERGE_SHAPES[:3])
def test_lokr_factor_grads_vs_autograd(N, K):
    from lycoris_b200.engine import kernels as k

    torch.manual_seed(1)
    w1 = (torch.randn(8, 8, device="cuda") * 0.3).requires_grad_(True)
    w2 = (torch.randn(N // 8, K // 8, device="cuda") * 0.02).requires_grad_(True)
    dW = torch.randn(N, K, device="cuda")
    (torch.kron(w1, w2) * dW).sum().backward()
    d = k.make_desc(k.ALGO_LOKR, N, K, factors=[w1.detach(), w2.detach()], w_dtype=torch.bfloat16, up=8, uq=8,
                    vp=N // 8, vq=K // 8)
    g1, g2 = k.factor_grads(d, dW, None, [w1.shape, w2.shape])
    assert float((g1 - w1.grad).abs().max()) <= 1e-4 * float(w1.grad.abs().max())
    assert float((g2 - w2.grad).abs().max()) <= 1e-4 * float(w2.grad.abs().max())


@pytest.mark.parametrize("algo,r", [("locon", 16), ("locon", 40), ("loha", 32)])
def test_lowrank_merge_and_grads(algo, r):
    from lycoris_b200.engine import kernels as k

    N, K = 1280, 1000  # K not a multiple of the 256-wide tile
    torch.manual_seed(2)
    W = (torch.rand(N, K, device="cuda") * 0.25 - 0.125).to(torch.bfloat16)
    nf = 2 if algo == "locon" else 4
    f = []
    for i in range(nf):
        shape = (N, r) if i % 2 == 0 else (r, K)
        f.append(torch.randn(shape, device="cuda") * 0.1)
    code = k.ALGO_LOCON if algo == "locon" else k.ALGO_LOHA
    d = k.make_desc(code, N, K, factors=f, w_dtype=torch.bfloat16, rank=r, pre_round=1, pre_dtype=torch.bfloat16,
                    m_pre=0.5)
    out = k.merge_weight(d, W)
    fb = [t.to(torch.bfloat16).float() for t in f]
    raw = (fb[0] @ fb[1]).to(torch.bfloat16)
    if algo == "loha":
        raw = raw * (fb[2] @ fb[3]).to(torch.bfloat16)
    ref = W + (raw * 0.5)
    # fp32 accumulation order differs from cuBLAS: allow rare 1-ulp flips of the bf16 product
    mism = float((out != ref).float().mean())
    assert mism < 2e-3, mism
    assert float((out.float() - ref.float()).abs().max()) <= 2.0 ** -7 * 0.25
    fr = [t.to(torch.bfloat16).float().requires_grad_(True) for t in f]
    dd = fr[0] @ fr[1]
    if algo == "loha":
        dd = dd * (fr[2] @ fr[3])
    dW = torch.randn(N, K, device="cuda")
    (dd * 0.5 * dW).sum().backward()
    gs = k.factor_grads(d, dW, None, [t.shape for t in f])
    for g, t in zip(gs, fr):
        assert float((g - t.grad).abs().max()) <= 2e-3 * float(t.grad.abs().max())


@pytest.mark.parametrize("Nb,C,H,W", [(2, 64, 8, 8), (3, 320, 16, 16), (1, 5, 7, 9), (2, 130, 3, 33), (8, 640, 64, 64)])
@pytest.mark.parametrize("src_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_transpose_cast_is_bit_exact(Nb, C, H, W, src_dtype):
    """lyco_transpose_cast == Tensor.to(dtype, memory_format=channels_last), including ragged tiles."""
    from lycoris_b200.engine import kernels as K

    torch.manual_seed(C * H + W)
    x = torch.randn(Nb, C, H, W, device="cuda", dtype=src_dtype)
    dst = torch.bfloat16 if src_dtype != torch.float16 else torch.float16
    before = K._lib.launch_count()
    got = K.as_nhwc(x, dst)
    assert K._lib.launch_count() == before + 1
    want = x.to(dst).contiguous(memory_format=torch.channels_last)
    assert got.dtype == dst and got.shape == x.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)
    assert K.as_nhwc(got, dst) is got  # already NHWC: no pass at all


@pytest.mark.parametrize("Nb,C,O,H,W,stride", [(2, 64, 64, 8, 8, 1), (1, 128, 200, 16, 16, 1), (2, 64, 96, 16, 16, 2)])
def test_conv_fprop_nchw_epilogue_equals_nhwc(Nb, C, O, H, W, stride):
    """The channel-major TMA-store epilogue writes the same numbers as the NHWC one (bit-exact), with bias."""
    from lycoris_b200.engine import kernels as K

    torch.manual_seed(O + H)
    x = torch.randn(Nb, C, H, W, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wk = (torch.randn(O, 9 * C, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(O, device="cuda", dtype=torch.bfloat16)
    a = K.conv2d_fprop(x, wk, bias, 3, 3, (1, 1), stride)
    b = K.conv2d_fprop(x, wk, bias, 3, 3, (1, 1), stride, out_nchw=True)
    assert a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous()
    assert torch.equal(a.contiguous(), b)
    c = K.conv2d_fprop(x, wk, bias, 3, 3, (1, 1), stride, out_nchw=True, out_dtype=torch.float32)
    assert c.dtype == torch.float32 and c.is_contiguous()
    assert torch.equal(c.to(torch.bfloat16), b)  # same accumulators, rounded once instead of not at all
    ref = torch.nn.functional.conv2d(x.float(), wk.view(O, 3, 3, C).permute(0, 3, 1, 2).float(), bias.float(),
                                     stride, 1)
    assert float((b.float() - ref).abs().max()) <= 2e-2 * float(ref.abs().max())


@pytest.mark.parametrize("O,C,R,S", [(64, 64, 3, 3), (72, 200, 3, 3), (1280, 640, 3, 3), (128, 64, 1, 1), (40, 24, 2, 3), (320, 960, 3, 3)])
def test_filter_relayouts_are_bit_exact(O, C, R, S):
    """lyco_filter_relayout == the permute / flip copies it replaces, for all three modes."""
    from lycoris_b200.engine import kernels as K

    torch.manual_seed(O + C)
    w = torch.randn(O, C, R, S, device="cuda").to(torch.bfloat16)
    before = K._lib.launch_count()
    wk = K.filter_relayout(w, K.FILTER_FPROP)
    wd = K.filter_relayout(w, K.FILTER_DGRAD)
    assert torch.equal(wk, w.permute(0, 2, 3, 1).reshape(O, R * S * C))
    assert torch.equal(wd, w.flip(2, 3).permute(1, 2, 3, 0).reshape(C, R * S * O))
    dwk = torch.randn(O, R * S * C, device="cuda")
    dw = K.filter_relayout((dwk, (O, C, R, S)), K.FILTER_WBACK)
    assert dw.shape == (O, C, R, S) and dw.is_contiguous()
    assert torch.equal(dw, dwk.view(O, R, S, C).permute(0, 3, 1, 2).contiguous())
    assert K._lib.launch_count() == before + 3  # all three ran on the engine


@pytest.mark.parametrize("Nb,C,O,H,W,stride", [
    (2, 128, 256, 16, 16, 1),    # CTA-pair tiles in fprop, dgrad and wgrad
    (1, 320, 512, 12, 12, 1),    # N tile padding inside a tap (320 channels), two pair rows
    (3, 640, 384, 8, 8, 1),      # O not a multiple of 256: ragged pair row
    (2, 256, 1280, 8, 8, 2),     # stride 2 wgrad
    (5, 64, 192, 10, 6, 1),      # pixels not a multiple of 128, ragged images
])
def test_conv_kernels_match_fp32_autograd(Nb, C, O, H, W, stride):
    """fprop / dgrad-as-fprop / wgrad on shapes that select the CTA-pair kernels, against fp32 autograd of the
    same bf16 operands (2^-7 relative at tensor scale: fp32 accumulation, one bf16 rounding)."""
    import torch.nn.functional as F

    from lycoris_b200.engine import kernels as K

    torch.manual_seed(C + O)
    x = torch.randn(Nb, C, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(O, C, 3, 3, device="cuda") * 0.05).to(torch.bfloat16)
    xf, wf = x.float().requires_grad_(True), w.float().requires_grad_(True)
    yf = F.conv2d(xf, wf, None, stride, 1)
    dy = torch.randn_like(yf).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    yf.backward(dy.float())

    y = K.conv2d_fprop(x, K.filter_relayout(w, K.FILTER_FPROP), None, 3, 3, (1, 1), stride)
    yf = yf.detach()
    assert float((y.float() - yf).abs().max()) <= 2 ** -7 * float(yf.abs().max())
    dwk = K.conv2d_wgrad(x, dy, 3, 3, (1, 1), stride)
    dw = K.filter_relayout((dwk, (O, C, 3, 3)), K.FILTER_WBACK)
    assert float((dw - wf.grad).abs().max()) <= 1e-3 * float(wf.grad.abs().max())
    if stride == 1:
        dx = K.conv2d_fprop(dy, K.filter_relayout(w, K.FILTER_DGRAD), None, 3, 3, (1, 1), 1)
        assert float((dx.float() - xf.grad).abs().max()) <= 2 ** -7 * float(xf.grad.abs().max())

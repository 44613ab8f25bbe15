"""GPU tests of the round-2 kernels, each against a plain PyTorch expression of the same operation on the device:
structured LoKr factor gradients (lyco_lokr_mix / lyco_lokr_w1grad + two skinny lyco_gemm calls) against the dense
dW' path, lyco_delta_weight (+ norm) against the modules' host formulas, lyco_dora_fwd / lyco_dora_bwd against
autograd through the reference's apply_weight_decompose formula, and merge_to / apply_max_norm on CUDA tensors."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ structured LoKr
@pytest.mark.parametrize("na,nb,nc,transpose", [(8, 8, 160, False), (8, 8, 160, True), (4, 8, 64, False), (8, 2, 24, True),
                                                 (1, 8, 8, False), (3, 5, 40, False)])
def test_lokr_mix_kernel(na, nb, nc, transpose):
    from lycoris_b200.engine import kernels as K

    g = torch.Generator().manual_seed(0)
    M = 1000
    x = torch.randn(M, nb * nc, generator=g).cuda().to(torch.bfloat16)
    w = torch.randn((nb, na) if transpose else (na, nb), generator=g).cuda()
    side = torch.full((777,), 3.0, device="cuda")  # zero-filled on the side by the same launch
    out = K.lokr_mix(x, w, na, nb, nc, transpose, zero=side)
    assert float(side.abs().sum()) == 0.0
    wm = w.t() if transpose else w
    ref = torch.einsum("ab,mbc->mac", wm.float(), x.float().view(M, nb, nc)).reshape(M, na * nc)
    assert out.shape == (M, na * nc) and out.dtype == torch.bfloat16
    assert float((out.float() - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())


@pytest.mark.parametrize("na,nb,nc", [(8, 8, 160), (4, 4, 64), (8, 3, 24), (2, 8, 1280)])
def test_lokr_w1grad_kernel(na, nb, nc):
    from lycoris_b200.engine import kernels as K

    g = torch.Generator().manual_seed(1)
    M = 3000
    P = torch.randn(M, na * nc, generator=g).cuda().to(torch.bfloat16)
    R = torch.randn(M, nb * nc, generator=g).cuda().to(torch.bfloat16)
    out = K.lokr_w1grad(P, R, na, nb, nc, 0.5)
    ref = 0.5 * torch.einsum("mac,mbc->ab", P.double().view(M, na, nc), R.double().view(M, nb, nc))
    assert rel_err(out, ref.float()) <= 1e-4
    pre = torch.zeros(na, nb, device="cuda")  # caller-zeroed buffer: added to, no memset inside
    assert K.lokr_w1grad(P, R, na, nb, nc, 0.5, out=pre) is pre and rel_err(pre, ref.float()) <= 1e-4


SHAPES = [
    # M, N, K  (factor 8)
    (256, 1280, 1280), (8192, 1280, 1280), (8192, 10240, 1280), (8192, 1280, 5120), (616, 1280, 2048), (64, 640, 320),
    (8, 1280, 1280), (1000, 320, 640),
]


@pytest.mark.parametrize("M,N,K", SHAPES, ids=[f"{m}x{n}x{k}" for m, n, k in SHAPES])
@pytest.mark.parametrize("scale", [1.0, 0.37])
def test_structured_lokr_grads_match_dense_path(M, N, K, scale):
    """Same layer, same inputs: factor gradients from the structured contractions vs the dense fp32 dW' + reduction
    pass, and both against an fp64 evaluation of the Kronecker gradient formula."""
    import lycoris_b200 as L
    from lycoris_b200.engine import ops

    torch.manual_seed(0)
    base = nn.Linear(K, N).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    mod = L.LokrModule("t", base, scale, 100000, 1, factor=8).cuda()
    with torch.no_grad():
        mod.lokr_w2.normal_(0, 0.02)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(M, K, generator=g).cuda().to(torch.bfloat16)
    dy = (torch.randn(M, N, generator=g) * 0.1).cuda().to(torch.bfloat16)
    mod.apply_to()

    def run(structured):
        saved = ops._LOKR_STRUCT
        ops._LOKR_STRUCT = structured
        try:
            for p in mod.parameters():
                p.grad = None
            xe = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = base(xe)
            y.backward(dy)
            return y.detach(), xe.grad, mod.lokr_w1.grad.clone(), mod.lokr_w2.grad.clone()
        finally:
            ops._LOKR_STRUCT = saved

    ys, dxs, g1s, g2s = run(True)
    yd, dxd, g1d, g2d = run(False)
    mod.restore()
    assert torch.equal(ys, yd) and torch.equal(dxs, dxd)  # forward / dX do not depend on the gradient path
    # fp64 truth of the factor gradients: dW' = dY^T X, contracted with the other Kronecker block
    dW = dy.double().t() @ x.double()
    up, uq = mod.lokr_w1.shape
    vp, vq = mod.lokr_w2.shape
    dW4 = dW.view(up, vp, uq, vq)
    t1 = torch.einsum("apbq,pq->ab", dW4, mod.lokr_w2.double()) * scale
    t2 = torch.einsum("apbq,ab->pq", dW4, mod.lokr_w1.double()) * scale
    for tag, s_, d_, t_ in (("w1", g1s, g1d, t1), ("w2", g2s, g2d, t2)):
        es, ed = rel_err(s_, t_.float()), rel_err(d_, t_.float())
        assert es <= 1e-2, (tag, "structured vs fp64", es)
        assert ed <= 1e-2, (tag, "dense vs fp64", ed)


@pytest.mark.parametrize("mode,N,K,expect_structured", [("all", 1280, 1280, True), ("auto", 1280, 1280, False),
                                                         ("auto", 5120, 1280, True), ("auto", 1280, 5120, True),
                                                         ("dense", 5120, 1280, False)])
def test_structured_path_selection(mode, N, K, expect_structured):
    """LYCO_LOKR_GRAD: `auto` takes the structured contractions where they pay (max(N,K) >= 4 min(N,K): the feed-forward
    projections) and the dense tensor-core wgrad on square layers; `all` / `dense` force one path."""
    import lycoris_b200 as L
    from lycoris_b200.engine import kernels as K_
    from lycoris_b200.engine import ops

    torch.manual_seed(0)
    base = nn.Linear(K, N).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    mod = L.LokrModule("t", base, 1.0, 100000, 1, factor=8).cuda()
    M = 2048
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    mod.apply_to()
    sink = []
    saved = ops._LOKR_STRUCT
    ops._LOKR_STRUCT = mode
    K_.set_gemm_profiler(sink)
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = base(x)
        y.float().pow(2).mean().backward()
    finally:
        K_.set_gemm_profiler(None)
        ops._LOKR_STRUCT = saved
        mod.restore()
    shapes = sorted((r[3], r[4], r[5]) for r in sink)
    dense_wgrad = (N, K, M)
    skinny = (N // 8, K // 8, M * 8)
    if expect_structured:
        assert dense_wgrad not in shapes and skinny in shapes, shapes
    else:
        assert dense_wgrad in shapes and skinny not in shapes, shapes


# ------------------------------------------------------------------------------------------ LoCon side path
@pytest.mark.parametrize("M,N,K,K2,b_mn", [(1024, 1280, 1280, 16, False), (1024, 1280, 1280, 16, True), (300, 640, 320, 8, False),
                                           (8192, 10240, 1280, 64, False), (616, 1280, 2048, 32, True), (64, 96, 64, 8, False)])
def test_gemm_dual_kernel(M, N, K, K2, b_mn):
    """C = A·Bᵀ + A2·B2ᵀ + bias in one accumulator vs fp32 matmuls of the same bf16 operands."""
    from lycoris_b200.engine import kernels as Kk

    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g).cuda().to(torch.bfloat16)
    a2 = torch.randn(M, K2, generator=g).cuda().to(torch.bfloat16)
    b = (torch.randn(N, K, generator=g) * 0.05).cuda().to(torch.bfloat16)
    b2 = (torch.randn(N, K2, generator=g) * 0.05).cuda().to(torch.bfloat16)
    bias = None if b_mn else torch.randn(N, generator=g).cuda().to(torch.bfloat16)
    ref = a.float() @ b.float().t() + a2.float() @ b2.float().t()
    if bias is not None:
        ref = ref + bias.float()
    if b_mn:
        out = Kk.gemm_dual(a, b.t().contiguous(), a2, b2.t().contiguous(), b_mn=True)
    else:
        out = Kk.gemm_dual(a, b, a2, b2, bias=bias)
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert float((out.float() - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())


@pytest.mark.parametrize("algo", ["locon", "dylora"])
@pytest.mark.parametrize("M,N,K,r", [(2048, 1280, 1280, 16), (8192, 10240, 1280, 16), (616, 1280, 2048, 32), (256, 640, 320, 8)])
def test_locon_side_path_matches_merged_path_and_fp64(algo, M, N, K, r):
    """LoCon / DyLoRA on nn.Linear: y = x·Wᵀ + (x·downᵀ)(s·up)ᵀ in one accumulator, no W', no dense dW' — against the
    round-1 merged-weight path on the same layer and against an fp64 evaluation of the same formula."""
    import random

    import lycoris_b200 as L
    from lycoris_b200.engine import kernels as Kk
    from lycoris_b200.engine import ops

    torch.manual_seed(0)
    base = nn.Linear(K, N).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    if algo == "locon":
        mod = L.LoConModule("t", base, 0.8, r, r / 2).cuda()
        with torch.no_grad():
            mod.lora_up.weight.normal_(0, 0.05)
    else:
        mod = L.DyLoraModule("t", base, 0.8, r, r / 2, block_size=r // 2).cuda()
        with torch.no_grad():
            for u in mod.up_list:
                u.normal_(0, 0.05)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(M, K, generator=gen).cuda().to(torch.bfloat16)
    dy = (torch.randn(M, N, generator=gen) * 0.1).cuda().to(torch.bfloat16)
    mod.apply_to()
    mod.train()

    def run(side):
        saved = ops._LOCON_SIDE
        ops._LOCON_SIDE = side
        sink = []
        Kk.set_gemm_profiler(sink)
        try:
            for p in mod.parameters():
                p.grad = None
            random.seed(5)  # DyLoRA: the same block on both paths (block_count = 2: rank r/2 or r)
            xe = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = base(xe)
            y.backward(dy)
            grads = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
            return y.detach(), xe.grad, grads, sorted((r_[3], r_[4], r_[5]) for r_ in sink)
        finally:
            Kk.set_gemm_profiler(None)
            ops._LOCON_SIDE = saved

    ys, dxs, gs, shapes_s = run(True)
    ym, dxm, gm, shapes_m = run(False)
    mod.restore()
    live_rank = [s_ for s_ in shapes_s if s_[1] <= r and s_[0] == M]  # T = x·downᵀ : (M, rank, K)
    if live_rank and live_rank[0][1] % 8 == 0:
        assert (N, K, M) not in shapes_s, "dense dW' contraction launched on the side path"
        assert (N, K, M) in shapes_m
    assert float((ys.float() - ym.float()).abs().max()) <= 2.0 ** -6 * float(ym.float().abs().max())
    assert float((dxs.float() - dxm.float()).abs().max()) <= 2.0 ** -6 * float(dxm.float().abs().max())
    assert set(gs) == set(gm)
    for k_ in gm:
        assert rel_err(gs[k_], gm[k_]) <= 3e-2, (k_, rel_err(gs[k_], gm[k_]))


@pytest.mark.parametrize("C,O,k,stride,side,r", [(64, 128, 3, 1, 16, 8), (64, 64, 3, 2, 16, 8), (320, 320, 3, 1, 64, 8),
                                                 (128, 64, 1, 1, 16, 16), (1280, 1280, 3, 1, 32, 8)])
def test_locon_conv_skinny_factor_gradients_match_dense_path(C, O, k, stride, side, r):
    """LoCon on an engine-run convolution: g_up / g_down from T = conv(X, down), U = dY.(s up), dY^T T and wgrad(X, U)
    (no dense dW' = wgrad(X, dY)) against the merged-weight path's dense dW' + reduction on the same layer."""
    import lycoris_b200 as L
    from lycoris_b200.engine import ops

    torch.manual_seed(0)
    base = nn.Conv2d(C, O, k, stride, k // 2).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    mod = L.LoConModule("t", base, 0.7, r, r / 2).cuda()
    with torch.no_grad():
        mod.lora_up.weight.normal_(0, 0.05)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, C, side, side, generator=gen).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    mod.apply_to()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        yshape = base(x).shape
    dy = (torch.randn(yshape, generator=gen) * 0.1).cuda().to(torch.bfloat16)

    def run(side_path):
        saved = ops._LOCON_SIDE
        ops._LOCON_SIDE = side_path
        try:
            for p in mod.parameters():
                p.grad = None
            xe = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = base(xe)
            y.backward(dy)
            return y.detach(), xe.grad, mod.lora_up.weight.grad.clone(), mod.lora_down.weight.grad.clone()
        finally:
            ops._LOCON_SIDE = saved

    ys, dxs, gus, gds = run(True)
    yd, dxd, gud, gdd = run(False)
    mod.restore()
    assert torch.equal(ys, yd) and torch.equal(dxs, dxd)  # forward and dX still use the merged weight
    assert gus.shape == gud.shape and gds.shape == gdd.shape
    assert rel_err(gus, gud) <= 2e-2, ("up", rel_err(gus, gud))
    assert rel_err(gds, gdd) <= 2e-2, ("down", rel_err(gds, gdd))


# ------------------------------------------------------------------------------------------ LoHa fused tile kernel
@pytest.mark.parametrize("N,K,r", [(1280, 1280, 32), (10240, 1280, 32), (1280, 11520, 16), (200, 72, 8), (320, 2880, 64),
                                   (96, 64, 8)])
def test_hada_kernels_are_bit_identical_to_the_split_sequence(N, K, r):
    """lyco_hada forms (w1a·w1b) ⊙ (w2a·w2b) per tile in TMEM and merges / builds the gradient operands in the epilogue;
    the round-1 sequence wrote both products with lyco_gemm and read them back in merge_raw / grad_prep.  Same MMA, same
    rounding points -> the results must be BIT-identical, ragged tile edges included."""
    from lycoris_b200.engine import kernels as Kk

    g = torch.Generator().manual_seed(0)
    f = [(torch.randn(N, r, generator=g) * 0.2).cuda().to(torch.bfloat16), (torch.randn(r, K, generator=g) * 0.2).cuda().to(torch.bfloat16),
         (torch.randn(N, r, generator=g) * 0.2).cuda().to(torch.bfloat16), (torch.randn(r, K, generator=g) * 0.2).cuda().to(torch.bfloat16)]
    W = (torch.randn(N, K, generator=g) * 0.05).cuda().to(torch.bfloat16)
    m = (0.5, 1.0, 0.75)
    fused = Kk.hada_merge(f, W, *m)
    P1 = Kk.gemm(f[0], f[1], b_mn=True)
    P2 = Kk.gemm(f[2], f[3], b_mn=True)
    desc = Kk.make_desc(Kk.ALGO_RAW, N, K, factors=[P1, P2], w_dtype=W.dtype, pre_round=1, pre_dtype=torch.bfloat16,
                        m_pre=m[0], m_post1=m[1], m_post2=m[2])
    split = Kk.merge_weight(desc, W)
    assert torch.equal(fused, split), float((fused.float() - split.float()).abs().max())
    ref = (W.float() + (f[0].float() @ f[1].float()) * (f[2].float() @ f[3].float()) * (m[0] * m[1] * m[2]))
    assert float((fused.float() - ref).abs().max()) <= 2.0 ** -6 * float(ref.abs().max())
    dW = torch.randn(N, K, generator=g).cuda()
    G1, G2 = Kk.hada_grad_operands(f, dW, 0.3)
    assert torch.equal(G1, Kk.grad_prep(dW, P2, 0.3, torch.bfloat16))
    assert torch.equal(G2, Kk.grad_prep(dW, P1, 0.3, torch.bfloat16))


def test_loha_layer_fused_vs_split_paths_agree():
    import lycoris_b200 as L
    from lycoris_b200.engine import ops

    torch.manual_seed(0)
    base = nn.Linear(1280, 2560).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    mod = L.LohaModule("t", base, 0.9, 32, 16).cuda()
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_((torch.randn(p.shape, generator=gen) * 0.1).to(p))
    x = torch.randn(4096, 1280, generator=gen).cuda().to(torch.bfloat16)
    dy = (torch.randn(4096, 2560, generator=gen) * 0.1).cuda().to(torch.bfloat16)
    mod.apply_to()

    def run(fused):
        saved = ops._LOHA_FUSED
        ops._LOHA_FUSED = fused
        try:
            for p in mod.parameters():
                p.grad = None
            xe = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = base(xe)
            y.backward(dy)
            return y.detach(), xe.grad, {k: p.grad.clone() for k, p in mod.named_parameters()}
        finally:
            ops._LOHA_FUSED = saved

    yf, dxf, gf = run(True)
    ys, dxs, gs = run(False)
    mod.restore()
    assert torch.equal(yf, ys) and torch.equal(dxf, dxs)
    for k in gs:
        assert rel_err(gf[k], gs[k]) <= 1e-3, (k, rel_err(gf[k], gs[k]))  # split-K atomics order only


# ------------------------------------------------------------------------------------------ delta weight
def _mods(dtype):
    import lycoris_b200 as L

    torch.manual_seed(0)
    lin = nn.Linear(320, 640).cuda()
    conv = nn.Conv2d(64, 128, 3, padding=1).cuda()
    out = []
    for base in (lin, conv):
        out.append(L.LoConModule("a", base, 1.0, 8, 4))
        out.append(L.LohaModule("b", base, 1.0, 8, 4))
        out.append(L.LokrModule("c", base, 1.0, 100000, 1, factor=8))
        out.append(L.LokrModule("d", base, 1.0, 4, 2, factor=4))
    g = torch.Generator().manual_seed(1)
    for m in out:
        m.cuda().to(dtype)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_((torch.randn(p.shape, generator=g) * 0.1).to(p))
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_delta_weight_kernel_matches_host_formulas(dtype):
    """get_diff_weight on CUDA parameters runs lyco_delta_weight; the host formula (what the reference computes) is
    evaluated by forcing the PyTorch path on the same module."""
    from lycoris_b200.engine import _lib

    for mod in _mods(dtype):
        mod.eval()
        before = _lib.launch_count()
        d_eng, _ = mod.get_diff_weight(0.7, mod.shape)
        assert _lib.launch_count() == before + 1, type(mod).__name__
        saved = mod._delta_via_engine
        mod._delta_via_engine = lambda *a, **k: None
        d_host, _ = mod.get_diff_weight(0.7, mod.shape)
        mod._delta_via_engine = saved
        assert d_eng.shape == d_host.shape and d_eng.dtype == d_host.dtype
        tol = 1e-5 if dtype == torch.float32 else 2.0 ** -7
        err = float((d_eng.float() - d_host.float()).abs().max()) / float(d_host.float().abs().max())
        assert err <= tol, (type(mod).__name__, mod.module_type, err)


def test_apply_max_norm_and_merge_on_cuda_match_cpu():
    """apply_max_norm reduces ||dW||_F inside the delta kernel; merge_to adds the kernel's dW.  Same modules on the
    CPU (host PyTorch path, pinned against the reference in tests/test_weights_side.py) give the same numbers."""
    for mod in _mods(torch.float32):
        cpu = copy.deepcopy(mod).cpu()
        cpu.org_module = [copy.deepcopy(mod.org_module[0]).cpu()]
        w0 = mod.org_module[0].weight.detach().clone()
        s_gpu, n_gpu = mod.apply_max_norm(0.05)
        s_cpu, n_cpu = cpu.apply_max_norm(0.05)
        assert bool(s_gpu) == bool(s_cpu)
        assert abs(float(n_gpu) - float(n_cpu)) <= 1e-4 * abs(float(n_cpu)), (type(mod).__name__, float(n_gpu), float(n_cpu))
        for (k, a), (_, b) in zip(mod.named_parameters(), cpu.named_parameters()):
            assert torch.allclose(a.cpu(), b, rtol=1e-4, atol=1e-6), (type(mod).__name__, k)
        mod.merge_to(0.5)
        cpu.merge_to(0.5)
        assert not torch.equal(mod.org_module[0].weight, w0)
        a, b = mod.org_module[0].weight.detach().cpu(), cpu.org_module[0].weight.detach()
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), type(mod).__name__


# ------------------------------------------------------------------------------------------ DoRA
def _dora_ref(Wm16, g, on_out, mult, eps, shape):
    """The reference's apply_weight_decompose (locon.py:239-260) as plain PyTorch on the device."""
    w = Wm16.view(shape).float()
    ones = [1] * (w.dim() - 1)
    if on_out:
        norm = w.reshape(w.shape[0], -1).norm(dim=1).reshape(w.shape[0], *ones) + eps
    else:
        norm = w.transpose(0, 1).reshape(w.shape[1], -1).norm(dim=1, keepdim=True).reshape(w.shape[1], *ones).transpose(0, 1) + eps
    scale = g / norm
    if mult != 1:
        scale = mult * (scale - 1) + 1
    return w * scale


@pytest.mark.parametrize("shape,on_out", [((640, 320), True), ((640, 320), False), ((128, 64, 3, 3), True),
                                          ((128, 64, 3, 3), False), ((1280, 1280), True)])
@pytest.mark.parametrize("mult", [1.0, 0.6])
def test_dora_kernels_match_autograd(shape, on_out, mult):
    from lycoris_b200.engine import kernels as K

    gen = torch.Generator().manual_seed(0)
    N = shape[0]
    taps = 1
    for d in shape[2:]:
        taps *= d
    Wm = (torch.randn(shape, generator=gen) * 0.05).cuda().to(torch.bfloat16)
    groups = N if on_out else shape[1]
    gshape = (N, *[1] * (len(shape) - 1)) if on_out else (1, shape[1], *[1] * (len(shape) - 2))
    g = (torch.rand(groups, generator=gen) + 0.5).cuda()
    eps = float(torch.finfo(torch.float32).eps)
    out, sumsq = K.dora_fwd(Wm.view(N, -1), g, on_out, taps, mult, eps)
    gp = g.view(gshape).clone().requires_grad_(True)
    wl = Wm.float().clone().requires_grad_(True)
    ref = _dora_ref(wl, gp, on_out, mult, eps, shape)
    assert float((out.view(shape).float() - ref.detach()).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())
    dW = torch.randn(shape, generator=gen).cuda()
    ref.backward(dW)
    dWk = dW.clone().view(N, -1).contiguous()
    gs = K.dora_bwd(dWk, Wm.view(N, -1), g, sumsq, on_out, taps, mult, eps)
    assert rel_err(gs, gp.grad.reshape(-1)) <= 1e-4
    assert rel_err(dWk.view(shape), wl.grad) <= 1e-4


@pytest.mark.parametrize("algo", ["locon", "loha", "lokr"])
@pytest.mark.parametrize("on_out", [True, False])
def test_dora_layer_runs_on_the_kernels_and_matches_assembled_path(algo, on_out):
    """dora_wd layers now go merge kernel -> lyco_dora_fwd -> contraction (and lyco_dora_bwd in backward); the
    round-1 path (W' assembled by PyTorch ops + autograd) is kept as the comparison."""
    import lycoris_b200 as L
    from lycoris_b200.engine import _lib

    torch.manual_seed(0)
    base = nn.Conv2d(64, 128, 3, padding=1).cuda().to(torch.bfloat16)
    base.requires_grad_(False)
    kw = dict(weight_decompose=True, wd_on_out=on_out)
    if algo == "locon":
        mod = L.LoConModule("t", base, 0.8, 8, 4, **kw)
    elif algo == "loha":
        mod = L.LohaModule("t", base, 0.8, 8, 4, **kw)
    else:
        mod = L.LokrModule("t", base, 0.8, 100000, 1, factor=8, **kw)
    mod = mod.cuda()
    gen = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for n_, p in mod.named_parameters():
            if float(p.abs().sum()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=gen) * 0.05).to(p))
    x = torch.randn(2, 64, 16, 16, generator=gen).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(2, 128, 16, 16, generator=gen).cuda().to(torch.bfloat16)
    mod.apply_to()

    def run(native):
        for p in mod.parameters():
            p.grad = None
        saved = mod._native_spec
        if not native:
            mod._native_spec = lambda: None
        try:
            xe = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = base(xe)
            y.backward(dy)
        finally:
            mod._native_spec = saved
        return y.detach(), xe.grad, {k: p.grad.clone() for k, p in mod.named_parameters()}

    before = _lib.launch_count()
    y1, dx1, g1 = run(True)
    n_native = _lib.launch_count() - before
    y0, dx0, g0 = run(False)
    mod.restore()
    assert n_native >= 8
    assert float((y1.float() - y0.float()).abs().max()) <= 2.0 ** -6 * float(y0.float().abs().max())
    assert float((dx1.float() - dx0.float()).abs().max()) <= 2.0 ** -6 * float(dx0.float().abs().max())
    assert set(g1) == set(g0) and "dora_scale" in g1
    for k in g0:
        assert rel_err(g1[k], g0[k]) <= 3e-2, (k, rel_err(g1[k], g0[k]))

"""The tcgen05 3xTF32 GEMM building block in isolation (GPU only): K-major and MN-major A operands, ragged
shapes, split reductions, the extra-column output -- against a float64 product."""
import ctypes as C

import numpy as np
import pytest

from drecpy_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    import torch
    lib = _lib.load()
    h = _lib.vp()
    _lib.check(lib.drb_ctx_create(0, C.byref(h)))
    _lib.check(lib.drb_ctx_set_stream(h, _lib.vp(torch.cuda.current_stream().cuda_stream)))
    yield h
    lib.drb_ctx_destroy(h)


def _split(ctx, x, transpose=False, ones_row=-1, t_rows=None):
    import torch
    lib = _lib.load()
    rows, ld = x.shape
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    t_hi = t_lo = None
    ldt = 0
    if transpose:
        ldt = (rows + 3) // 4 * 4
        t_hi = torch.zeros((t_rows or ld, ldt), device='cuda')
        t_lo = torch.zeros_like(t_hi)
    _lib.check(lib.drb_debug_split_tf32(ctx, _lib.t_ptr(x), rows, ld, ld, _lib.t_ptr(hi), _lib.t_ptr(lo),
                                        _lib.t_ptr(t_hi), _lib.t_ptr(t_lo), ldt, ones_row))
    return hi, lo, t_hi, t_lo


def _gemm(ctx, a_hi, a_lo, b_hi, b_lo, a_mn, M, N, K, splits, n_store, extra_idx=-1):
    import torch
    lib = _lib.load()
    ldc = (n_store + 3) // 4 * 4
    Cm = torch.full((splits, M, ldc), float('nan'), device='cuda')
    extra = torch.full((M,), float('nan'), device='cuda') if extra_idx >= 0 else None
    _lib.check(lib.drb_debug_umma_gemm(ctx, _lib.t_ptr(a_hi), _lib.t_ptr(a_lo), a_hi.shape[1], _lib.t_ptr(b_hi),
                                       _lib.t_ptr(b_lo), b_hi.shape[1], b_hi.shape[0], int(a_mn), M, N, K, splits,
                                       _lib.t_ptr(Cm), ldc, n_store, _lib.t_ptr(extra), extra_idx))
    torch.cuda.synchronize()
    return Cm.sum(0), extra


def _report(got, want, tag):
    err = (got.double() - want).abs()
    scale = want.abs().max().item()
    bad = (err > 1e-5 * scale).nonzero()
    msg = f'{tag}: max err {err.max().item():.3e} (scale {scale:.3e}), {len(bad)} bad of {err.numel()}'
    if len(bad):
        msg += f'; first bad {bad[:6].tolist()} got {got[tuple(bad[0])].item():.6f} want {want[tuple(bad[0])].item():.6f}'
        rows = sorted(set(bad[:, 0].tolist()))[:20]
        cols = sorted(set(bad[:, 1].tolist()))[:20]
        msg += f'; bad rows {rows} cols {cols}'
    return err.max().item() <= 1e-5 * scale, msg


@pytest.mark.parametrize('M,N,K,splits', [(128, 64, 32, 1), (128, 128, 64, 1), (300, 208, 200, 1), (257, 56, 1000, 3),
                                          (64, 256, 100, 1), (4096, 208, 2048, 4)])
def test_umma_kmajor(ctx, M, N, K, splits):
    import torch
    g = torch.Generator(device='cuda').manual_seed(1)
    Kp = (K + 3) // 4 * 4
    A = torch.zeros((M, Kp), device='cuda'); A[:, :K] = torch.randn((M, K), device='cuda', generator=g)
    B = torch.zeros((N, Kp), device='cuda'); B[:, :K] = torch.randn((N, K), device='cuda', generator=g)
    a_hi, a_lo, _, _ = _split(ctx, A)
    b_hi, b_lo, _, _ = _split(ctx, B)
    assert torch.equal(a_hi + a_lo, A)
    n_store = N // 4 * 4
    got, _ = _gemm(ctx, a_hi, a_lo, b_hi, b_lo, False, M, N, K, splits, n_store)
    want = A.double() @ B.double().t()
    ok, msg = _report(got[:, :n_store], want[:, :n_store], f'kmajor M{M} N{N} K{K} s{splits}')
    assert ok, msg


@pytest.mark.parametrize('M,N,K', [(128, 64, 32), (128, 64, 64), (1682, 64, 64), (777, 208, 300), (2049, 80, 129),
                                   (26744, 208, 512)])
def test_umma_mnmajor_with_ones_column(ctx, M, N, K):
    """dW'^T = dz^T h shape: A given as G[k][m] (m contiguous), B = transposed activations with a ones row."""
    import torch
    g = torch.Generator(device='cuda').manual_seed(2)
    Mp = (M + 3) // 4 * 4
    hidden = N - 3                                  # ones feature lives at column `hidden` of the product
    G = torch.zeros((K, Mp), device='cuda'); G[:, :M] = torch.randn((K, M), device='cuda', generator=g)
    ld = (hidden + 3) // 4 * 4
    H = torch.zeros((K, ld), device='cuda'); H[:, :hidden] = torch.randn((K, hidden), device='cuda', generator=g)
    g_hi, g_lo, _, _ = _split(ctx, G)
    _, _, ht_hi, ht_lo = _split(ctx, H, transpose=True, ones_row=hidden, t_rows=N)
    assert torch.all(ht_hi[hidden, :K] == 1) and torch.all(ht_lo[hidden] == 0)
    got, extra = _gemm(ctx, g_hi, g_lo, ht_hi, ht_lo, True, M, N, K, 1, ld, extra_idx=hidden)
    want = G[:, :M].double().t() @ H.double()
    ok, msg = _report(got[:, :hidden], want[:, :hidden], f'mnmajor M{M} N{N} K{K}')
    assert ok, msg
    ok, msg = _report(extra[:, None], G[:, :M].double().sum(0)[:, None], 'ones column')
    assert ok, msg


# ------------------------------------------------------------------------------------------------ 3xFP16 building blocks
def _split_f16(ctx, x, alpha, transpose=False, ones_row=-1, t_rows=None):
    import torch
    lib = _lib.load()
    rows, ld = x.shape
    ldh = (ld + 7) // 8 * 8
    hi = torch.zeros((rows, ldh), dtype=torch.float16, device='cuda')
    lo = torch.zeros_like(hi)
    t_hi = t_lo = None
    ldt = 0
    if transpose:
        ldt = (rows + 7) // 8 * 8
        t_hi = torch.zeros((t_rows or ld, ldt), dtype=torch.float16, device='cuda')
        t_lo = torch.zeros_like(t_hi)
    _lib.check(lib.drb_debug_split_f16(ctx, _lib.t_ptr(x), rows, ld, ld, alpha, _lib.t_ptr(hi), _lib.t_ptr(lo), ldh,
                                       _lib.t_ptr(t_hi), _lib.t_ptr(t_lo), ldt, ones_row))
    return hi, lo, t_hi, t_lo


@pytest.mark.parametrize('M,N,K,splits', [(128, 64, 32, 1), (128, 128, 64, 1), (300, 208, 200, 1), (257, 56, 1000, 3),
                                          (64, 256, 100, 1), (4096, 208, 2048, 4), (26744, 208, 512, 2)])
def test_umma_f16_kmajor(ctx, M, N, K, splits):
    """C = A B^T from fp16 hi/lo operands scaled into the fp16 range: within 1e-5 of the float64 product, i.e. the same
    accuracy as the 3xTF32 form.  Operand magnitudes span what the CDAE step feeds (weights ~1e-2, activations in (0,1),
    gradients with a wide dynamic range)."""
    import torch
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N)
    a = torch.rand((M, K), device='cuda', generator=g) * 0.999                       # activations
    b = (torch.rand((N, K), device='cuda', generator=g) - 0.5) * 0.03               # weights
    b = b * torch.pow(10.0, -3 * torch.rand((N, K), device='cuda', generator=g))     # with a wide dynamic range
    alpha_a = 32768.0
    alpha_b = float(2.0 ** (14 - np.floor(np.log2(b.abs().max().item()))))
    a_hi, a_lo, _, _ = _split_f16(ctx, a, alpha_a)
    b_hi, b_lo, _, _ = _split_f16(ctx, b, alpha_b)
    assert float(a_hi.float().abs().max()) <= 32768 and float(b_hi.float().abs().max()) < 32768
    # the split itself keeps ~22 bits of every element that matters
    rec = (a_hi.double() + a_lo.double())[:, :K] / alpha_a
    assert float((rec - a.double()).abs().max()) <= 2.0 ** -22
    lib = _lib.load()
    ldc = (N + 3) // 4 * 4
    Cm = torch.full((splits, M, ldc), float('nan'), device='cuda')
    _lib.check(lib.drb_debug_umma_gemm_f16(ctx, _lib.t_ptr(a_hi), _lib.t_ptr(a_lo), a_hi.shape[1], _lib.t_ptr(b_hi),
                                           _lib.t_ptr(b_lo), b_hi.shape[1], N, 0, M, N, K, splits,
                                           1.0 / (alpha_a * alpha_b), _lib.t_ptr(Cm), ldc, N, None, -1))
    torch.cuda.synchronize()
    got = Cm.sum(0)[:, :N]
    want = a.double() @ b.double().t()
    ok, msg = _report(got, want, f'f16 kmajor {M}x{N}x{K}/{splits}')
    assert ok, msg


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (300, 208, 200), (1000, 56, 1000), (26744, 208, 512)])
def test_umma_f16_mnmajor_with_ones_column(ctx, M, N, K):
    """dW'^T = dz^T h in fp16 hi/lo form with A MN-major (A given as G[k][m], the layout dz is written in): the transposed
    operand is read through the MN-major SWIZZLE_128B descriptor; the constant-one row of B folds db' into the product."""
    import torch
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    G = (torch.rand((K, M), device='cuda', generator=g) - 0.5) * 2.0          # dz / inv_count in [-1, 1]
    G = G * torch.pow(10.0, -4 * torch.rand((K, M), device='cuda', generator=g))
    H_ = torch.rand((K, N - 1), device='cuda', generator=g) * 0.999           # activations [batch, hidden]
    alpha_g, alpha_h = 16384.0, 32768.0
    g_hi, g_lo, _, _ = _split_f16(ctx, G, alpha_g)                            # [K][ldm] : m contiguous
    _, _, ht_hi, ht_lo = _split_f16(ctx, H_, alpha_h, transpose=True, ones_row=N - 1, t_rows=N)
    lib = _lib.load()
    ldc = (N + 3) // 4 * 4
    Cm = torch.full((1, M, ldc), float('nan'), device='cuda')
    extra = torch.full((M,), float('nan'), device='cuda')
    _lib.check(lib.drb_debug_umma_gemm_f16(ctx, _lib.t_ptr(g_hi), _lib.t_ptr(g_lo), g_hi.shape[1], _lib.t_ptr(ht_hi),
                                           _lib.t_ptr(ht_lo), ht_hi.shape[1], N, 1, M, N, K, 1,
                                           1.0 / (alpha_g * alpha_h), _lib.t_ptr(Cm), ldc, N - 1, _lib.t_ptr(extra), N - 1))
    torch.cuda.synchronize()
    want = G.double().t() @ H_.double()
    ok, msg = _report(Cm[0][:, :N - 1], want, f'f16 mn-major {M}x{N}x{K}')
    assert ok, msg
    colsum = G.double().sum(0)
    assert float((extra.double() - colsum).abs().max()) <= 1e-5 * float(colsum.abs().max()), 'ones column (db)'

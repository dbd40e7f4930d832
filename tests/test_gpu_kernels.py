"""GPU parity of the individual C-ABI kernels against plain torch fp32 references of the same op."""
import math

import pytest
import torch

from icka_b200 import ops
from icka_b200._lib import ACT_GELU_ERF, ACT_NONE

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def rel(a, b):
    return float(((a - b).abs() / b.abs().clamp(min=1.0)).max())


def test_cast_bf16():
    for n in (8, 1000, 128 * 768 + 3):
        x = rnd(n, seed=n)
        y = ops.cast_bf16(x.to(DEV)).cpu()
        assert torch.equal(y, x.bfloat16())


@pytest.mark.parametrize('B,C,R', [(2, 2048, 49), (1, 2048, 196), (3, 64, 9), (2, 100, 7), (3, 512, 50), (2, 256, 1), (1, 2048, 8), (2, 128, 36), (3, 64, 4), (2, 192, 100), (5, 2048, 196),
                                   (70, 2048, 49)])
@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
def test_region_rows(B, C, R, dt):
    g = rnd(B, C, R, seed=R)
    got = ops.region_rows(g.to(DEV), dt).cpu()
    want = g.permute(0, 2, 1).reshape(B * R, C).to(dt)
    assert torch.equal(got, want)


@pytest.mark.parametrize('M,N', [(5, 768), (64, 1024), (9, 128), (3, 4)])
def test_layernorm(M, N):
    x = rnd(M, N, seed=M, scale=3.0) + 0.5
    w, b = rnd(N, seed=1) * 0.1 + 1, rnd(N, seed=2) * 0.1
    for eps in (1e-12, 1e-5):
        y32, y16 = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), eps, want_f32=True, want_bf16=True)
        u = x.mean(-1, keepdim=True)
        s = (x - u).pow(2).mean(-1, keepdim=True)
        want = w * ((x - u) / torch.sqrt(s + eps)) + b
        assert rel(y32.cpu(), want) <= 2e-6
        assert torch.equal(y16.cpu(), y32.cpu().bfloat16())


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (100, 768, 768), (49 * 3, 768, 2048), (7, 3072, 768),
                                    (130, 768, 3072), (3, 1, 128), (2, 768, 512), (257, 130, 36)])
@pytest.mark.parametrize('act', [ACT_NONE, ACT_GELU_ERF])
def test_linear_fp32(M, N, K, act):
    a, w, bias = rnd(M, K, seed=1), rnd(N, K, seed=2) / math.sqrt(K), rnd(N, seed=3)
    res = rnd(M, N, seed=4)
    got = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), residual=res.to(DEV), act=act).cpu()
    want = (a.double() @ w.double().t() + bias.double())
    want = (gelu(want) if act else want) + res.double()
    # fp32 accumulation over K terms vs an fp64 reference: sqrt(K)-growth of the rounding error
    assert rel(got.double(), want) <= 4e-7 * math.sqrt(K) + 1e-6


def test_linear_fp32_pitched_views():
    a_full, w = rnd(50, 256, seed=1), rnd(64, 128, seed=2)
    a = a_full[:, 128:]
    got = ops.linear(a.to(DEV) if False else a_full.to(DEV)[:, 128:], w.to(DEV), None).cpu()
    assert rel(got, a @ w.t()) <= 3e-6


def test_cross_attn_core_fp32_and_bf16():
    for (B, Sq, Skv, nh) in [(3, 128, 49, 12), (2, 1, 128, 12), (1, 256, 196, 12), (2, 16, 9, 2),
                             (70, 1, 37, 12), (3, 1, 1, 2), (2, 1, 300, 16)]:      # single query: attention_sq1.cu
        H = nh * 64
        q, kv = rnd(B * Sq, H, seed=1), rnd(B * Skv, 2 * H, seed=2)
        mask = torch.zeros(B, Skv); mask[:, Skv // 2:] = -10000.0; mask[0] = 0
        qh = q.view(B, Sq, nh, 64).permute(0, 2, 1, 3)
        kh = kv[:, :H].reshape(B, Skv, nh, 64).permute(0, 2, 1, 3)
        vh = kv[:, H:].reshape(B, Skv, nh, 64).permute(0, 2, 1, 3)
        sc = qh @ kh.transpose(-1, -2) / 8.0 + mask.view(B, 1, 1, Skv)
        want = (torch.softmax(sc, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Sq, H)
        kvd = kv.to(DEV)
        got = ops.cross_attn_core(q.to(DEV), kvd[:, :H], kvd[:, H:], mask.to(DEV), B, Sq, Skv, nh, 64).cpu()
        assert rel(got, want) <= 3e-6
        kvb = kvd.bfloat16()
        gotb = ops.cross_attn_core(q.to(DEV).bfloat16(), kvb[:, :H], kvb[:, H:], mask.to(DEV), B, Sq, Skv, nh, 64)
        assert rel(gotb.float().cpu(), want) <= 3e-2


@pytest.mark.parametrize('B,Sq,Skv,nh', [(3, 128, 49, 12), (2, 16, 9, 2), (5, 128, 64, 1), (2, 300, 49, 3), (40, 128, 49, 12),
                                          (1, 1, 1, 1), (2, 256, 196, 12), (30, 256, 196, 12), (3, 128, 128, 4), (3, 100, 100, 2),
                                          (2, 128, 224, 1), (2, 64, 225, 2), (4, 1, 128, 12)])
@pytest.mark.parametrize('attn_mode', [0, 1, 2, 3], ids=['tcgen05', 'mma_sync', 'tcgen05_wide', 'tcgen05_wide2'])
def test_cross_attn_core_bf16_kernels(B, Sq, Skv, nh, attn_mode):
    """Both bf16 attention kernels (tcgen05/TMEM for Skv <= 64, mma.sync) against the fp64 formula on the same
    bf16-rounded operands; ragged Sq (rows of the next sentence / zero fill inside the Q box) and key masks."""
    from icka_b200 import _lib
    H = nh * 64
    q = rnd(B * Sq, H, seed=3).bfloat16()
    kv = rnd(B * Skv, 2 * H, seed=4).bfloat16()
    mask = torch.zeros(B, Skv)
    mask[:, (Skv + 1) // 2:] = -10000.0
    mask[0] = 0
    qh = q.double().view(B, Sq, nh, 64).permute(0, 2, 1, 3)
    kh = kv[:, :H].double().reshape(B, Skv, nh, 64).permute(0, 2, 1, 3)
    vh = kv[:, H:].double().reshape(B, Skv, nh, 64).permute(0, 2, 1, 3)
    sc = qh @ kh.transpose(-1, -2) / 8.0 + mask.double().view(B, 1, 1, Skv)
    want = (torch.softmax(sc, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Sq, H)
    _lib.check(_lib.load().icka_set_attn_mode(attn_mode), 'icka_set_attn_mode')
    try:
        kvd = kv.to(DEV)
        got = ops.cross_attn_core(q.to(DEV), kvd[:, :H], kvd[:, H:], mask.to(DEV), B, Sq, Skv, nh, 64)
        torch.cuda.synchronize()
    finally:
        _lib.load().icka_set_attn_mode(0)
    err = float(((got.double().cpu() - want).abs() / want.abs().clamp(min=1.0)).max())
    assert err <= 2e-2, err     # P rounded to bf16 for the second GEMM + bf16 output


@pytest.mark.parametrize('B,S,H', [(5, 128, 768), (3, 37, 64), (2, 4, 6), (40, 256, 1024)])
def test_gate_fold_and_blend(B, S, H):
    fused, tok = rnd(B, S, H, seed=1), rnd(B, S, H, seed=2)
    lw, lb = rnd(H, seed=3) * 0.1 + 1, rnd(H, seed=4) * 0.1
    wp, bp, wa, ba = rnd(H, H, seed=5) / math.sqrt(H), rnd(H, seed=6), rnd(H, seed=7) / math.sqrt(H), rnd(1, seed=8)
    wf, cf = ops.gate_fold(wp.to(DEV), bp.to(DEV), wa.to(DEV), ba.to(DEV))
    assert rel(wf.cpu(), wa @ wp) <= 2e-6
    out, gate = ops.gate_blend(fused.to(DEV), tok.to(DEV), lw.to(DEV), lb.to(DEV), 1e-5, wf, cf)
    feat = torch.nn.functional.layer_norm(fused[:, 0] + tok[:, 0], (H,), lw, lb, 1e-5)
    g = torch.sigmoid((feat @ wp.t() + bp) @ wa + ba).view(B, 1, 1)
    assert rel(gate.cpu(), g.view(-1)) <= 2e-6
    assert rel(out.cpu(), g * tok + (1 - g) * fused) <= 2e-6


@pytest.mark.parametrize('B,S,H,nh', [(3, 128, 768, 12), (2, 256, 768, 12), (2, 40, 1024, 16), (1, 7, 768, 12), (300, 128, 768, 12),
                                      (5, 100, 768, 16), (300, 256, 768, 12), (7, 200, 768, 12), (3, 129, 1024, 16), (2, 300, 768, 12)])
@pytest.mark.parametrize('attn_mode', [0, 1], ids=['tcgen05', 'mma_sync'])
def test_i2t_pool(B, S, H, nh, attn_mode):
    from icka_b200 import _lib
    u = (rnd(B, nh * H, seed=1) / math.sqrt(H)).bfloat16()
    x = rnd(B * S, H, seed=2).bfloat16()
    mask = torch.zeros(B, S); mask[:, (S * 2) // 3:] = -10000.0; mask[0] = 0
    uf, xf = u.float().view(B, nh, H), x.float().view(B, S, H)
    sc = torch.einsum('bhd,bsd->bhs', uf, xf) / 8.0 + mask.view(B, 1, S)
    want = torch.einsum('bhs,bsd->bhd', torch.softmax(sc, -1), xf).reshape(B, nh * H)
    _lib.check(_lib.load().icka_set_attn_mode(attn_mode), 'icka_set_attn_mode')
    try:
        got = ops.i2t_pool(u.to(DEV), x.to(DEV), mask.to(DEV), B, S, H, nh).float().cpu()
        torch.cuda.synchronize()
    finally:
        _lib.load().icka_set_attn_mode(0)
    assert rel(got, want) <= 2e-2


def test_tcgen05_kernels_long_pipelines_are_repeatable():
    """Many work items per persistent CTA (the TMA rings wrap several times, both softmax groups and both TMEM slots
    are re-used) -- results must match the mma.sync kernels' within bf16 noise and be bit-identical run to run."""
    from icka_b200 import _lib
    lib = _lib.load()
    B, Sq, Skv, nh, H = 260, 128, 49, 12, 768
    q = rnd(B * Sq, H, seed=21).bfloat16().to(DEV)
    kv = rnd(B * Skv, 2 * H, seed=22).bfloat16().to(DEV)
    mask = torch.zeros(B, Skv)
    mask[1::2, 40:] = -10000.0
    mask = mask.to(DEV)
    a1 = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask, B, Sq, Skv, nh, 64)
    a2 = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask, B, Sq, Skv, nh, 64)
    u = (rnd(B * 3, nh * H, seed=23) / math.sqrt(H)).bfloat16().to(DEV)
    x = rnd(B * 3 * 128, H, seed=24).bfloat16().to(DEV)
    tm = torch.zeros(B * 3, 128)
    tm[::3, 30:] = -10000.0
    tm = tm.to(DEV)
    p1 = ops.i2t_pool(u, x, tm, B * 3, 128, H, nh)
    p2 = ops.i2t_pool(u, x, tm, B * 3, 128, H, nh)
    _lib.check(lib.icka_set_attn_mode(1), 'icka_set_attn_mode')
    try:
        a_ref = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask, B, Sq, Skv, nh, 64)
        p_ref = ops.i2t_pool(u, x, tm, B * 3, 128, H, nh)
        torch.cuda.synchronize()
    finally:
        lib.icka_set_attn_mode(0)
    assert torch.equal(a1, a2) and torch.equal(p1, p2)
    assert float((a1.float() - a_ref.float()).abs().max()) <= 3e-2
    assert float((p1.float() - p_ref.float()).abs().max()) <= 3e-2

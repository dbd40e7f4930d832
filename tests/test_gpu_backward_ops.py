"""GPU parity of the backward kernels (training path) against torch autograd on the CPU oracle formulas.

GEMM backward: dgrad / wgrad on the tcgen05 kernel with MN-major operands (bf16) and on the FFMA kernels
(fp32), vs float64 matmuls of the same (bf16-rounded) operands.  LayerNorm / attention / gate / CRF
backward vs autograd through oracle/fusion_ref.py and oracle/crf_ref.py in float64.
"""
import math

import pytest
import torch

from icka_b200 import ops, synth
from oracle import crf_ref, fusion_ref

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


# ---- dgrad ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (128, 128, 256), (256, 768, 768), (300, 768, 3072),
                                    (1000, 3072, 768), (7, 1536, 768), (128 * 9, 768, 768), (64, 768, 128)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_linear_dgrad(M, N, K, dtype):
    dy = rnd(M, N, seed=1).to(dtype)
    w = (rnd(N, K, seed=2) / math.sqrt(N)).to(dtype)
    res = rnd(M, K, seed=3)
    got = ops.linear_dgrad(dy.to(DEV), w.to(DEV), residual=res.to(DEV), out_dtype=torch.float32).cpu()
    want = dy.double() @ w.double() + res.double()
    err = float((got.double() - want).abs().max())
    assert err <= 1e-6 * math.sqrt(N) + 3e-6, err


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_linear_dgrad_gelu(dtype):
    M, N, K = 300, 768, 3072
    dy = rnd(M, N, seed=4).to(dtype)
    w = (rnd(N, K, seed=5) / math.sqrt(N)).to(dtype)
    u = rnd(M, K, seed=6, scale=1.5).to(dtype)
    got = ops.linear_dgrad(dy.to(DEV), w.to(DEV), gelu_pre=u.to(DEV), out_dtype=dtype).cpu()
    want = (dy.double() @ w.double()) * gelu_grad(u.double())
    err = float(((got.double() - want).abs() / want.abs().clamp(min=1.0)).max())
    assert err <= (2 ** -7 if dtype == torch.bfloat16 else 1e-5), err


# ---- wgrad ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K', [(64, 64, 64), (128, 128, 128), (512, 768, 768), (1000, 3072, 768),
                                    (300, 768, 3072), (49 * 7, 768, 2048), (128 * 40, 1536, 768), (5, 768, 768)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_linear_wgrad(M, N, K, dtype):
    dy = rnd(M, N, seed=7).to(dtype)
    x = rnd(M, K, seed=8).to(dtype)
    got = ops.linear_wgrad(dy.to(DEV), x.to(DEV)).cpu()
    want = dy.double().t() @ x.double()
    err = float((got.double() - want).abs().max())
    assert err <= 2e-6 * M + 1e-5, err
    # accumulate onto an existing gradient
    base = rnd(N, K, seed=9)
    acc = base.clone().to(DEV)
    ops.linear_wgrad(dy.to(DEV), x.to(DEV), out=acc, accumulate=True)
    err = float((acc.cpu().double() - (want + base.double())).abs().max())
    assert err <= 2e-6 * M + 1e-5, err


def test_linear_wgrad_pitched_halves():
    """dK and dV are the column halves of one [dK|dV] buffer (row pitch 2H)."""
    M, H = 392, 768
    dkv = rnd(M, 2 * H, seed=10).bfloat16().to(DEV)
    y = rnd(M, H, seed=11).bfloat16().to(DEV)
    got = ops.linear_wgrad(dkv, y).cpu()
    want = dkv.cpu().double().t() @ y.cpu().double()
    assert float((got.double() - want).abs().max()) <= 2e-6 * M + 1e-5


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('M,N', [(1, 768), (1000, 768), (4096, 3072), (333, 130)])
def test_colsum(M, N, dtype):
    x = rnd(M, N, seed=12).to(dtype)
    got = ops.colsum(x.to(DEV)).cpu()
    want = x.double().sum(0)
    assert float((got.double() - want).abs().max()) <= 1e-6 * M + 1e-5


# ---- LayerNorm --------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,eps', [(1, 768, 1e-12), (1000, 768, 1e-12), (4100, 768, 1e-5), (300, 1024, 1e-12),
                                      (77, 128, 1e-5)])
def test_layernorm_bwd(M, N, eps):
    x = rnd(M, N, seed=13, scale=2.0)
    dy = rnd(M, N, seed=14)
    gamma, beta = 1.0 + 0.1 * rnd(N, seed=15), 0.1 * rnd(N, seed=16)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    fusion_ref.bert_layer_norm(xd, gd, bd, eps).backward(dy.double())
    dx32, dx16, dg, db, dbias = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), gamma.to(DEV), eps, want_f32=True,
                                                  want_bf16=True)
    assert float((dx32.cpu().double() - xd.grad).abs().max()) <= 2e-5
    assert float((dx16.cpu().double() - xd.grad).abs().max()) <= 2 ** -8 * float(xd.grad.abs().max()) + 1e-5
    tol = 2e-5 * math.sqrt(M) + 1e-5
    assert float((dg.cpu().double() - gd.grad).abs().max()) <= tol
    assert float((db.cpu().double() - bd.grad).abs().max()) <= tol
    assert float((dbias.cpu().double() - xd.grad.sum(0)).abs().max()) <= tol


# ---- attention core ---------------------------------------------------------------------------------
def attn_ref(q, k, v, mask_add, B, Sq, Skv, nh, d):
    qh = q.view(B, Sq, nh, d).permute(0, 2, 1, 3)
    kh = k.view(B, Skv, nh, d).permute(0, 2, 1, 3)
    vh = v.view(B, Skv, nh, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if mask_add is not None:
        s = s + mask_add.view(B, 1, 1, Skv)
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Sq, nh * d)


@pytest.mark.parametrize('B,Sq,Skv,nh', [(3, 128, 49, 12), (2, 1, 128, 12), (2, 77, 49, 4), (1, 200, 64, 2),
                                          (2, 128, 100, 3), (2, 128, 196, 2), (70, 1, 37, 12), (3, 1, 1, 2),
                                          (2, 1, 300, 16)])
@pytest.mark.parametrize('kind', ['fp32', 'bf16', 'bf16_tensor_core'])
def test_cross_attn_core_bwd(B, Sq, Skv, nh, kind):
    """fp32 / bf16: CUDA-core kernel; bf16_tensor_core: mma.sync kernel (needs the forward output, Sq <= 128)."""
    if kind != 'bf16_tensor_core' and Skv > 150 and Sq > 1:
        pytest.skip('the CUDA-core kernel keeps all keys in shared memory')
    dtype = torch.float32 if kind == 'fp32' else torch.bfloat16
    d, H = 64, nh * 64
    q = rnd(B * Sq, H, seed=17).to(dtype)
    kv = rnd(B * Skv, 2 * H, seed=18).to(dtype)
    dctx = rnd(B * Sq, H, seed=19).to(dtype)
    lens = torch.randint(1, Skv + 1, (B,), generator=torch.Generator().manual_seed(20))
    mask_add = (1.0 - synth.prefix_mask(lens, Skv).float()) * -10000.0
    qd = q.double().requires_grad_(True)
    kvd = kv.double().requires_grad_(True)
    attn_ref(qd, kvd[:, :H], kvd[:, H:], mask_add.double(), B, Sq, Skv, nh, d).backward(dctx.double())
    q_dev, kv_dev, m_dev = q.to(DEV), kv.to(DEV), mask_add.to(DEV)
    ctx = None
    if kind == 'bf16_tensor_core':
        ctx = ops.cross_attn_core(q_dev, kv_dev[:, :H], kv_dev[:, H:], m_dev, B, Sq, Skv, nh, d)
    dq, dkv = ops.cross_attn_core_bwd(q_dev, kv_dev[:, :H], kv_dev[:, H:], m_dev, dctx.to(DEV), B, Sq, Skv, nh, d,
                                      ctx=ctx)
    # bf16: outputs rounded to bf16; the tensor-core kernel also rounds P and dS to bf16 before the second GEMMs
    tol = {'fp32': 2e-5, 'bf16': 2 ** -7, 'bf16_tensor_core': 4e-2}[kind]
    for got, want in ((dq, qd.grad), (dkv, kvd.grad)):
        err = float(((got.cpu().double() - want).abs() / want.abs().clamp(min=1.0)).max())
        assert err <= tol, err


# ---- gate + blend -----------------------------------------------------------------------------------
def test_gate_blend_bwd():
    B, S, H = 5, 128, 768
    fused, tok, dout = rnd(B, S, H, seed=21), rnd(B, S, H, seed=22), rnd(B, S, H, seed=23)
    p = {'cls_layer.proj_norm.weight': 1.0 + 0.1 * rnd(H, seed=24), 'cls_layer.proj_norm.bias': 0.1 * rnd(H, seed=25),
         'cls_layer.proj.weight': rnd(H, H, seed=26) / math.sqrt(H), 'cls_layer.proj.bias': 0.1 * rnd(H, seed=27),
         'aux_head.weight': rnd(1, H, seed=28) / math.sqrt(H), 'aux_head.bias': 0.1 * rnd(1, seed=29)}
    pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
    fd, td = fused.double().requires_grad_(True), tok.double().requires_grad_(True)
    out = fusion_ref.gate_blend(fd, td, pd)
    res = out[0] if isinstance(out, (tuple, list)) else out
    res.backward(dout.double())

    dev = {k: v.to(DEV) for k, v in p.items()}
    w_fold, c_fold = ops.gate_fold(dev['cls_layer.proj.weight'], dev['cls_layer.proj.bias'],
                                   dev['aux_head.weight'].view(-1).contiguous(), dev['aux_head.bias'])
    f_dev, t_dev = fused.to(DEV), tok.to(DEV)
    _, gate = ops.gate_blend(f_dev, t_dev, dev['cls_layer.proj_norm.weight'], dev['cls_layer.proj_norm.bias'], 1e-5,
                             w_fold, c_fold)
    dfused, dtok, d_ln_w, d_ln_b, d_wf, d_cf = ops.gate_blend_bwd(
        dout.to(DEV), f_dev, t_dev, gate, dev['cls_layer.proj_norm.weight'], dev['cls_layer.proj_norm.bias'], 1e-5,
        w_fold)
    dwp, dbp, dwa, dba = ops.gate_fold_bwd(dev['cls_layer.proj.weight'], dev['cls_layer.proj.bias'],
                                           dev['aux_head.weight'].view(-1).contiguous(), d_wf, d_cf)

    def close(got, want, tol=2e-4):   # d(gate) is an fp32 sum over S*H = 98k products
        err = float(((got.cpu().double() - want).abs() / want.abs().clamp(min=1.0)).max())
        assert err <= tol, err

    close(dfused, fd.grad)
    close(dtok, td.grad)
    close(d_ln_w, pd['cls_layer.proj_norm.weight'].grad)
    close(d_ln_b, pd['cls_layer.proj_norm.bias'].grad)
    close(dwp, pd['cls_layer.proj.weight'].grad)
    close(dbp, pd['cls_layer.proj.bias'].grad)
    close(dwa.view(1, -1), pd['aux_head.weight'].grad)
    close(dba, pd['aux_head.bias'].grad)


# ---- CRF log-likelihood -----------------------------------------------------------------------------
@pytest.mark.parametrize('B,S,T,holes', [(33, 128, 15, False), (9, 40, 15, True), (5, 7, 3, False), (6, 33, 20, False),
                                          (3, 1, 15, False)])
def test_crf_llh_bwd(B, S, T, holes):
    sh = synth.Shape(S=S, T=T)
    batch = synth.crf_batch(B, sh, seed=41, median_len=max(1.0, S / 3))
    mask = batch['mask'].clone()
    if holes:
        mask = torch.rand(B, S, generator=torch.Generator().manual_seed(42)) > 0.3
    mask[:, 0] = True
    tags = batch['tags'] % T
    cp = synth.crf_params(T, 43, 'normal')
    w = rnd(B, seed=44)
    ed = batch['emissions'].double().requires_grad_(True)
    pd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    llh = crf_ref.log_likelihood(ed, tags, mask, pd['start_transitions'], pd['end_transitions'], pd['transitions'],
                                 'none')
    (llh * w.double()).sum().backward()
    de, ds, dend, dtr = ops.crf_llh_bwd(batch['emissions'].to(DEV), tags.to(DEV), mask.to(torch.uint8).to(DEV),
                                        cp['start_transitions'].to(DEV), cp['end_transitions'].to(DEV),
                                        cp['transitions'].to(DEV), w.to(DEV))
    # marginals are exp(alpha + beta - logZ) with |alpha|, |logZ| ~ 5 S in fp32: relative error ~ eps * |logZ|
    tol_e = max(2e-5, 1.5e-6 * S)
    assert float((de.cpu().double() - ed.grad).abs().max()) <= tol_e
    tol = tol_e * math.sqrt(B * S)
    zero = lambda g, like: torch.zeros_like(like) if g is None else g      # S == 1: transitions are unused
    assert float((ds.cpu().double() - pd['start_transitions'].grad).abs().max()) <= tol
    assert float((dend.cpu().double() - pd['end_transitions'].grad).abs().max()) <= tol
    assert float((dtr.cpu().double() - zero(pd['transitions'].grad, pd['transitions'])).abs().max()) <= tol

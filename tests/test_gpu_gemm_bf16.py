"""GPU parity of the tcgen05 bf16 GEMM (icka_linear_fwd, in_dtype=bf16) vs torch fp32 matmul on the same
bf16-rounded operands.  fp32 accumulation => errors are at the fp32 reassociation level for fp32 outputs
and half a bf16 ulp for bf16 outputs."""
import math

import pytest
import torch

from icka_b200 import ops
from icka_b200._lib import ACT_GELU_ERF, ACT_NONE

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


SHAPES = [(128, 256, 64), (128, 256, 16), (128, 128, 64), (128, 256, 768), (256, 768, 768), (49 * 5, 768, 2048),
          (1024, 3072, 768), (1000, 768, 3072), (7, 768, 512), (130, 1536, 768), (300, 136, 40), (128 * 40, 768, 768)]


@pytest.fixture(params=[0, 1, 2], ids=['auto', 'single_cta', 'cta_pair'])
def gemm_mode(request):
    from icka_b200 import _lib
    _lib.check(_lib.load().icka_set_gemm_mode(request.param), 'icka_set_gemm_mode')
    yield request.param
    _lib.load().icka_set_gemm_mode(0)


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_linear_bf16_fp32_out(M, N, K, gemm_mode):
    a = rnd(M, K, seed=1).bfloat16()
    w = (rnd(N, K, seed=2) / math.sqrt(K)).bfloat16()
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4)
    got = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), residual=res.to(DEV), out_dtype=torch.float32).cpu()
    want = a.double() @ w.double().t() + bias.double() + res.double()
    err = float((got.double() - want).abs().max())
    assert err <= 1e-6 * math.sqrt(K) + 2e-6, err     # fp32 accumulation over K exact bf16 products


@pytest.mark.parametrize('M,N,K', [(256, 3072, 768), (77, 256, 128), (128, 768, 768)])
def test_linear_bf16_gelu_bf16_out(M, N, K, gemm_mode):
    a = rnd(M, K, seed=5).bfloat16()
    w = (rnd(N, K, seed=6) / math.sqrt(K)).bfloat16()
    bias = rnd(N, seed=7)
    got = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), act=ACT_GELU_ERF, out_dtype=torch.bfloat16).cpu()
    want = gelu(a.double() @ w.double().t() + bias.double())
    err = float(((got.double() - want).abs() / want.abs().clamp(min=1.0)).max())
    # half a bf16 ulp (2^-9) from the output rounding + the epilogue's GELU evaluation (fitted tanh form on
    # MUFU.TANH, csrc/common.cuh: about another 2^-9 relative where |gelu| > 1)
    assert err <= 2 ** -7, err


def test_linear_bf16_pitched_kv_views(gemm_mode):
    """Consumers read K and V as column halves of one [K|V] buffer; A may be a pitched view too."""
    M, H = 200, 768
    a_full = rnd(M, 2 * H, seed=8).bfloat16().to(DEV)
    w = (rnd(H, H, seed=9) / math.sqrt(H)).bfloat16().to(DEV)
    got = ops.linear(a_full[:, H:], w, None, out_dtype=torch.float32).cpu()
    want = a_full[:, H:].cpu().double() @ w.cpu().double().t()
    assert float((got.double() - want).abs().max()) <= 2e-5


@pytest.fixture(params=[0, 1, 2], ids=['ln_auto', 'ln_single_cta', 'ln_cluster'])
def ln_mode(request):
    from icka_b200 import _lib
    _lib.check(_lib.load().icka_set_ln_mode(request.param), 'icka_set_ln_mode')
    yield request.param
    _lib.load().icka_set_ln_mode(0)


@pytest.mark.parametrize('M,N,K', [(128, 768, 768), (1000, 768, 3072), (300, 256, 64), (77, 1024, 512), (128 * 160, 768, 768),
                                    (5, 136, 40), (128 * 300 + 9, 768, 768), (1, 1024, 1024), (4100, 512, 128)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_linear_ln_fused_epilogue(M, N, K, dtype, ln_mode):
    """Dense + bias + residual + BertLayerNorm in one call (LayerNorm in the tcgen05 epilogue for bf16 operands: the
    cluster kernel for N = 512 / 768 / 1024, the single-CTA kernel otherwise)."""
    from oracle import fusion_ref
    if ln_mode == 2 and (dtype != torch.bfloat16 or N not in (512, 768, 1024)):
        pytest.skip('the cluster kernel serves bf16 operands with N = 512 / 768 / 1024')
    if ln_mode != 0 and dtype != torch.bfloat16:
        pytest.skip('fp32 operands take the FFMA path in every mode')
    a = rnd(M, K, seed=11).to(dtype)
    w = (rnd(N, K, seed=12) / math.sqrt(K)).to(dtype)
    bias, res = rnd(N, seed=13), rnd(M, N, seed=14)
    gamma, beta = 1.0 + 0.1 * rnd(N, seed=15), 0.1 * rnd(N, seed=16)
    eps = 1e-12
    y32, y16 = ops.linear_ln(a.to(DEV), w.to(DEV), bias.to(DEV), res.to(DEV), gamma.to(DEV), beta.to(DEV), eps,
                             want_bf16=True)
    pre = a.double() @ w.double().t() + bias.double() + res.double()
    want = fusion_ref.bert_layer_norm(pre, gamma.double(), beta.double(), eps)
    err = float((y32.cpu().double() - want).abs().max())
    assert err <= 2e-5, err
    err16 = float((y16.float().cpu().double() - want).abs().max())
    assert err16 <= 2 ** -8 * float(want.abs().max()) + 1e-5, err16

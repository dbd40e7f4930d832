"""GPU parity: Viterbi tags bit-exact vs the oracle; CRF log-likelihood within rel 1e-5 (fp32)."""
import numpy as np
import pytest
import torch

import icka_b200
from icka_b200 import synth
from oracle import crf_ref, viterbi_c

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def make_crf(T, seed, kind):
    cp = synth.crf_params(T, seed, kind)
    crf = icka_b200.CRF(T, batch_first=True).to(DEV)
    crf.load_state_dict(cp)
    return crf, cp


def oracle_c(e, mask, cp):
    tags, lens = viterbi_c.viterbi(e.numpy(), None if mask is None else mask.numpy(),
                                   cp['start_transitions'].numpy(), cp['end_transitions'].numpy(),
                                   cp['transitions'].numpy())
    return viterbi_c.to_lists(tags, lens)


@pytest.mark.parametrize('kind', ['normal', 'ties', 'near_ties'])
@pytest.mark.parametrize('pkind', ['uniform', 'normal'])
def test_viterbi_std_shape_bit_exact(kind, pkind):
    sh = synth.STD
    batch = synth.crf_batch(300, sh, seed=11, kind=kind)
    crf, cp = make_crf(sh.T, 12, pkind)
    got = crf.decode(batch['emissions'].to(DEV), batch['mask'].to(DEV))
    assert got == oracle_c(batch['emissions'], batch['mask'], cp)
    # and the torch restatement agrees on a slice (the C port is itself pinned to it on CPU)
    want = crf_ref.viterbi_decode(batch['emissions'][:32], batch['mask'][:32], cp['start_transitions'],
                                  cp['end_transitions'], cp['transitions'])
    assert got[:32] == want


@pytest.mark.parametrize('B,S,T', [(1, 1, 15), (3, 2, 15), (5, 7, 3), (17, 128, 16), (9, 33, 17), (4, 50, 32),
                                    (2, 256, 15), (33, 13, 1), (7, 19, 9)])
def test_viterbi_shapes(B, S, T):
    sh = synth.Shape(S=S, T=T)
    batch = synth.crf_batch(B, sh, seed=B * 7 + S, kind='ties', median_len=max(1.0, S / 3))
    mask = batch['mask'].clone(); mask[:, 0] = True
    crf, cp = make_crf(T, 5, 'normal')
    got = crf.decode(batch['emissions'].to(DEV), mask.to(DEV))
    assert got == oracle_c(batch['emissions'], mask, cp)


def test_viterbi_no_mask_and_full_length():
    sh = synth.STD
    e = synth.emissions(40, sh.S, sh.T, seed=3, kind='near_ties')
    crf, cp = make_crf(sh.T, 2, 'uniform')
    got = crf.decode(e.to(DEV))
    assert got == oracle_c(e, None, cp)
    assert all(len(g) == sh.S for g in got)


def test_viterbi_masks_with_holes_follow_the_procedure():
    sh = synth.Shape(S=40, T=15)
    g = torch.Generator().manual_seed(9)
    e = synth.emissions(64, sh.S, sh.T, seed=4, kind='ties')
    mask = torch.rand(64, sh.S, generator=g) > 0.35
    mask[:, 0] = True
    crf, cp = make_crf(sh.T, 8, 'normal')
    got = crf.decode(e.to(DEV), mask.to(DEV))
    assert got == oracle_c(e, mask, cp)


def test_viterbi_time_major_and_int_mask():
    sh = synth.Shape(S=21, T=15)
    batch = synth.crf_batch(10, sh, seed=1, median_len=8)
    cp = synth.crf_params(sh.T, 3, 'normal')
    crf = icka_b200.CRF(sh.T).to(DEV)          # batch_first=False, pytorch-crf's default
    crf.load_state_dict(cp)
    got = crf.decode(batch['emissions'].transpose(0, 1).to(DEV), batch['mask'].long().transpose(0, 1).to(DEV))
    assert got == oracle_c(batch['emissions'], batch['mask'], cp)


def test_viterbi_rounding_induced_tie():
    T = 2
    crf = icka_b200.CRF(T, batch_first=True).to(DEV)
    one_up = float(np.nextafter(np.float32(1.0), np.float32(2.0)))
    crf.load_state_dict(dict(start_transitions=torch.tensor([1.0, one_up]), end_transitions=torch.zeros(T),
                             transitions=torch.zeros(T, T)))
    e = torch.zeros(1, 2, T); e[0, 1, :] = 1024.0
    assert crf.decode(e.to(DEV))[0][0] == 0


def test_viterbi_full_batch_property():
    """BASELINE sweep size (B=4096): bit-exact vs the C oracle, and the returned path re-scores to the DP max."""
    sh = synth.STD
    batch = synth.crf_batch(4096, sh, seed=21, kind='normal')
    crf, cp = make_crf(sh.T, 22, 'uniform')
    tags, lens = crf.decode_tensors(batch['emissions'].to(DEV), batch['mask'].to(DEV))
    t_ref, l_ref = viterbi_c.viterbi(batch['emissions'].numpy(), batch['mask'].numpy(), cp['start_transitions'].numpy(),
                                     cp['end_transitions'].numpy(), cp['transitions'].numpy())
    assert np.array_equal(lens.cpu().numpy(), l_ref)
    assert np.array_equal(tags.cpu().numpy(), t_ref)


@pytest.mark.parametrize('kind', ['ties', 'near_ties'])
def test_viterbi_large_batch_two_sentences_per_warp(kind):
    """Above 32 sentences per SM the decode switches from one sentence per warp to two (odd tail included)."""
    sh = synth.STD
    B = 32 * 148 + 1265
    batch = synth.crf_batch(B, sh, seed=77, kind=kind)
    crf, cp = make_crf(sh.T, 78, 'normal')
    tags, lens = crf.decode_tensors(batch['emissions'].to(DEV), batch['mask'].to(DEV))
    t_ref, l_ref = viterbi_c.viterbi(batch['emissions'].numpy(), batch['mask'].numpy(), cp['start_transitions'].numpy(),
                                     cp['end_transitions'].numpy(), cp['transitions'].numpy())
    assert np.array_equal(lens.cpu().numpy(), l_ref)
    assert np.array_equal(tags.cpu().numpy(), t_ref)


@pytest.mark.parametrize('reduction', ['none', 'sum', 'mean', 'token_mean'])
def test_crf_llh(reduction):
    sh = synth.STD
    batch = synth.crf_batch(65, sh, seed=31)
    crf, cp = make_crf(sh.T, 32, 'normal')
    got = crf(batch['emissions'].to(DEV), batch['tags'].to(DEV), batch['mask'].to(DEV), reduction=reduction).cpu()
    want = crf_ref.log_likelihood(batch['emissions'], batch['tags'], batch['mask'], cp['start_transitions'],
                                  cp['end_transitions'], cp['transitions'], reduction)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-4), float((got - want).abs().max())


@pytest.mark.parametrize('value', [float('-inf'), -10000.0])
@pytest.mark.parametrize('T', [15, 20])
def test_viterbi_forbidden_cells(value, T):
    """Constrained decoding: -inf / -10000 emissions on a random 30 % of the (step, tag) cells, one step with no allowed
    tag at all.  Tags stay bit-exact (both the T <= 16 lock-step kernel and the generic one).  NaN emissions are outside
    the contract (the kernels' max is fmaxf-based; torch.max would propagate the NaN)."""
    sh = synth.Shape(S=128, T=T)
    batch = synth.crf_batch(130, sh, seed=21, kind='ties')
    e = synth.forbid_cells(batch['emissions'], 4, value=value)
    e[0, 2, :] = value
    crf, cp = make_crf(T, 13, 'normal')
    got = crf.decode(e.to(DEV), batch['mask'].to(DEV))
    assert got == oracle_c(e, batch['mask'], cp)


# ---- long sequences: fewer sentences per block so the per-sentence shared-memory slabs fit -----------
@pytest.mark.parametrize('S', [256, 600, 900])
def test_viterbi_long_sequences(S):
    """S = 256 is the hi-res shape (BASELINE configs[3]); from S ~ 470 the 8-sentence layout no longer fits 227 KB."""
    sh = synth.Shape(S=S, T=15)
    batch = synth.crf_batch(37, sh, seed=S, kind='ties', median_len=S * 0.6)
    crf, cp = make_crf(15, 41, 'normal')
    got = crf.decode(batch['emissions'].to(DEV), batch['mask'].to(DEV))
    assert got == oracle_c(batch['emissions'], batch['mask'], cp)


@pytest.mark.parametrize('S', [256, 400])
def test_crf_llh_forward_and_backward_long_sequences(S):
    """The CRF loss of the hi-res training shape (S = 256, `bench.py --mode train --hires`): the backward kept alpha[S][16]
    for 8 sentences per block (271 KB) and was refused; now 4 (or 2) sentences per block."""
    sh = synth.Shape(S=S, T=15)
    batch = synth.crf_batch(21, sh, seed=S + 1, median_len=S * 0.5)
    crf, cp = make_crf(15, 43, 'uniform')
    e = batch['emissions'].to(DEV).requires_grad_(True)
    loss = -crf(e, batch['tags'].to(DEV), batch['mask'].to(DEV), reduction='token_mean')
    loss.backward()
    ed = batch['emissions'].double().requires_grad_(True)
    pd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    want = -crf_ref.log_likelihood(ed, batch['tags'], batch['mask'], pd['start_transitions'], pd['end_transitions'],
                                   pd['transitions'], 'token_mean')
    want.backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert float((e.grad.cpu().double() - ed.grad).abs().max()) <= 1e-6
    for name, p in crf.named_parameters():
        assert float((p.grad.cpu().double() - pd[name].grad).abs().max()) <= 2e-5, name


def test_crf_tags_outside_the_label_range_raise_index_error():
    """pytorch-crf indexes its parameter tensors with every tag id: -100 (ignore-index padding) or T is an IndexError."""
    sh = synth.Shape(S=12, T=15)
    batch = synth.crf_batch(4, sh, seed=3)
    crf, _ = make_crf(15, 5, 'uniform')
    for bad in (-100, 15):
        tags = batch['tags'].clone()
        tags[2, 11] = bad                     # a masked (padding) position counts too
        with pytest.raises(IndexError):
            crf(batch['emissions'].to(DEV), tags.to(DEV), batch['mask'].to(DEV))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_entry_points_run_on_the_tensor_device_not_the_current_one():
    """Tensors on cuda:1 while cuda:0 is the current device (one nn.DataParallel worker thread per GPU,
    My_cross_attention.py:777-779): every C entry point sets the handle's device for the call and restores it."""
    sh = synth.STD
    batch = synth.crf_batch(64, sh, seed=5, kind='ties')
    cp = synth.crf_params(sh.T, 6, 'normal')
    crf = icka_b200.CRF(sh.T, batch_first=True).to('cuda:1')
    crf.load_state_dict(cp)
    assert torch.cuda.current_device() == 0
    got = crf.decode(batch['emissions'].to('cuda:1'), batch['mask'].to('cuda:1'))
    assert torch.cuda.current_device() == 0
    assert got == oracle_c(batch['emissions'], batch['mask'], cp)
    from icka_b200 import ops
    a = torch.randn(256, 768, device='cuda:1').bfloat16()
    w = torch.randn(768, 768, device='cuda:1').bfloat16()
    out = ops.linear(a, w, None, out_dtype=torch.float32)          # tcgen05 path: cudaFuncSetAttribute + TMA descriptors
    torch.cuda.synchronize('cuda:1')
    assert float((out.double() - a.double() @ w.double().t()).abs().max()) <= 1e-3

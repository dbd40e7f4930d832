"""Anchor oracle/crf_ref.py + oracle/viterbi_ref.c (parity unpinned: no torchcrf, no reference vectors)."""
import numpy as np
import pytest
import torch

from icka_b200 import synth
from oracle import crf_ref, viterbi_c


def P(T, seed=0, kind='uniform'):
    p = synth.crf_params(T, seed, kind)
    return p['start_transitions'], p['end_transitions'], p['transitions']


@pytest.mark.parametrize('T,S', [(2, 1), (3, 4), (4, 6), (2, 7)])
@pytest.mark.parametrize('kind', ['uniform', 'normal'])
def test_viterbi_vs_brute_force(T, S, kind):
    st, en, tr = P(T, 3, kind)
    e = synth.emissions(5, S, T, seed=T * 100 + S, kind='normal')
    for L in {1, max(1, S // 2), S}:
        mask = torch.zeros(5, S, dtype=torch.bool); mask[:, :L] = True
        paths = crf_ref.viterbi_decode(e, mask, st, en, tr)
        for b in range(5):
            best, arg = crf_ref.brute_force_best(e[b, :L], L, st, en, tr)
            assert len(paths[b]) == L
            got = crf_ref._seq_score(e[b].double(), paths[b], st.double(), en.double(), tr.double())
            assert abs(got - best) < 1e-4
            if len(arg) == 1:
                assert paths[b] == arg[0]


@pytest.mark.parametrize('kind', ['normal', 'ties', 'near_ties'])
@pytest.mark.parametrize('pkind', ['uniform', 'normal'])
def test_c_port_matches_torch_restatement(kind, pkind):
    sh = synth.Shape(S=40, T=15)
    batch = synth.crf_batch(37, sh, seed=5, kind=kind, median_len=12)
    st, en, tr = P(sh.T, 9, pkind)
    want = crf_ref.viterbi_decode(batch['emissions'], batch['mask'], st, en, tr)
    tags, lens = viterbi_c.viterbi(batch['emissions'].numpy(), batch['mask'].numpy(), st.numpy(), en.numpy(), tr.numpy())
    assert viterbi_c.to_lists(tags, lens) == want
    assert (lens == batch['lens'].numpy()).all()


def test_c_port_holes_in_mask_follow_the_procedure():
    """Non-prefix masks: replicate upstream's procedure (history[:len-1]), not an idealised semantics."""
    sh = synth.Shape(S=12, T=5)
    g = torch.Generator().manual_seed(1)
    e = synth.emissions(16, sh.S, sh.T, seed=2, kind='ties')
    mask = torch.rand(16, sh.S, generator=g) > 0.4
    mask[:, 0] = True
    st, en, tr = P(sh.T, 4, 'normal')
    want = crf_ref.viterbi_decode(e, mask, st, en, tr)
    tags, lens = viterbi_c.viterbi(e.numpy(), mask.numpy(), st.numpy(), en.numpy(), tr.numpy())
    assert viterbi_c.to_lists(tags, lens) == want


def test_first_index_wins_on_exact_ties():
    T, S = 4, 5
    z = torch.zeros
    paths = crf_ref.viterbi_decode(z(2, S, T), None, z(T), z(T), z(T, T))
    assert paths == [[0] * S, [0] * S]
    tags, lens = viterbi_c.viterbi(np.zeros((2, S, T), np.float32), None, np.zeros(T), np.zeros(T), np.zeros((T, T)))
    assert viterbi_c.to_lists(tags, lens) == paths


def test_rounding_induced_tie():
    """Two predecessors 1 ulp apart tie after adding a large e[t][j]; the lower index must win."""
    T = 2
    st = torch.tensor([1.0, float(np.nextafter(np.float32(1.0), np.float32(2.0)))])
    en, tr = torch.zeros(T), torch.zeros(T, T)
    e = torch.zeros(1, 2, T); e[0, 1, :] = 1024.0
    paths = crf_ref.viterbi_decode(e, None, st, en, tr)
    assert paths[0][0] == 0                    # un-hoisted add rounds both to 1025 -> first index
    tags, lens = viterbi_c.viterbi(e.numpy(), None, st.numpy(), en.numpy(), tr.numpy())
    assert viterbi_c.to_lists(tags, lens) == paths


@pytest.mark.parametrize('T,S', [(2, 3), (3, 5), (4, 4)])
def test_log_partition_vs_enumeration(T, S):
    st, en, tr = P(T, 1, 'normal')
    e = synth.emissions(4, S, T, seed=8)
    for L in (1, S - 1, S):
        mask = torch.zeros(4, S, dtype=torch.bool); mask[:, :L] = True
        z = crf_ref.log_partition(e, mask, st, en, tr)
        for b in range(4):
            assert abs(float(z[b]) - crf_ref.brute_force_logZ(e[b, :L], L, st, en, tr)) < 1e-4


def test_llh_reductions_and_gradient():
    sh = synth.Shape(S=9, T=4)
    batch = synth.crf_batch(6, sh, seed=3, median_len=5)
    st, en, tr = [t.clone().requires_grad_(True) for t in P(sh.T, 2, 'normal')]
    e = batch['emissions'].clone().requires_grad_(True)
    none = crf_ref.log_likelihood(e, batch['tags'], batch['mask'], st, en, tr, 'none')
    assert none.shape == (6,) and (none <= 1e-5).all()
    tm = crf_ref.log_likelihood(e, batch['tags'], batch['mask'], st, en, tr, 'token_mean')
    assert torch.allclose(tm, none.sum() / batch['mask'].float().sum())
    assert torch.allclose(crf_ref.log_likelihood(e, batch['tags'], batch['mask'], st, en, tr, 'mean'), none.mean())
    (-tm).backward()
    # d(-llh)/d e = (marginals - one-hot gold) / n_tokens: rows of valid steps sum to 0, padded rows are 0
    gsum = e.grad.sum(-1)
    assert gsum.abs().max() < 1e-5
    assert e.grad[~batch['mask']].abs().max() == 0
    with pytest.raises(ValueError):
        crf_ref.log_likelihood(e, batch['tags'], batch['mask'], st, en, tr, 'bogus')


def test_validate_errors():
    e = torch.zeros(2, 3, 4)
    with pytest.raises(ValueError):
        crf_ref.validate(torch.zeros(2, 3), 4)
    with pytest.raises(ValueError):
        crf_ref.validate(e, 5)
    with pytest.raises(ValueError):
        crf_ref.validate(e, 4, tags=torch.zeros(2, 4, dtype=torch.long))
    with pytest.raises(ValueError):
        crf_ref.validate(e, 4, mask=torch.ones(3, 3, dtype=torch.bool))
    m = torch.ones(2, 3, dtype=torch.bool); m[1, 0] = False
    with pytest.raises(ValueError):
        crf_ref.validate(e, 4, mask=m)


@pytest.mark.parametrize('value', [float('-inf'), -10000.0])
def test_c_port_forbidden_cells(value):
    """-inf / -10000 emissions (forbidden tags; whole steps can be all -inf): every comparison is still ordered, the
    first maximal index wins, and the C port follows the torch restatement (torch.max) exactly."""
    sh = synth.Shape(S=30, T=15)
    batch = synth.crf_batch(41, sh, seed=8, kind='ties', median_len=14)
    e = synth.forbid_cells(batch['emissions'], 3, value=value)
    e[0, 2, :] = value                      # a step with no allowed tag
    st, en, tr = P(sh.T, 6, 'normal')
    want = crf_ref.viterbi_decode(e, batch['mask'], st, en, tr)
    tags, lens = viterbi_c.viterbi(e.numpy(), batch['mask'].numpy(), st.numpy(), en.numpy(), tr.numpy())
    assert viterbi_c.to_lists(tags, lens) == want

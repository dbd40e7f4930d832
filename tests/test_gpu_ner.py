"""GPU parity of the tag post-processing + chunk-F1 kernel (icka_ner_chunk_counts) against the oracle restatement of
My_cross_attention.py:879-903 + ner_evaluate.py and the golden vectors produced by the reference's own file."""
import json
import os
import random

import pytest
import torch

from icka_b200 import CRF, ner
from oracle import ner_ref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ner_eval.json')


def cases():
    return json.load(open(GOLDEN))['cases']


@pytest.mark.parametrize('case', cases(), ids=lambda c: c['name'])
def test_evaluate_matches_reference_golden(case):
    got = ner.evaluate(case['pred'], case['gold'], None, None, None, case['tags'])
    assert list(map(float, got)) == case['evaluate']


def same(a, b):
    """tuple equality where nan == nan (np.mean([]) of the reference when no token is kept)"""
    return all(x == y or (x != x and y != y) for x, y in zip(a, b)) and len(a) == len(b)


def random_batch(rng, B, S, p_same, prefix=True):
    gold = torch.zeros(B, S, dtype=torch.int64)
    pred = torch.full((B, S), -1, dtype=torch.int32)
    mask = torch.zeros(B, S, dtype=torch.uint8)
    for b in range(B):
        n = rng.randint(1, S)
        mask[b, :n] = 1
        if not prefix and n > 4 and rng.random() < 0.3:
            mask[b, rng.randint(1, n - 1)] = 0                  # a hole: the driver loop stops there
        body = [rng.choice([1, 1, 1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]) for _ in range(n)]
        body[0] = 13
        body[-1] = 14
        gold[b, :n] = torch.tensor(body)
        pred[b, :n] = torch.tensor([x if rng.random() < p_same else rng.randint(0, 14) for x in body],
                                   dtype=torch.int32)
    return pred, gold, mask


@pytest.mark.parametrize('B,S,p_same,prefix', [(64, 128, 0.8, True), (37, 33, 0.5, False), (5, 1, 0.5, True),
                                               (300, 256, 0.9, False), (16, 31, 0.2, True), (9, 64, 1.0, True)])
def test_counts_match_oracle(B, S, p_same, prefix):
    rng = random.Random(B * 1000 + S)
    pred, gold, mask = random_batch(rng, B, S, p_same, prefix)
    ev = ner.ChunkF1(device='cuda')
    per = ev.update(pred.cuda(), gold.cuda(), mask.cuda(), per_sentence=True).cpu()
    y_pred, y_true = ner_ref.filter_tokens(pred.tolist(), gold.tolist(), mask.tolist())
    tags = ner_ref.tag_dict()
    for b in range(B):
        want = ner_ref.counts([y_pred[b]], [y_true[b]], tags)
        assert tuple(per[b].tolist()) == want, f'sentence {b}'
    want = ner_ref.counts(y_pred, y_true, tags)
    assert ev.counts() == want
    assert same(ev.result(), ner_ref.scores(*want))
    # accumulation over batches and bool / int64 masks
    ev.update(pred.cuda(), gold.cuda(), mask.cuda().bool())
    ev.update(pred.cuda(), gold.cuda(), mask.cuda().long())
    assert ev.counts() == tuple(3 * x for x in want)
    ev.reset()
    assert ev.counts() == (0, 0, 0, 0, 0)


def test_viterbi_tags_to_f1_stay_on_device():
    """decode_tensors -> ChunkF1 with no host round trip equals decode() -> driver loop -> evaluate on the host."""
    torch.manual_seed(7)
    B, S, T = 96, 128, 15
    rng = random.Random(11)
    _, gold, mask = random_batch(rng, B, S, 1.0)
    em = torch.randn(B, S, T) * 2
    for b in range(B):                                          # make the gold path likely so chunks overlap
        n = int(mask[b].sum())
        em[b, torch.arange(n), gold[b, :n]] += 3.0
    crf = CRF(T, batch_first=True).cuda()
    tags_dev, _ = crf.decode_tensors(em.cuda(), mask.cuda().bool())
    ev = ner.ChunkF1(device='cuda')
    ev.update(tags_dev, gold.cuda(), mask.cuda())
    lists = crf.decode(em.cuda(), mask.cuda().bool())
    padded = [row + [0] * (S - len(row)) for row in lists]
    y_pred, y_true = ner_ref.filter_tokens(padded, gold.tolist(), mask.tolist())
    want = ner_ref.evaluate(y_pred, y_true, ner_ref.tag_dict())
    assert ev.result() == want
    assert want[1] > 0.2                                        # a non-trivial F1, not the all-zero corner


def test_errors():
    ev = ner.ChunkF1(device='cuda')
    pred = torch.zeros(2, 4, dtype=torch.int32, device='cuda')
    with pytest.raises(ValueError):
        ev.update(pred, torch.zeros(2, 5, dtype=torch.int64, device='cuda'))
    with pytest.raises(RuntimeError):
        ev.update(pred.cpu(), torch.zeros(2, 4, dtype=torch.int64))
    ev.update(pred, torch.full((2, 4), 99, dtype=torch.int64, device='cuda'))
    with pytest.raises(KeyError):
        ev.counts()
    with pytest.raises(RuntimeError):
        ner.ChunkF1(device='cpu')

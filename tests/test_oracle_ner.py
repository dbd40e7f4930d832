"""Oracle for the tag post-processing + chunk-F1 row against the golden vectors generated from
the reference's own ner_evaluate.py (oracle/make_golden_ner.py), plus the host-side label table of icka_b200.ner."""
import json
import os

import pytest

from icka_b200 import ner
from oracle import ner_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ner_eval.json')


def cases():
    return json.load(open(GOLDEN))['cases']


@pytest.mark.parametrize('case', cases(), ids=lambda c: c['name'])
def test_oracle_matches_reference_golden(case):
    tags = case['tags']
    for seq, want in zip(case['pred'], case['chunks_pred']):
        assert [list(t) for t in ner_ref.get_chunks(seq, tags)] == want
    for seq, want in zip(case['gold'], case['chunks_gold']):
        assert [list(t) for t in ner_ref.get_chunks(seq, tags)] == want
    got = ner_ref.evaluate(case['pred'], case['gold'], tags)
    assert list(map(float, got)) == case['evaluate']            # same integers -> same floats, bit for bit


def test_known_answers_of_the_reference_file():
    # ner_evaluate.py:13-17 (docstring) and :153-170 (__main__ block)
    tags = {'B-PER': 4, 'I-PER': 5, 'B-LOC': 3, 'O': 0}
    assert ner_ref.get_chunks([4, 5, 0, 3], tags) == [('PER', 0, 2), ('LOC', 3, 4)]
    c = [c for c in cases() if c['name'] == 'main_block'][0]
    assert c['evaluate'] == [0.85, 0.5714285714285715, 0.5, 0.6666666666666666]


def test_filter_tokens_follows_the_driver_loop():
    gold = [[13, 1, 4, 10, 5, 14, 0, 0], [13, 1, 14, 0, 0, 0, 0, 0]]
    pred = [[1, 1, 4, 5, 5, 1, 1, 1], [1, 2, 3, 4, 5, 6, 7, 8]]
    mask = [[1, 1, 1, 1, 1, 1, 0, 1], [1, 1, 1, 0, 0, 0, 0, 0]]
    y_pred, y_true = ner_ref.filter_tokens(pred, gold, mask)
    assert y_true == [[1, 4, 5], [1]] and y_pred == [[1, 4, 5], [2]]


def test_label_info_table():
    tags = ner.tag_dict()
    info = ner.label_info(tags, ner.SKIP_LABELS)
    assert len(info) == 15
    assert info[tags['O']] & ner.NER_OUTSIDE and not info[tags['B-PER']] & ner.NER_OUTSIDE
    assert info[tags['B-PER']] & ner.NER_BEGIN and not info[tags['I-PER']] & ner.NER_BEGIN
    assert (info[tags['B-PER']] & 0xff) == (info[tags['I-PER']] & 0xff) != (info[tags['B-LOC']] & 0xff)
    for name in ner.SKIP_LABELS:
        assert info[tags[name]] & ner.NER_SKIP
    assert not info[tags['PAD']] & ner.NER_SKIP
    with pytest.raises(KeyError):
        ner.label_info({'B-PER': 1})
    assert ner.scores(20, 17, 4, 8, 6) == ner_ref.scores(20, 17, 4, 8, 6)
    assert ner.scores(5, 0, 0, 3, 2)[1:] == (0, 0, 0)

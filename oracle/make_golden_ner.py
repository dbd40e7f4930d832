"""Generate tests/golden/ner_eval.json from the reference's OWN ner_evaluate.py (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_ner          (build container only: needs /root/reference)

Cases: the two known-answer calls of SURVEY section 4 (the docstring example of get_chunks and the arrays of
ner_evaluate.py:153-170) plus seeded random label sequences over ICKA's 14-label set (with the pad id 0 and the
special labels allowed among the predictions).  Stored: inputs, get_chunks output per sequence, evaluate() result.
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

import ner_evaluate as ref          # noqa: E402  (imports only numpy / codecs)
from oracle import ner_ref          # noqa: E402


def run_eval(pred, gold, tags):
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:      # evaluate() writes ./test_results.txt
        os.chdir(tmp)
        try:
            words = [['w'] * len(g) for g in gold]
            acc, f1, p, r = ref.evaluate(pred, gold, [[str(x) for x in s] for s in pred],
                                         [[str(x) for x in s] for s in gold], words, tags)
        finally:
            os.chdir(cwd)
    return [float(acc), float(f1), float(p), float(r)]


def main():
    cases = []
    tags0 = {'B-PER': 4, 'I-PER': 5, 'B-LOC': 3, 'O': 0}
    cases.append(dict(name='docstring', tags=tags0, pred=[[4, 5, 0, 3]], gold=[[4, 5, 0, 3]]))
    tags1 = {'0': 0, 'B-PER': 1, 'I-PER': 2, 'B-LOC': 3, 'I-LOC': 4, 'B-ORG': 5, 'I-ORG': 6, 'B-OTHER': 7, 'I-OTHER': 8, 'O': 9}
    cases.append(dict(name='main_block', tags=tags1,
                      pred=[[9, 9, 9, 1, 3, 1, 2, 2, 0, 0], [9, 9, 9, 1, 3, 1, 2, 0, 0, 0]],
                      gold=[[9, 9, 9, 9, 3, 1, 2, 2, 0, 0], [9, 9, 9, 9, 3, 1, 2, 2, 0, 0]]))
    tags = ner_ref.tag_dict()
    rng = random.Random(20260818)
    for k, (n_sent, max_len, p_same) in enumerate([(40, 30, 0.8), (25, 128, 0.6), (30, 12, 0.95), (10, 1, 0.5)]):
        gold, pred = [], []
        for _ in range(n_sent):
            n = rng.randint(1, max_len)
            g = [rng.choice([1, 1, 1, 2, 3, 4, 5, 6, 7, 8, 9]) for _ in range(n)]
            p = [x if rng.random() < p_same else rng.randint(0, 14) for x in g]
            gold.append(g)
            pred.append(p)
        cases.append(dict(name=f'random_{k}', tags=tags, pred=pred, gold=gold))
    for c in cases:
        c['chunks_pred'] = [[list(t) for t in ref.get_chunks(s, c['tags'])] for s in c['pred']]
        c['chunks_gold'] = [[list(t) for t in ref.get_chunks(s, c['tags'])] for s in c['gold']]
        c['evaluate'] = run_eval(c['pred'], c['gold'], c['tags'])
    out = os.path.join(ROOT, 'tests', 'golden', 'ner_eval.json')
    json.dump(dict(source='/root/reference/ner_evaluate.py get_chunks + evaluate', cases=cases), open(out, 'w'))
    print('wrote', out, [(c['name'], c['evaluate']) for c in cases[:2]])


if __name__ == '__main__':
    main()

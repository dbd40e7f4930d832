"""Tag post-processing and chunk-F1 on the device (SURVEY 8f "next" row 2).

Host-side mirror of what the reference does with the decoded tags:

  * the driver's filter loops, ``My_cross_attention.py:879-903`` (dev) and ``:1052-1077`` (test): walk each sentence
    while its mask is on and keep the positions whose GOLD label is not X / <s> / </s> / [CLS] / [SEP];
  * ``ner_evaluate.get_chunks`` / ``ner_evaluate.evaluate`` (``ner_evaluate.py:4-48``, ``:64-110``).

The B x S Python loops and per-token ``.item()``-style reads become ONE kernel over the device-resident tag tensor
that ``CRF.decode_tensors`` produced (``icka_ner_chunk_counts``), accumulating six exact integer counters; the only
host traffic is reading those 48 bytes once per evaluation.  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .ops import _ctx, _need

LABEL_LIST = ["O", "B-MISC", "I-MISC", "B-PER", "I-PER", "B-ORG", "I-ORG", "B-LOC", "I-LOC", "X", "[CLS]", "[SEP]",
              "<s>", "</s>"]                                    # MNERProcessor.get_labels, My_cross_attention.py:215
SKIP_LABELS = ("X", "</s>", "<s>", "[CLS]", "[SEP]")            # My_cross_attention.py:890-892

NER_SKIP, NER_OUTSIDE, NER_BEGIN = 1 << 8, 1 << 9, 1 << 10      # include/icka_b200.h


def tag_dict(label_list: Sequence[str] = LABEL_LIST) -> Dict[str, int]:
    """``reverse_label_map`` of My_cross_attention.py:911-912: labels from 1, 'PAD' = 0."""
    d = {label: i for i, label in enumerate(label_list, 1)}
    d['PAD'] = 0
    return d


def label_info(tags: Dict[str, int], skip: Sequence[str] = ()) -> List[int]:
    """Per label id: chunk-type id | SKIP | OUTSIDE | BEGIN, from a {tag name: id} dictionary (get_chunks' ``tags``).

    Follows ``get_chunk_type`` (ner_evaluate.py:50-62): class = name.split('-')[0], type = name.split('-')[-1];
    ids that no tag maps to get a private type (the reference would raise ``KeyError`` on them).
    """
    if 'O' not in tags:
        raise KeyError('O')                                     # get_chunks: default = tags['O']
    n_ids = max(tags.values()) + 1
    if min(tags.values()) < 0 or n_ids > 256:
        raise ValueError(f'label ids must lie in [0, 256), got {min(tags.values())}..{n_ids - 1}')
    idx_to_tag = {idx: tag for tag, idx in tags.items()}
    type_ids: Dict[str, int] = {}
    info = []
    for i in range(n_ids):
        name = idx_to_tag.get(i)
        if name is None:
            name = f'\0unmapped-{i}'
        tag_class, tag_type = name.split('-')[0], name.split('-')[-1]
        v = type_ids.setdefault(tag_type, len(type_ids))
        if v > 255:
            raise ValueError('more than 256 chunk types')
        if name in skip:
            v |= NER_SKIP
        if i == tags['O']:
            v |= NER_OUTSIDE
        if tag_class == 'B':
            v |= NER_BEGIN
        info.append(v)
    return info


def scores(n_tok: int, n_ok: int, correct: int, n_pred: int, n_gold: int) -> Tuple[float, float, float, float]:
    """(acc, f1, p, r) exactly as ner_evaluate.py:104-107 forms them."""
    p = correct / n_pred if correct > 0 else 0
    r = correct / n_gold if correct > 0 else 0
    f1 = 2 * p * r / (p + r) if correct > 0 else 0
    acc = n_ok / n_tok if n_tok else float('nan')               # np.mean([]) is nan in the reference
    return acc, f1, p, r


class ChunkF1:
    """Streaming evaluator: ``update`` per batch on the device, ``result`` once at the end.

    ``label_list`` as ``MNERProcessor.get_labels()``; ids are ``enumerate(label_list, 1)`` with 0 = 'PAD'.
    """

    def __init__(self, label_list: Sequence[str] = LABEL_LIST, device='cuda', skip: Sequence[str] = SKIP_LABELS,
                 tags: Optional[Dict[str, int]] = None):
        self.tags = dict(tags) if tags is not None else tag_dict(label_list)
        info = label_info(self.tags, skip)
        self.n_ids = len(info)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('icka_b200.ner needs a CUDA device (there is no CPU fallback)')
        self.info = torch.tensor(info, dtype=torch.int16).to(self.device)     # < 2^11: same bits as the ABI's u16
        self.totals = torch.zeros(6, dtype=torch.int64, device=self.device)

    def reset(self) -> None:
        self.totals.zero_()

    def update(self, pred_tags: torch.Tensor, label_ids: torch.Tensor, mask: Optional[torch.Tensor] = None,
               per_sentence: bool = False) -> Optional[torch.Tensor]:
        """pred_tags [B,S] int32 (``CRF.decode_tensors``), label_ids [B,S] int64, mask [B,S] (bool / u8 / int).

        Returns the per-sentence counters [B,5] int32 (tokens, correct tokens, |gold & pred|, |pred|, |gold|) when
        ``per_sentence`` is set.  No host synchronisation.
        """
        _need(pred_tags, torch.int32, 'ner.update(pred_tags)')
        gold = label_ids.long().contiguous()
        if pred_tags.dim() != 2 or gold.shape != pred_tags.shape:
            raise ValueError(f'pred_tags and label_ids must both be [B,S], got {tuple(pred_tags.shape)} and '
                             f'{tuple(gold.shape)}')
        m = None
        if mask is not None:
            if mask.shape != pred_tags.shape:
                raise ValueError(f'mask must be [B,S], got {tuple(mask.shape)}')
            m = mask.contiguous().view(torch.uint8) if mask.dtype in (torch.bool, torch.uint8) \
                else (mask != 0).contiguous().view(torch.uint8)
        B, S = pred_tags.shape
        lib, h, st = _ctx(pred_tags)
        out = torch.empty(B, 5, dtype=torch.int32, device=pred_tags.device) if per_sentence else None
        _lib.check(lib.icka_ner_chunk_counts(h, pred_tags.data_ptr(), gold.data_ptr(), None if m is None else m.data_ptr(),
                                             self.info.data_ptr(), self.n_ids, self.totals.data_ptr(),
                                             None if out is None else out.data_ptr(), B, S, st),
                   'icka_ner_chunk_counts')
        return out

    def counts(self) -> Tuple[int, int, int, int, int]:
        t = self.totals.tolist()                                # the one D2H read
        if t[5]:
            raise KeyError(f'{t[5]} label ids outside [0, {self.n_ids}) (the reference raises KeyError in label_map)')
        return tuple(t[:5])

    def result(self) -> Tuple[float, float, float, float]:
        """(acc, f1, p, r) -- the tuple ``ner_evaluate.evaluate`` returns."""
        return scores(*self.counts())


def evaluate(labels_pred_id: Sequence[Sequence[int]], labels_id: Sequence[Sequence[int]], labels_pred=None,
             labels=None, words=None, tags: Optional[Dict[str, int]] = None, device='cuda'):
    """Drop-in for ``ner_evaluate.evaluate(labels_pred_id, labels_id, labels_pred, labels, words, tags)`` on
    already-filtered id lists (ner_evaluate.py:64-110) -> (acc, f1, p, r).  The string / word arguments only fed the
    reference's ``./test_results.txt`` dump and are ignored.  Lists are padded into one [B,S] upload."""
    if tags is None:
        raise TypeError("evaluate() missing required argument: 'tags'")
    if len(labels_pred_id) != len(labels_id):
        raise ValueError('labels_pred_id and labels_id must hold the same number of sentences')
    B = len(labels_id)
    S = max([1] + [len(s) for s in labels_id])
    pred = torch.zeros(B, S, dtype=torch.int32)
    gold = torch.zeros(B, S, dtype=torch.int64)
    mask = torch.zeros(B, S, dtype=torch.uint8)
    for b, (p, g) in enumerate(zip(labels_pred_id, labels_id)):
        if len(p) != len(g):
            raise ValueError(f'sentence {b}: {len(p)} predicted vs {len(g)} gold labels')
        n = len(g)
        if n:
            pred[b, :n] = torch.as_tensor(list(p), dtype=torch.int32)
            gold[b, :n] = torch.as_tensor(list(g), dtype=torch.int64)
            mask[b, :n] = 1
    ev = ChunkF1(tags=tags, skip=(), device=device)
    if B:
        ev.update(pred.to(ev.device), gold.to(ev.device), mask.to(ev.device))
    return ev.result()

"""icka_b200 -- B200 (sm_100a) drop-in for ICKA's cross-modal fusion + CRF hot path.

Public surface (mirrors /root/reference/Cross_Modal_Interaction_Module.py and torchcrf):
    modules.BertCrossEncoder, BertCrossAttentionLayer, BertCrossAttention, BertCoAttention,
    BertSelfOutput, BertIntermediate, BertOutput, BertLayerNorm, cls_layer_both, CrossModalFusion
    crf.CRF
    model.MTCCMBertForMMTokenClassificationCRF (the top-level class, CMIM:886-1057: same ctor / 19-argument forward, 'dev' / 'test')
    emission.LSTM, emission.EmissionHead (the BiLSTM + classifier between fusion and CRF, CMIM:905-910, 1042-1043)
    prompt.PromptMapping (prompt mapping networks + prefix assembly, CMIM:913-930, 995-1009)
    resnet_tail.myResnet (region-producer tail: means, adaptive pooling, K-major region rows; resnet/resnet_utils.py)
    ner.ChunkF1 / ner.evaluate (tag post-processing + chunk-F1, My_cross_attention.py:879-903, ner_evaluate.py)
    set_precision('bf16' | 'fp32') (process default), `with precision(mode):` (thread-local override)
Everything computes through libicka_b200.so (include/icka_b200.h); there is no CPU fallback.
"""
from .config import FusionConfig  # noqa: F401
from .modules import (BertCoAttention, BertCrossAttention, BertCrossAttentionLayer, BertCrossEncoder,  # noqa: F401
                      BertIntermediate, BertLayerNorm, BertOutput, BertSelfOutput, CrossModalFusion,
                      cls_layer_both)
from .precision import get_precision, precision, set_precision  # noqa: F401
from .crf import CRF  # noqa: F401
from .emission import LSTM, EmissionHead  # noqa: F401
from .prompt import PromptMapping  # noqa: F401
from .resnet_tail import myResnet  # noqa: F401
from .model import MTCCMBertForMMTokenClassificationCRF  # noqa: F401

__all__ = ['FusionConfig', 'BertCoAttention', 'BertCrossAttention', 'BertCrossAttentionLayer', 'BertCrossEncoder',
           'BertIntermediate', 'BertLayerNorm', 'BertOutput', 'BertSelfOutput', 'CrossModalFusion', 'cls_layer_both',
           'CRF', 'LSTM', 'EmissionHead', 'PromptMapping', 'myResnet', 'MTCCMBertForMMTokenClassificationCRF', 'get_precision', 'set_precision', 'precision']

"""Top-level drop-in: ``MTCCMBertForMMTokenClassificationCRF`` (Cross_Modal_Interaction_Module.py:886-1057).

Same constructor ``(config, embedding, last_encoder, layer_num1=1, layer_num2=1, layer_num3=1, num_labels=2)`` and the
same 19-argument ``forward`` with ``mode`` in {'train', 'dev', 'test'} (returns ``loss`` / ``(pred_tags, loss)`` /
``pred_tags``, CMIM:1046-1057).  The two
transformer encoders are the caller's own torch modules (``embedding`` = ``self.bert``, ``last_encoder``; outside the
hot path); everything between them and the tag lists runs on libicka_b200.so:

    region projection, text->image and image->text cross encoders        CrossModalFusion.encode      CMIM:954-989
    prompt mapping networks + prefix assembly                            PromptMapping                CMIM:995-1009
    gate + blend                                                         CrossModalFusion.blend       CMIM:1029-1036
    BiLSTM + classifier                                                  EmissionHead                 CMIM:1042-1043
    CRF decode / negative log-likelihood                                 CRF                          CMIM:1045-1057

Parameter names are the reference's (``vismap2text.*``, ``txt2img_attention.*``, ``cls_layer_Y.*``, ``cls_layer.*``,
``aux_head.*``, ``lstm.*``, ``classifier.*``, ``crf.*``, ``mapping_network_alignment.*``, ``mapping_network_vision.*``,
``lastproj.*``), so ``model.load_state_dict(torch.load(path)['net'], False)`` (My_cross_attention.py:997-998) works.  The
members the reference constructs but never uses (``self_attention``, ``self_attention_v2``, ``embedding_layer``,
``LayerNorm``; CMIM:894-895, 903, 935) are not created -- their checkpoint keys are ignored by the non-strict load.
``mode='train'`` (the reference's main call, My_cross_attention.py:814-817) builds the autograd graph out of the
kernel-backed nodes of ``icka_b200.autograd`` (cross layers, dense layers of the prompt networks, gate + blend, BiLSTM with
BPTT, classifier, CRF log-likelihood); gradients flow through the caller's two encoders by ordinary PyTorch autograd.
"""
from __future__ import annotations

import torch
from torch import nn

from .crf import CRF
from .emission import LSTM, EmissionHead
from .modules import CrossModalFusion
from .prompt import PromptMapping


class MTCCMBertForMMTokenClassificationCRF(CrossModalFusion):
    def __init__(self, config, embedding, last_encoder, layer_num1=1, layer_num2=1, layer_num3=1, num_labels=2):
        super().__init__(config, layer_num1=layer_num1)          # layer_num2 / layer_num3 are ignored, as in CMIM:888
        self.num_labels = num_labels
        self.last_encoder = last_encoder
        self.bert = embedding
        self.hidden_size = config.hidden_size
        self.dropout = nn.Dropout(config.hidden_dropout_prob)                               # CMIM:896
        self.lstm = LSTM(input_size=config.hidden_size, hidden_size=config.hidden_size, batch_first=True,
                         bidirectional=True)                                                # CMIM:905-908
        self.classifier = nn.Linear(config.hidden_size * 2, num_labels)                     # CMIM:910
        self.crf = CRF(num_tags=num_labels, batch_first=True)                               # CMIM:911-912
        self.prompt_len = 5                                                                 # CMIM:913
        pm = PromptMapping(config, prompt_len=self.prompt_len)
        self.mapping_network_alignment = pm.mapping_network_alignment                       # CMIM:914-920
        self.mapping_network_vision = pm.mapping_network_vision                             # CMIM:922-928
        self.lastproj = pm.lastproj                                                         # CMIM:930
        head = EmissionHead.__new__(EmissionHead)
        nn.Module.__init__(head)
        head.lstm, head.classifier = self.lstm, self.classifier
        # helpers that SHARE the registered submodules above; kept out of the module tree so the state_dict keeps the
        # reference's flat key names
        object.__setattr__(self, '_prompt', pm)
        object.__setattr__(self, '_head', head)

    def train(self, mode: bool = True):
        super().train(mode)
        self._prompt.train(mode)
        self._head.train(mode)
        return self

    def forward(self, input_ids, segment_ids, input_mask, ori_input_ids, ori_input_mask, ori_segment_ids,
                added_attention_mask, clip_features, visual_embeds_mean, visual_embeds_att, offsets, output_mask,
                rela_score, temp=None, temp_lamb=None, lamb=None, labels=None, negative_rate=None, mode=None):
        if mode not in ('train', 'dev', 'test'):
            return None             # the reference's if / elif chain falls through (CMIM:1046-1057)
        from .precision import precision
        # 'train' records the autograd graph (when autograd is enabled); 'dev' / 'test' never do (the reference wraps them
        # in torch.no_grad(), My_cross_attention.py:870, :1045)
        with torch.set_grad_enabled(torch.is_grad_enabled() and mode == 'train'), precision(self.precision):
            offset = offsets.tolist()[0]                                                        # CMIM:948
            sequence_output = self.bert(ori_input_ids, token_type_ids=ori_segment_ids,
                                        attention_mask=ori_input_mask)[0].float()               # CMIM:949-950
            sequence_output = self.dropout(sequence_output)                                     # CMIM:953
            cross_output_layer, clip = self.encode(sequence_output, visual_embeds_att, clip_features,
                                                   added_attention_mask, ori_input_mask)        # CMIM:954-989
            prefix_emb, prompt_mask = self._prompt(clip, visual_embeds_mean, input_mask)        # CMIM:995-1009
            roberta_encoder_output = self.last_encoder(
                input_ids=input_ids, token_type_ids=segment_ids, attention_mask=input_mask,
                prompt_embeddings=prefix_emb, input_mask=prompt_mask, offset=offset)[0]         # CMIM:1010-1013
            offset = offset - 2 + prefix_emb.size(1)                                            # CMIM:1022
            token_embedding = roberta_encoder_output[:, offset: offset + 128, :]                # CMIM:1024
            result = self.blend(cross_output_layer, token_embedding)                            # CMIM:1029-1036
            emissions = self._head(result)                                                      # CMIM:1042-1043
            output_mask = (output_mask != 0)                                                    # CMIM:1045
            if mode == 'train':
                return -self.crf(emissions, tags=labels, mask=output_mask, reduction='token_mean')   # CMIM:1047-1048
            pred_tags = self.crf.decode(emissions, mask=output_mask)                            # CMIM:1051 / 1056
            if mode == 'test':
                return pred_tags
            loss = -self.crf(emissions, tags=labels, mask=output_mask, reduction='token_mean')  # CMIM:1052-1053
            return pred_tags, loss

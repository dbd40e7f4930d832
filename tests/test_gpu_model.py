"""The top-level drop-in MTCCMBertForMMTokenClassificationCRF (CMIM:886-1057) with stub encoders, against the oracle chain
(fusion_ref -> prompt_ref -> stub last_encoder -> fusion_ref gate -> lstm_ref -> crf_ref): emissions within the bf16
gate, tags bit-exact given the emissions, loss within 1e-4."""
import pytest
import torch
from torch import nn

import icka_b200
from icka_b200 import synth
from oracle import crf_ref, fusion_ref, lstm_ref, prompt_ref

pytestmark = pytest.mark.gpu
H, T, S, L, OFFSET = 768, 15, 128, 140, 4


class StubEmbedding(nn.Module):
    """Stands in for the BERT `embedding` (self.bert): (ids, token_type_ids=, attention_mask=) -> (states,)."""

    def __init__(self):
        super().__init__()
        self.emb = nn.Embedding(50, H)

    def forward(self, ids, token_type_ids=None, attention_mask=None):
        return (nn.functional.layer_norm(self.emb(ids), (H,)),)


class StubLastEncoder(nn.Module):
    """Stands in for the RoBERTa `last_encoder`: splices the 10 prompt rows in place of 2 tokens (CMIM:1010-1022)."""

    def __init__(self):
        super().__init__()
        self.emb = nn.Embedding(50, H)
        self.proj = nn.Linear(1024, H)

    def forward(self, input_ids=None, token_type_ids=None, attention_mask=None, prompt_embeddings=None, input_mask=None,
                offset=None):
        e = self.emb(input_ids)
        p = self.proj(prompt_embeddings)
        out = torch.cat([e[:, :offset], p, e[:, offset + 2:]], dim=1) + 0.05 * p.mean(dim=1, keepdim=True)
        return (nn.functional.layer_norm(out, (H,)),)


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = icka_b200.get_precision()
    yield
    icka_b200.set_precision(prev)


def test_full_model_dev_and_test_modes():
    icka_b200.set_precision('bf16')
    torch.manual_seed(1)
    B = 5
    shape = synth.Shape(L=1)
    cfg = icka_b200.FusionConfig(hidden_size=H, num_attention_heads=12, intermediate_size=3072)
    model = icka_b200.MTCCMBertForMMTokenClassificationCRF(cfg, StubEmbedding(), StubLastEncoder(), 1, 1, 1, num_labels=T)
    model = model.eval()
    inp = synth.fusion_inputs(B, shape, seed=2)
    crf_b = synth.crf_batch(B, shape, seed=2)
    g = torch.Generator().manual_seed(3)
    input_ids = torch.randint(0, 50, (B, L), generator=g)
    ori_input_ids = torch.randint(0, 50, (B, S), generator=g)
    input_mask = torch.ones(B, L, dtype=torch.long)
    segment_ids = torch.zeros(B, L, dtype=torch.long)
    ori_segment_ids = torch.zeros(B, S, dtype=torch.long)
    vmean = inp['visual_embeds_att'].mean(3).mean(2)
    offsets = torch.full((B,), OFFSET, dtype=torch.long)
    output_mask, labels = crf_b['mask'].long(), crf_b['tags']

    # ---- oracle chain on the CPU (fp32 / fp64) with the same parameters
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        seq = model.bert(ori_input_ids, token_type_ids=ori_segment_ids, attention_mask=inp['text_mask'])[0].float()
        fparams = {k: v for k, v in sd.items() if k.split('.')[0] in ('vismap2text', 'vismapping', 'txt2img_attention',
                                                                      'cls_layer_Y', 'cls_layer', 'aux_head')}
        seg = lambda tok: fusion_ref.fusion_segment(seq, inp['visual_embeds_att'], inp['clip_features'], tok,
                                                    inp['img_mask'], inp['text_mask'], fparams, num_layers=1,
                                                    num_heads=12, layer_norm_eps=cfg.layer_norm_eps)
        first = seg(torch.zeros(B, S, H))                       # fused / clip do not depend on token_embedding
        pparams = {k: v for k, v in sd.items() if k.split('.')[0] in ('mapping_network_alignment',
                                                                      'mapping_network_vision', 'lastproj')}
        prefix, pmask = prompt_ref.prompt_prefix(first['clip'], vmean, input_mask, pparams)
        enc = model.last_encoder(input_ids=input_ids, token_type_ids=segment_ids, attention_mask=input_mask,
                                 prompt_embeddings=prefix, input_mask=pmask, offset=OFFSET)[0]
        off = OFFSET - 2 + prefix.size(1)
        tok = enc[:, off:off + 128, :]
        result = seg(tok)['result']
        lp = {k[5:]: v.double() for k, v in sd.items() if k.startswith('lstm.')}
        want_e = lstm_ref.emission_head(result.double(), lp, sd['classifier.weight'].double(), sd['classifier.bias'].double())

    # ---- the drop-in on the GPU
    model = model.cuda()
    dev = lambda t: t.cuda()
    args = [dev(input_ids), dev(segment_ids), dev(input_mask), dev(ori_input_ids), dev(inp['text_mask']), dev(ori_segment_ids),
            dev(inp['img_mask']), dev(inp['clip_features']), dev(vmean), dev(inp['visual_embeds_att']), dev(offsets),
            dev(output_mask), None]
    tags = model(*args, mode='test')
    tags2, loss = model(*args, labels=dev(labels), mode='dev')
    assert tags == tags2
    # emissions of the GPU path (same calls the forward makes) for the tag / loss checks
    with torch.no_grad():
        seq_g = model.bert(dev(ori_input_ids), token_type_ids=dev(ori_segment_ids), attention_mask=dev(inp['text_mask']))[0].float()
        fused, clip = model.encode(seq_g, dev(inp['visual_embeds_att']), dev(inp['clip_features']), dev(inp['img_mask']),
                                   dev(inp['text_mask']))
        prefix_g, pmask_g = model._prompt(clip, dev(vmean), dev(input_mask))
        enc_g = model.last_encoder(input_ids=dev(input_ids), token_type_ids=dev(segment_ids), attention_mask=dev(input_mask),
                                   prompt_embeddings=prefix_g, input_mask=pmask_g, offset=OFFSET)[0]
        em = model._head(model.blend(fused, enc_g[:, off:off + 128, :]))
    err = (em.double().cpu() - want_e).abs().max().item()
    print(f'full model bf16: max |d emissions| {err:.2e}, prefix {float((prefix_g.cpu() - prefix).abs().max()):.2e}')
    assert err <= 2e-2
    cp = {k[4:]: v for k, v in sd.items() if k.startswith('crf.')}
    mask = crf_b['mask'].bool()
    assert tags == crf_ref.viterbi_decode(em.cpu(), mask, cp['start_transitions'], cp['end_transitions'], cp['transitions'])
    want_loss = -crf_ref.log_likelihood(em.cpu(), labels, mask, cp['start_transitions'], cp['end_transitions'],
                                            cp['transitions'], reduction='token_mean')
    assert abs(float(loss) - float(want_loss)) <= 1e-4 * max(1.0, abs(float(want_loss)))
    assert model(*args, labels=dev(labels), mode='predict') is None          # the reference's if/elif chain falls through


def test_full_model_train_mode_loss_and_gradients():
    """mode='train' (the reference's main call, My_cross_attention.py:814-817; CMIM:1046-1048): the loss and the gradient
    of EVERY parameter -- fusion stack, prompt networks, gate, BiLSTM, classifier, CRF and, through ordinary autograd, the
    caller's two encoders -- against autograd through the fp64 oracle chain.  fp32 kernels, dropout off (p = 0 / eval)."""
    icka_b200.set_precision('fp32')
    torch.manual_seed(11)
    B = 2
    shape = synth.Shape(L=1)
    cfg = icka_b200.FusionConfig(hidden_size=H, num_attention_heads=12, intermediate_size=3072, hidden_dropout_prob=0.0,
                                 attention_probs_dropout_prob=0.0)
    model = icka_b200.MTCCMBertForMMTokenClassificationCRF(cfg, StubEmbedding(), StubLastEncoder(), 1, 1, 1, num_labels=T)
    model = model.eval()                                   # Dropout(0.3) of the prompt networks off; autograd stays on
    inp = synth.fusion_inputs(B, shape, seed=12)
    crf_b = synth.crf_batch(B, shape, seed=12)
    g = torch.Generator().manual_seed(13)
    input_ids = torch.randint(0, 50, (B, L), generator=g)
    ori_input_ids = torch.randint(0, 50, (B, S), generator=g)
    input_mask = torch.ones(B, L, dtype=torch.long)
    segment_ids = torch.zeros(B, L, dtype=torch.long)
    ori_segment_ids = torch.zeros(B, S, dtype=torch.long)
    vmean = inp['visual_embeds_att'].mean(3).mean(2)
    offsets = torch.full((B,), OFFSET, dtype=torch.long)
    output_mask, labels = crf_b['mask'].long(), crf_b['tags']

    # ---- fp64 oracle chain with autograd, parameters under the model's own state_dict names
    sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in model.state_dict().items()}
    ln = lambda x: nn.functional.layer_norm(x, (H,))
    seq = ln(sd['bert.emb.weight'][ori_input_ids])
    fparams = {k: v for k, v in sd.items() if k.split('.')[0] in ('vismap2text', 'vismapping', 'txt2img_attention',
                                                                  'cls_layer_Y', 'cls_layer', 'aux_head')}
    grid64, clip64 = inp['visual_embeds_att'].double(), inp['clip_features'].double()
    seg = lambda tok: fusion_ref.fusion_segment(seq, grid64, clip64, tok, inp['img_mask'], inp['text_mask'], fparams,
                                                num_layers=1, num_heads=12, layer_norm_eps=cfg.layer_norm_eps)
    first = seg(torch.zeros(B, S, H, dtype=torch.float64))
    pparams = {k: v for k, v in sd.items() if k.split('.')[0] in ('mapping_network_alignment', 'mapping_network_vision',
                                                                  'lastproj')}
    prefix, pmask = prompt_ref.prompt_prefix(first['clip'], vmean.double(), input_mask, pparams)
    e = sd['last_encoder.emb.weight'][input_ids]
    pr = prefix @ sd['last_encoder.proj.weight'].t() + sd['last_encoder.proj.bias']
    enc = ln(torch.cat([e[:, :OFFSET], pr, e[:, OFFSET + 2:]], dim=1) + 0.05 * pr.mean(dim=1, keepdim=True))
    off = OFFSET - 2 + prefix.size(1)
    result = fusion_ref.gate_blend(first['fused'], enc[:, off:off + 128, :], fparams)[0]
    lp = {k[5:]: v for k, v in sd.items() if k.startswith('lstm.')}
    em = lstm_ref.emission_head(result, lp, sd['classifier.weight'], sd['classifier.bias'])
    want_loss = -crf_ref.log_likelihood(em, labels, crf_b['mask'].bool(), sd['crf.start_transitions'],
                                        sd['crf.end_transitions'], sd['crf.transitions'], reduction='token_mean')
    want_loss.backward()

    # ---- the drop-in on the GPU
    model = model.cuda()
    dev = lambda t: t.cuda()
    args = [dev(input_ids), dev(segment_ids), dev(input_mask), dev(ori_input_ids), dev(inp['text_mask']), dev(ori_segment_ids),
            dev(inp['img_mask']), dev(inp['clip_features']), dev(vmean), dev(inp['visual_embeds_att']), dev(offsets),
            dev(output_mask), None]
    loss = model(*args, labels=dev(labels), mode='train')
    assert loss.dim() == 0 and loss.grad_fn is not None
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(want_loss)) <= 1e-4 * max(1.0, abs(float(want_loss)))
    bad = {}
    for k, v in model.named_parameters():
        want = sd[k].grad
        assert want is not None and v.grad is not None, k
        scale = float(want.abs().max())
        if k.endswith('key.bias'):
            scale = float(sd[k.replace('key.bias', 'query.bias')].grad.abs().max())
        err = float((v.grad.double().cpu() - want).abs().max())
        if not err <= 1e-3 * scale + 1e-9:
            bad[k] = (err, scale)
    assert not bad, bad
    # 'dev' right after must not record a graph (the reference calls it under torch.no_grad())
    _, dev_loss = model(*args, labels=dev(labels), mode='dev')
    assert dev_loss.grad_fn is None

"""The widened path end to end on the GPU (TaggingPipeline: fusion -> BiLSTM + classifier -> Viterbi -> chunk-F1)
against the oracle chain; also under CUDA-graph capture (the recurrent kernel is a cooperative launch)."""
import pytest
import torch

import icka_b200
from icka_b200 import synth
from icka_b200.pipeline import TaggingPipeline
from oracle import crf_ref, fusion_ref, lstm_ref, ner_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = icka_b200.get_precision()
    yield
    icka_b200.set_precision(prev)


def build(B, seed=5):
    shape = synth.Shape(L=1)
    pipe = TaggingPipeline(shape, 'cuda:0', 'bf16', seed=seed)
    host = pipe.make_host_batch(B, shape, seed, pin=False)
    labels = synth.crf_batch(B, shape, seed=seed)['tags']
    return shape, pipe, host, labels


def test_emissions_tags_and_f1_against_the_oracle_chain():
    B = 6
    shape, pipe, host, labels = build(B)
    d = pipe.to_device(host)
    emissions, tags, lens = pipe.step_tagging(d, labels.cuda())
    torch.cuda.synchronize()
    # oracle chain in fp32 on the CPU with the same parameters
    params = {k: v.detach().cpu() for k, v in pipe.fusion.state_dict().items()}
    want = fusion_ref.fusion_segment(host['text_states'], host['visual_embeds_att'], host['clip_features'],
                                     host['token_embedding'], host['img_mask'], host['text_mask'], params,
                                     num_layers=shape.L, num_heads=shape.heads, layer_norm_eps=shape.eps)
    lp = {k: v.detach().cpu().double() for k, v in pipe.head.lstm.named_parameters()}
    want_e = lstm_ref.emission_head(want['result'].double(), lp, pipe.head.classifier.weight.detach().cpu().double(),
                                    pipe.head.classifier.bias.detach().cpu().double())
    err = (emissions.cpu().double() - want_e).abs().max().item()
    print(f'tagging pipeline bf16: max |d emissions| {err:.2e}')
    assert err <= 2e-2
    # Viterbi is bit-exact GIVEN the emissions: decode the GPU's own emissions with the oracle
    cp = {k: v.detach().cpu() for k, v in pipe.crf.state_dict().items()}
    mask = host['crf_mask'].bool()
    want_tags = crf_ref.viterbi_decode(emissions.cpu(), mask, cp['start_transitions'], cp['end_transitions'],
                                       cp['transitions'])
    got_tags = [tags[b, :int(lens[b])].tolist() for b in range(B)]
    assert got_tags == want_tags
    padded = [row + [0] * (shape.S - len(row)) for row in want_tags]
    y_pred, y_true = ner_ref.filter_tokens(padded, labels.tolist(), host['crf_mask'].tolist())
    assert pipe.f1.counts() == ner_ref.counts(y_pred, y_true, ner_ref.tag_dict())


def test_cuda_graph_capture_of_the_widened_step():
    B = 130                                    # two sentence tiles, the second ragged
    shape, pipe, host, labels = build(B, seed=9)
    d = pipe.to_device(host)
    lab = labels.cuda()
    with torch.no_grad():
        for _ in range(2):
            eager = pipe.step_tagging(d)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        pipe.step_tagging(d)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            outs = pipe.step_tagging(d)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(outs[0], eager[0]) and torch.equal(outs[1], eager[1]) and torch.equal(outs[2], eager[2])
    del lab

"""Prompt mapping oracle (oracle/prompt_ref.py) against the golden outputs of the reference's own modules
(tests/golden/prompt_prefix.npz, oracle/make_golden_prompt.py) and, where /root/reference exists, live."""
import pytest
import torch

from oracle import prompt_ref, reference_shim

load_case = prompt_ref.load_golden_case


def test_oracle_matches_reference_golden():
    p, clip, vmean, mask, want, want_mask = load_case('std')
    got, got_mask = prompt_ref.prompt_prefix(clip, vmean, mask, p)
    assert got.shape == (3, 10, 1024)
    assert (got - want).abs().max().item() <= 1e-5
    assert torch.equal(got_mask, want_mask)


def test_state_dict_keys_match_the_reference_names():
    import icka_b200
    from icka_b200.prompt import PromptMapping
    m = PromptMapping(icka_b200.FusionConfig(hidden_size=768))
    assert set(m.state_dict()) == set(prompt_ref.make_params(768, seed=1))


@pytest.mark.skipif(not reference_shim.available(), reason='/root/reference not present')
def test_reference_class_has_these_members():
    cmim = reference_shim.load()
    src = open(cmim.__file__).read()
    for name in ('mapping_network_alignment', 'mapping_network_vision', 'lastproj', 'prompt_len = 5'):
        assert name in src

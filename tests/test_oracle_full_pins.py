"""Pin the oracle at the FULL BASELINE sizes against numbers recorded from the reference itself
(tests/golden/full_size_pins.json): weights are rebuilt from the seed (bit-identical to the reference's, see
test_host_logic.py), inputs from the seeded generator, and the oracle's output is compared with the
reference's sub-sampled output."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import describe as D
from oracle import separator_ref as R
from puresound_b200 import recipes, testing


def load_pin(name):
    """full_size_pins.json (the BASELINE configs), gated_pins.json (cfg4 with GatedTCN blocks) or
    round2_pins.json (cfg-1b and the benched cfg2 shape: the cfg1 model at batch 64, items 0 / 31 / 63 recorded)."""
    f = {"cfg4_gated": "gated_pins.json", "cfg1b": "round2_pins.json", "cfg2_b64": "round2_pins.json"}.get(name, "full_size_pins.json")
    with open(os.path.join(GOLDEN, f)) as fh:
        return json.load(fh)[name]


@pytest.mark.parametrize("name", ["cfg1b", "cfg2_b64"])
def test_oracle_matches_reference_round2_pins(name):
    """cfg-1b and cfg2 itself: the reference ran the whole 64-utterance batch; items never mix in eval
    mode, so the oracle is checked on the recorded items only (0 / 31 / 63 of the same seeded batch)."""
    pin = load_pin(name)
    torch.manual_seed(0)
    m = recipes.baseline_config("cfg1b" if name == "cfg1b" else "cfg2").eval()
    testing.perturb_(m, seed=1)
    assert m.overall_parameters == pin["params"]
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pin["state_checksum"], rel=1e-12)
    mix, _ = testing.noisy_speech(pin["batch"], pin["length"], seed=pin["input_seed"])
    y = R.inference(m.state_dict(), D.describe(m), mix[pin["items"]])
    assert y.shape[-1] == pin["out_len"]
    assert (y[:, :: pin["stride"]] - torch.tensor(pin["samples"])).abs().max().item() <= 2e-5
    assert abs(float(y.abs().mean()) - pin["out_abs_mean"]) <= 1e-6


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4", "cfg5_offline", "veve_dprnn_v0_causal", "cfg4_gated"])
def test_oracle_matches_reference_at_full_size(name):
    pin = load_pin(name)
    torch.manual_seed(0)
    m = recipes.baseline_config(name).eval()
    testing.perturb_(m, seed=1)
    mix, _ = testing.noisy_speech(pin["batch"], pin["length"], seed=pin["input_seed"])
    enr = testing.noisy_speech(pin["batch"], pin["enroll_length"], seed=pin["enroll_seed"])[0] if pin["enroll_length"] else None
    y = R.inference(m.state_dict(), D.describe(m), mix, enr)
    assert y.shape[-1] == pin["out_len"]
    got = y[:, :: pin["stride"]]
    want = torch.tensor(pin["samples"])
    assert (got - want).abs().max().item() <= 2e-5
    assert abs(float(y.abs().mean()) - pin["out_abs_mean"]) <= 1e-6


@pytest.mark.parametrize("name", ["tse_unet_tcn_v0", "tse_unet_tcn_v0_causal", "tse_unet_tcn_v1"])
def test_oracle_matches_reference_unet_recipes(name):
    """The reference's STFT-domain TSE recipes (egs/tse/model.py:184-369) at full size, 2 x (4 s + 6 s)."""
    with open(os.path.join(GOLDEN, "unet_pins.json")) as fh:
        pin = json.load(fh)[name]
    torch.manual_seed(0)
    m = recipes.init_model(name, verbose=False).eval()
    testing.perturb_(m, seed=1)
    assert m.overall_parameters == pin["params"]
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pin["state_checksum"], rel=1e-12)
    mix, _ = testing.noisy_speech(pin["batch"], pin["length"], seed=pin["input_seed"])
    enr = testing.noisy_speech(pin["batch"], pin["enroll_length"], seed=pin["enroll_seed"])[0]
    y = R.inference(m.state_dict(), D.describe(m), mix, enr)
    assert y.shape[-1] == pin["out_len"]
    assert (y[:, :: pin["stride"]] - torch.tensor(pin["samples"])).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", ["ns_dpcrn_v0", "ns_dpcrn_v0_causal", "ns_dparn_v0", "ns_dparn_v0_causal"])
def test_oracle_matches_reference_dpcrn_recipes(name):
    """The egs/ns noise-suppression recipes (egs/ns/model.py:38-216: complex mask on the STFT) at full size, 2 x 4 s."""
    with open(os.path.join(GOLDEN, "dparn_pins.json" if "dparn" in name else "dpcrn_pins.json")) as fh:
        pin = json.load(fh)[name]
    torch.manual_seed(0)
    m = recipes.init_model(name, verbose=False).eval()
    testing.perturb_(m, seed=1)
    assert m.overall_parameters == pin["params"] == (1215179 if "dparn" in name else 1380043)  # the counts the reference documents
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pin["state_checksum"], rel=1e-12)
    mix, _ = testing.noisy_speech(pin["batch"], pin["length"], seed=pin["input_seed"])
    y = R.inference(m.state_dict(), D.describe(m), mix, None)
    assert y.shape[-1] == pin["out_len"]
    assert (y[:, :: pin["stride"]] - torch.tensor(pin["samples"])).abs().max().item() <= 2e-5


def test_oracle_matches_reference_skim_recipe():
    """`tse_skim_v0_causal` (egs/tse/model.py:418-463, the reference's demo model) at full size."""
    with open(os.path.join(GOLDEN, "skim_pins.json")) as fh:
        pin = json.load(fh)["tse_skim_v0_causal"]
    torch.manual_seed(0)
    m = recipes.init_model("tse_skim_v0_causal", verbose=False).eval()
    testing.perturb_(m, seed=1)
    assert m.overall_parameters == pin["params"]
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pin["state_checksum"], rel=1e-12)
    mix, _ = testing.noisy_speech(pin["batch"], pin["length"], seed=pin["input_seed"])
    enr = testing.noisy_speech(pin["batch"], pin["enroll_length"], seed=pin["enroll_seed"])[0]
    y = R.inference(m.state_dict(), D.describe(m), mix, enr)
    assert y.shape[-1] == pin["out_len"]
    assert (y[:, :: pin["stride"]] - torch.tensor(pin["samples"])).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", ["tse_skim_v1_causal", "tse_skim_v2_causal", "tse_skim_v0_causal_vad"])
def test_oracle_matches_reference_mel_and_rnn_speaker_recipes(name):
    """`tse_skim_v1_causal` (bidirectional-LSTM speaker net) and `tse_skim_v2_causal` (mel front-end + SpecAugment, which the
    reference applies at inference too: the global seed recorded with the pin is set right before the call) at full size."""
    with open(os.path.join(GOLDEN, "mel_pins.json")) as fh:
        pin = json.load(fh)[name]
    torch.manual_seed(0)
    m = recipes.init_model(name, verbose=False).eval()
    testing.perturb_(m, seed=1)
    assert m.overall_parameters == pin["params"]
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pin["state_checksum"], rel=1e-9)
    mix, _ = testing.noisy_speech(pin["batch"], pin["length"], seed=pin["input_seed"])
    enr = testing.noisy_speech(pin["batch"], pin["enroll_length"], seed=pin["enroll_seed"])[0]
    torch.manual_seed(pin["rng_seed"])
    y = R.inference(m.state_dict(), D.describe(m), mix, enr)
    assert y.shape[-1] == pin["out_len"]
    assert (y[:, :: pin["stride"]] - torch.tensor(pin["samples"])).abs().max().item() <= 2e-5
    torch.manual_seed(pin["rng_seed"])
    emb = R.tse_embedding(m.state_dict(), D.describe(m), enr)
    assert (emb.flatten() - torch.tensor(pin["embedding"])).abs().max().item() <= 2e-5


def _real_inputs(g, tag):
    mix = (g["mix_i16"].float() / 32768.0)[None]
    enr = (g["enroll_i16"].float() / 32768.0)[None]
    if tag.startswith("white"):
        return testing.white(1, 64000, amp=1.0, seed=g["white_seed"]), None
    return mix, (enr if g["pins"][tag]["config"] in ("cfg4", "veve_dprnn_v0_causal") else None)


@pytest.mark.parametrize("tag", ["speech_cfg1", "speech_cfg3", "speech_cfg4", "speech_veve", "white_a1_cfg1"])
def test_oracle_matches_reference_on_real_speech_and_full_scale_noise(tag):
    """SURVEY.md 8d inputs (iii) — the reference's own two-speaker speech fixture, 4 s (+ 6 s enrollment) — and (i) at
    a = 1.0, where 58 % of the output samples sit on the [-1, 1] clamp; outputs recorded from the reference."""
    g = torch.load(os.path.join(GOLDEN, "real_input_pins.pt"))
    pin = g["pins"][tag]
    torch.manual_seed(0)
    m = recipes.baseline_config(pin["config"]).eval()
    testing.perturb_(m, seed=1)
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pin["state_checksum"], rel=1e-12)
    mix, enr = _real_inputs(g, tag)
    y = R.inference(m.state_dict(), D.describe(m), mix, enr)
    assert y.shape[-1] == pin["out_len"]
    assert (y[0, :: pin["stride"]] - torch.tensor(pin["samples"])).abs().max().item() <= 2e-5
    assert float((y.abs() >= 1).float().mean()) == pytest.approx(pin["out_clamped_frac"], abs=1e-4)

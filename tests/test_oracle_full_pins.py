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


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4", "cfg5_offline", "veve_dprnn_v0_causal"])
def test_oracle_matches_reference_at_full_size(name):
    with open(os.path.join(GOLDEN, "full_size_pins.json")) as fh:
        pin = json.load(fh)[name]
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

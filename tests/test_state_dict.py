"""CPU: the mirror modules expose the reference's state_dict keys and shapes (checkpoints are interchangeable) and
the reference's CLI flags / defaults.  Fixture: tests/golden/state_dict.json, generated from the unmodified reference."""
import json
import os

import pytest

from normalizing_flows_dpfs_b200.arguments import parse_args
from normalizing_flows_dpfs_b200.DPFs import DPF

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("meas", ["gaussian", "cos", "CRNVP", "NN"])
def test_state_dict_matches_reference(meas):
    ref = json.load(open(os.path.join(HERE, "golden", "state_dict.json")))[meas]
    mine = {k: list(v.shape) for k, v in DPF(parse_args(["--measurement", meas])).state_dict().items()}
    assert list(mine.keys()) == list(ref.keys()) or set(mine) == set(ref), (set(mine) ^ set(ref))
    assert mine == ref
    n_params = sum(p.numel() for p in DPF(parse_args(["--measurement", meas])).parameters())
    assert n_params == {"gaussian": 1672168, "cos": 1672168, "CRNVP": 1677032}.get(meas, n_params)


def test_cli_defaults_match_reference():
    a = parse_args([])
    assert (a.measurement, a.resampler_type, a.alpha, a.epsilon, a.scaling, a.threshold, a.max_iter) == ("cos", "ot", 0.5, 0.1, 0.75, 1e-3, 100)
    assert (a.num_particles, a.batchsize, a.hiddensize, a.sequence_length, a.pos_noise, a.width) == (100, 32, 32, 50, 20.0, 128)
    assert a.NF_dyn is False and a.NF_cond is False and a.e2e_train is True and a.gpu is True
    b = parse_args(["--NF-dyn", "--NF-cond", "--measurement", "CRNVP", "--num-particles", "7"])
    assert b.NF_dyn and b.NF_cond and b.measurement == "CRNVP" and b.num_particles == 7

"""GPU: the reference's own main.py (tests/golden/ref_main.py: a byte-for-byte copy of /root/reference/main.py kept as a test
fixture, see tests/golden/README.md) executed UNCHANGED on top of the dropin/ shims -- dataset -> DataLoader -> DPF(args).to(device) ->
train_val -> torch.save(dpf) -> testing (reference main.py:23-64)."""
import os
import runpy
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _toy_npz(path, n_seq, T, rng):
    def split(n):
        return {"start_image": rng.random((n, 128, 128, 3), dtype=np.float32), "start_state": rng.normal(0, 20, (n, 4)),
                "image": rng.random((n, T, 128, 128, 3), dtype=np.float32), "state": rng.normal(0, 20, (n, T, 4)),
                "q": rng.normal(0, 1, (n, T, 4)), "visible": np.ones((n, T), np.int64)}
    for name, n in (("train", n_seq), ("val", 50), ("test", 50)):     # main.py:51,63 use batch_size=50, drop_last=True
        np.savez(os.path.join(path, "toy_pn=2.0_d=25_const_%s.npz" % name), **{name + "_data": split(n)})


@pytest.mark.parametrize("flags", [["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"],
                                   ["--measurement", "CRNVP", "--resampler_type", "ot", "--fast"]])
def test_reference_main_runs_unchanged(tmp_path, monkeypatch, flags):
    monkeypatch.chdir(tmp_path)
    data = tmp_path / "data" / "disk" / "TwentyfiveDistractors"
    data.mkdir(parents=True)
    _toy_npz(str(data), 8, 3, np.random.default_rng(0))
    monkeypatch.setattr(sys, "argv", ["main.py", "--batchsize", "4", "--num-particles", "32", "--sequence-length", "3", "--num-epochs", "1"] + flags)
    drop = os.path.join(ROOT, "dropin")
    monkeypatch.syspath_prepend(drop)
    for name in ("DPFs", "dataset", "arguments", "losses", "utils", "model", "model.models", "nf", "nf.flows", "nf.models", "resamplers",
                 "resamplers.resamplers"):
        monkeypatch.delitem(sys.modules, name, raising=False)      # the reference's top-level module names must resolve to dropin/
    ns = runpy.run_path(os.path.join(ROOT, "tests", "golden", "ref_main.py"), run_name="__main__")
    dpf = ns["dpf"]
    assert type(dpf).__module__ == "normalizing_flows_dpfs_b200.DPFs"
    assert os.path.exists("model/dpf.pkl"), "torch.save(dpf, ...) of main.py:57"
    run = ns["run_id"]
    assert os.path.exists(os.path.join("logs", run, "models", "e2e_model_bestval_e2e.pth"))
    assert os.path.exists(os.path.join("logs", run, "data", "test_result.npz"))
    again = torch.load("model/dpf.pkl", weights_only=False)          # the pickled module loads back and still filters
    assert sorted(again.state_dict().keys()) == sorted(dpf.state_dict().keys())
    for a, b in zip(again.parameters(), dpf.parameters()):
        assert torch.equal(a, b)

"""N3 (supervised stop-head training): pins oracle/lstm_train_oracle.py against the UNMODIFIED reference
(PPOV2.1/train_lstm.py + model.py:67-91, container only), checks it against the committed golden run, and checks
the product's host-side plateau scheduler against torch's."""
import contextlib
import io
import random

import numpy as np
import pytest
import torch

from oracle import lstm_train_oracle as lo
from oracle.ref_harness import FakeNC, reference_available, reference_modules
from tests.helpers import load_golden

needs_ref = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@needs_ref
def test_dataset_matches_reference():
    nc = lo.synthetic_nc(40, seed=3)
    with reference_modules("2.1") as ref:
        ref.model.Dataset = lambda path, mode="r": FakeNC(nc)
        with contextlib.redirect_stdout(io.StringIO()):
            segments = ref.model.load_trajectory_segments("training_data.nc", tail_steps=60)
            random.seed(11)
            ds = ref.train_lstm.TrajectoryDataset(segments, window_size=20)
    elig = lo.eligible_episodes(nc, 20)
    assert 0 < len(elig) < 40                                    # some episodes are too short and are skipped
    sel = [int(elig[i]) for i in lo.select_episodes(len(elig), 11)]
    f, l = lo.build_dataset(nc, sel)
    assert len(ds) == len(l) == 2 * len(elig)
    ref_f = np.stack([ds[i][0].numpy() for i in range(len(ds))]).squeeze(-1)
    ref_l = np.stack([ds[i][1].numpy() for i in range(len(ds))])
    assert np.array_equal(ref_f, f) and np.array_equal(ref_l, l)
    assert 0.2 < l[1::2, 1].mean() < 0.8                         # both stop labels occur


@needs_ref
def test_training_loop_matches_reference(monkeypatch):
    """The reference's train() (100 epochs, CPU) under a seeded RNG vs the restatement: every checkpoint it
    saves and every epoch mean it prints must be reproduced."""
    nc = lo.synthetic_nc(70, seed=5)
    elig = lo.eligible_episodes(nc, 20)
    saved = []
    monkeypatch.setattr(torch, "save", lambda obj, path: saved.append({k: v.clone() for k, v in obj.items()}))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    out = io.StringIO()
    with reference_modules("2.1") as ref:
        ref.model.Dataset = lambda path, mode="r": FakeNC(nc)
        ref.train_lstm.load_trajectory_segments = ref.model.load_trajectory_segments
        random.seed(21)
        torch.manual_seed(21)
        with contextlib.redirect_stdout(out):
            ref.train_lstm.train()
    ref_lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("Epoch")]
    assert len(ref_lines) == 100

    sel = [int(elig[i]) for i in lo.select_episodes(len(elig), 21)]
    f, l = lo.build_dataset(nc, sel)
    torch.manual_seed(21)
    model = lo.PeakAndStopPredictor()
    res = lo.train(model, f, l, epochs=100)
    mine = [f"Epoch {e + 1:03d} | Loss: {avg:.4f} | LR: {lr_after:.2e}" for e, (avg, _, lr_after) in enumerate(res["history"])]
    assert mine == ref_lines
    assert len(saved) == len(res["checkpoints"]) > 3
    for a, b in zip(saved, res["checkpoints"]):
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_oracle_reproduces_golden_run():
    g = load_golden("lstm_train_s13.npz")
    nc = {k[3:]: g[k] for k in g.files if k.startswith("nc_")}
    f, l = lo.build_dataset(nc, g["selected"].tolist())
    assert np.array_equal(f, g["features"]) and np.array_equal(l, g["labels"])
    model = lo.PeakAndStopPredictor()
    model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init_")})
    loss0, grad0 = lo.loss_and_grad(model, torch.from_numpy(f[g["orders"][0][:64]]), torch.from_numpy(l[g["orders"][0][:64]]))
    assert np.allclose(loss0, g["batch_losses"][0], rtol=1e-6)
    assert np.allclose(grad0.numpy(), g["grad0"], rtol=1e-4, atol=1e-7)
    res = lo.train(model, f, l, epochs=len(g["orders"]), orders=g["orders"])
    assert np.allclose(res["batch_losses"], g["batch_losses"], rtol=1e-5)
    for k, v in res["final"].items():
        assert np.allclose(v.numpy(), g["final_" + k], rtol=1e-4, atol=1e-6), k


def test_plateau_scheduler_matches_torch():
    import uav_wrf_les_ppo_lstm_b200 as pb
    rng = np.random.default_rng(0)
    for trial in range(20):
        losses = np.abs(np.cumsum(rng.normal(size=80) * 0.01) + 1.0 - 0.01 * np.arange(80) * (trial % 3 == 0))
        losses[30:] = losses[30]            # a plateau
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=1e-3)
        ts = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", patience=5)
        mine = pb.lstm_train.ReduceLROnPlateau(1e-3, patience=5)
        for v in losses:
            ts.step(float(v))
            assert mine.step(float(v)) == opt.param_groups[0]["lr"]
        assert mine.lr < 1e-3

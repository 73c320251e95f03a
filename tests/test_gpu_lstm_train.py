"""N3: supervised stop-head training on the GPU (csrc/lstm_train_kernels.cu behind LstmTrainer) against the
torch-CPU restatement of PPOV2.1/train_lstm.py (oracle/lstm_train_oracle.py, pinned bit-equal against the
reference's own train()) and its committed golden run.  Tolerances: losses fp32 rel 1e-5 on the first
minibatch (north_star bar), 2e-4 after 24 optimiser steps (drift of a chaotic recursion: the two float32
implementations round differently); gradients within 1e-5 of the largest entry."""
import numpy as np
import pytest
import torch

from oracle import lstm_train_oracle as lo
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _golden():
    g = load_golden("lstm_train_s13.npz")
    nc = {k[3:]: g[k] for k in g.files if k.startswith("nc_")}
    init = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init_")}
    return g, nc, init


def _model(init):
    import uav_wrf_les_ppo_lstm_b200 as m
    model = m.PeakAndStopPredictor(device="cuda")
    model.load_state_dict(init)
    return model


def test_dataset_kernel_matches_reference_dataset():
    import uav_wrf_les_ppo_lstm_b200 as m
    g, nc, _ = _golden()
    elig = m.eligible_episodes((~np.isnan(nc["x"])).sum(axis=1), 20)
    assert np.array_equal(elig, lo.eligible_episodes(nc, 20))
    f, l = m.build_dataset(nc, g["selected"])
    assert np.array_equal(f.cpu().numpy(), g["features"])
    assert np.array_equal(l.cpu().numpy(), g["labels"])


def test_first_minibatch_loss_and_gradient():
    import uav_wrf_les_ppo_lstm_b200 as m
    g, nc, init = _golden()
    model = _model(init)
    f, l = torch.from_numpy(g["features"]).cuda(), torch.from_numpy(g["labels"]).cuda()
    tr = m.LstmTrainer(model, f[g["orders"][0][:64]], l[g["orders"][0][:64]], batch_size=64)
    before = tr.flat.clone()
    tr.train_epoch(order=np.arange(64))
    loss = float(tr.batch_losses[0].item())
    assert abs(loss - g["batch_losses"][0]) <= 1e-5 * abs(g["batch_losses"][0])          # fp32 rel 1e-5
    grad = tr.grad.cpu().numpy()
    scale = np.abs(g["grad0"]).max()
    assert np.abs(grad - g["grad0"]).max() <= 1e-5 * scale
    assert abs(float(tr.grad_norms[0].item()) - g["grad_norms"][0]) <= 1e-5 * g["grad_norms"][0]
    # one AdamW step with a clipped gradient moved every parameter by about lr
    step = (tr.flat - before).abs()
    assert 0.5e-3 < float(step.max().item()) < 1.2e-3
    # the nn.Parameters are views of the flat buffer: state_dict() sees the update
    sd = model.state_dict()
    assert torch.equal(sd["lstm.weight_hh_l0"].reshape(-1), tr.flat[128:128 + 4096])
    assert torch.equal(sd["fc_stop.0.bias"], tr.flat[4545:4546])


def test_training_run_follows_the_golden_run():
    import uav_wrf_les_ppo_lstm_b200 as m
    g, nc, init = _golden()
    model = _model(init)
    f, l = m.build_dataset(nc, g["selected"])
    tr = m.LstmTrainer(model, f, l, batch_size=64)
    assert tr.n == 148 and tr.n_batches == 3                     # last minibatch is ragged (20 samples)
    losses = []
    for e in range(len(g["orders"])):
        tr.train_epoch(order=g["orders"][e])
        losses.append(tr.batch_losses.cpu().numpy().copy())
    losses = np.concatenate(losses)
    assert np.allclose(losses, g["batch_losses"], rtol=2e-4)
    assert np.allclose(losses[:3], g["batch_losses"][:3], rtol=2e-5)
    assert np.allclose([h[0] for h in tr.history], g["epoch_means"], rtol=1e-4)
    for k, v in model.state_dict().items():
        assert np.allclose(v.cpu().numpy(), g["final_" + k], atol=3e-5, rtol=1e-3), k
    assert tr.best_loss == min(h[0] for h in tr.history) and tr.best_state is not None
    # the trained head through the inference kernel == the oracle model with the same weights
    ref = lo.PeakAndStopPredictor()
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    with torch.no_grad():
        rp, rs = ref(f.cpu())
    p, s = model(f)
    assert np.allclose(p.cpu().numpy(), rp.numpy(), rtol=1e-4, atol=1e-5)
    assert np.allclose(s.cpu().numpy(), rs.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n,batch", [(1, 64), (7, 7), (9, 8), (130, 64), (300, 128)])
def test_ragged_and_odd_batches_against_torch(n, batch):
    """Edge cases of the minibatch split (single sample, batch not a multiple of the 8 sequences per CTA, short
    last minibatch, batch larger than the reference's 64) against the torch-CPU loop on random windows."""
    import uav_wrf_les_ppo_lstm_b200 as m
    rng = np.random.default_rng(n)
    feats = rng.random((n, 20)).astype(np.float32)
    labels = np.stack([rng.random(n), (rng.random(n) < 0.4).astype(np.float64)], axis=1).astype(np.float32)
    torch.manual_seed(n)
    ref = lo.PeakAndStopPredictor()
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    orders = [np.random.default_rng(e).permutation(n) for e in range(3)]
    res = lo.train(ref, feats, labels, epochs=3, batch_size=batch, orders=orders)
    model = _model(init)
    tr = m.LstmTrainer(model, torch.from_numpy(feats).cuda(), torch.from_numpy(labels).cuda(), batch_size=batch)
    got = []
    for e in range(3):
        tr.train_epoch(order=orders[e])
        got.append(tr.batch_losses.cpu().numpy().copy())
    assert np.allclose(np.concatenate(got), res["batch_losses"], rtol=5e-5)
    for k, v in model.state_dict().items():
        assert np.allclose(v.cpu().numpy(), res["final"][k].numpy(), atol=2e-5, rtol=1e-3), k


def test_saturated_stop_probability_is_finite():
    """BCELoss clamps log at -100 and its backward divides by max(p(1-p), 1e-12): weights that saturate the
    sigmoid must give the same finite loss as torch and a finite update."""
    import uav_wrf_les_ppo_lstm_b200 as m
    torch.manual_seed(1)
    ref = lo.PeakAndStopPredictor()
    with torch.no_grad():
        ref.fc_stop[0].weight.fill_(400.0)
        ref.fc_stop[0].bias.fill_(300.0)
        ref.lstm.bias_ih_l0.fill_(3.0)
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    feats = np.random.default_rng(0).random((16, 20)).astype(np.float32)
    labels = np.stack([np.full(16, 0.5), np.zeros(16)], axis=1).astype(np.float32)     # p -> 1 with target 0
    want, _ = lo.loss_and_grad(ref, torch.from_numpy(feats), torch.from_numpy(labels))
    model = _model(init)
    tr = m.LstmTrainer(model, torch.from_numpy(feats).cuda(), torch.from_numpy(labels).cuda(), batch_size=16)
    tr.train_epoch(order=np.arange(16))
    got = float(tr.batch_losses[0].item())
    assert np.isfinite(got) and abs(got - want) <= 1e-5 * abs(want)
    assert bool(torch.isfinite(tr.flat).all())


def test_rollout_to_trained_stop_head_on_the_device():
    """PPO rollout -> TrajectoryLogger (training_data.nc variables on the device) -> dataset kernel -> training
    kernel: the loop the reference closes through files (train_ppo2.0.py -> training_data.nc -> train_lstm.py)."""
    import uav_wrf_les_ppo_lstm_b200 as m
    torch.manual_seed(0)
    N, T = 256, 64
    env = m.VecMethaneEnv(N, version="2.1", seed=9, field_mode="procedural", auto_reset=True)
    env.curriculum[0] = 120.0
    env.reset()
    policy = m.PPOActorCritic(device="cuda")
    with torch.no_grad():
        policy.actor.weight.mul_(30.0)
    head = m.PeakAndStopPredictor(device="cuda")
    eng = m.RolloutEngine(env, policy, head, horizon=T, with_info=True, with_trajectory=True)
    log = m.TrajectoryLogger(env, max_episodes=2000)
    for _ in range(6):
        log.consume(eng.collect())
    nc = log.nc_variables()
    elig = m.eligible_episodes(log.steps[:log.count], 20)
    assert len(elig) >= 32
    sel = np.random.default_rng(0).permutation(elig)[:1000]
    f, l = m.build_dataset(nc, sel)
    fo, lo_ = lo.build_dataset(nc, sel.tolist())
    assert np.array_equal(f.cpu().numpy(), fo) and np.array_equal(l.cpu().numpy(), lo_)
    tr = m.LstmTrainer(head, f, l)
    hist = tr.train(epochs=30)
    assert hist[-1][0] < 0.8 * hist[0][0], hist
    assert bool(torch.isfinite(tr.flat).all())

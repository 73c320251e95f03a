"""The N > 1 path on CPU: two gloo ranks exercise the product's host-side exchange logic
(dist.py) -- env sharding, the three-double advantage statistics and the summed gradient -- and
check 2 ranks x M/2 samples == 1 rank x M samples with the oracle's arithmetic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import plume_oracle as po
from oracle import ppo_oracle as pp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_data(T, N, seed):
    rng = np.random.default_rng(seed)
    return (torch.from_numpy(rng.normal(size=(T, N)).astype(np.float32)),
            torch.from_numpy(rng.normal(size=(T, N)).astype(np.float32)),
            torch.from_numpy((rng.random((T, N)) < 0.05).astype(np.float32)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import uav_wrf_les_ppo_lstm_b200.dist as pd
    r, w, pg = pd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    cfg = po.config_for("2.1")
    T, N = 32, 8                                   # global: 8 envs, 4 per rank
    rew, val, don = _make_data(T, N, 0)
    base, ids = pd.env_shard(rank, N // world)
    assert list(ids) == list(range(rank * 4, rank * 4 + 4))
    sl = slice(base, base + N // world)
    adv_local = pp.gae_quirk(rew[:, sl], val[:, sl], don[:, sl], cfg.gamma, cfg.lam)
    stats = torch.tensor([adv_local.double().sum(), (adv_local.double() ** 2).sum(), adv_local.numel()],
                         dtype=torch.float64)
    pd.allreduce_stats(stats, pg)
    mean, denom = pd.normalisation_from_stats(stats)
    adv_n = (adv_local - np.float32(mean)) / np.float32(denom)
    # gradient: each rank's loss is divided by the GLOBAL minibatch size, gradients are summed
    torch.manual_seed(0)
    model = pp.OracleActorCritic()
    rng = np.random.default_rng(1)
    M = 64
    S = torch.from_numpy(rng.random((M, 6)).astype(np.float32))
    A = torch.from_numpy(rng.integers(0, 5, M))
    mine = slice(rank * (M // world), (rank + 1) * (M // world))
    with torch.no_grad():
        P, V = model(S)
    LP = pp.categorical_log_prob(P, A) + 0.1
    ADV = torch.from_numpy(rng.normal(size=M).astype(np.float32))
    RET = V.squeeze(-1) + 0.3
    total, *_ = pp.ppo_loss(model, S[mine], A[mine], LP[mine], ADV[mine], RET[mine], V.squeeze(-1)[mine], cfg)
    local_mb = M // world
    (total * local_mb / pd.global_minibatch(local_mb, pg)).backward()
    flat = torch.cat([p.grad.flatten() for p in model.parameters()])
    pd.allreduce_gradient(flat, pg)
    # curriculum flags: canonical order = step-major, then GLOBAL env id
    dn = torch.full((3, 4), float(rank)) + torch.arange(4)[None, :] * 0.1 + torch.arange(3)[:, None] * 10
    rc = (torch.arange(12).reshape(3, 4) + 100 * rank).to(torch.uint8)
    d_all, r_all = pd.gather_episode_flags(dn, rc, pg)
    out[rank] = (adv_n.numpy(), flat.numpy(), stats.numpy(), d_all.numpy(), r_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    cfg = po.config_for("2.1")
    rew, val, don = _make_data(32, 8, 0)
    adv = pp.gae_quirk(rew, val, don, cfg.gamma, cfg.lam)
    adv_ref, _ = pp.normalise_advantages(adv, val)
    got = np.concatenate([out[0][0], out[1][0]], axis=1)
    assert np.allclose(got, adv_ref.numpy(), rtol=1e-5, atol=1e-6)
    assert np.array_equal(out[0][2], out[1][2]) and out[0][2][2] == 32 * 8
    # gradient of the global mean loss
    torch.manual_seed(0)
    model = pp.OracleActorCritic()
    rng = np.random.default_rng(1)
    M = 64
    S = torch.from_numpy(rng.random((M, 6)).astype(np.float32))
    A = torch.from_numpy(rng.integers(0, 5, M))
    with torch.no_grad():
        P, V = model(S)
    LP = pp.categorical_log_prob(P, A) + 0.1
    ADV = torch.from_numpy(rng.normal(size=M).astype(np.float32))
    RET = V.squeeze(-1) + 0.3
    total, *_ = pp.ppo_loss(model, S, A, LP, ADV, RET, V.squeeze(-1), cfg)
    total.backward()
    want = torch.cat([p.grad.flatten() for p in model.parameters()]).numpy()
    assert np.allclose(out[0][1], want, rtol=1e-4, atol=1e-7)
    assert np.array_equal(out[0][1], out[1][1])
    # gathered curriculum flags: identical on both ranks, [T, world*N] with rank-major env ids
    d_all, r_all = out[0][3], out[0][4]
    assert np.array_equal(d_all, out[1][3]) and np.array_equal(r_all, out[1][4])
    assert d_all.shape == (3, 8)
    for t in range(3):
        for rk in range(2):
            for e in range(4):
                assert np.isclose(d_all[t, rk * 4 + e], rk + 0.1 * e + 10 * t)
                assert r_all[t, rk * 4 + e] == (t * 4 + e + 100 * rk) % 256


def test_normalisation_from_stats_matches_oracle():
    import uav_wrf_les_ppo_lstm_b200.dist as pd
    a = torch.randn(1000)
    stats = torch.tensor([a.double().sum(), (a.double() ** 2).sum(), 1000.0], dtype=torch.float64)
    mean, denom = pd.normalisation_from_stats(stats)
    want, _ = pp.normalise_advantages(a, torch.zeros(1000))
    assert torch.allclose((a - mean) / denom, want, rtol=1e-5, atol=1e-6)
    z = torch.zeros(10)
    stats = torch.tensor([0.0, 0.0, 10.0], dtype=torch.float64)
    assert pd.normalisation_from_stats(stats) == (0.0, 1.0 + 1e-6)

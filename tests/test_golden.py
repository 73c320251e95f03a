"""The oracle against the committed golden vectors, which were produced by the real reference
(oracle/make_golden.py).  Runs anywhere (no reference tree, no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import plume_oracle as po
from oracle import ppo_oracle as pp
from tests.helpers import INFO_KEYS, golden_oracle_env, golden_update_inputs, load_golden

ENV_FIXTURES = ["env_v21_s11.npz", "env_v21_s12.npz", "env_v20_s21.npz", "env_v11_s31.npz"]


@pytest.mark.parametrize("name", ENV_FIXTURES)
def test_env_trace(name):
    g = load_golden(name)
    cfg, env, z_steps = golden_oracle_env(g)
    assert np.array_equal(env.src[0], g["source_pos"])
    assert np.array_equal(env.observe()[0], g["obs0"])
    assert np.array_equal(env.fields.conc[0][::50, ::50], g["conc_probe"])
    assert np.array_equal(env.fields.tke[0][::50, ::50], g["tke_probe"])
    for t, a in enumerate(g["actions"]):
        o, r, d, info = env.step(np.array([a]), z_steps[t][None])
        assert np.array_equal(o[0], g["obs"][t]), t
        assert r[0] == g["reward"][t]
        assert d[0] == g["done"][t] and info["reached"][0] == g["reached"][t]
        for j, k in enumerate(INFO_KEYS):
            assert float(info[k][0]) == g["info"][t, j]
        assert np.array_equal(env.pos32[0], g["pos"][t])
    assert np.array_equal(env.visited[0], g["visited"])


def test_update_golden():
    g = load_golden("update_s5.npz")
    cfg = po.config_for("2.1")
    model = pp.OracleActorCritic()
    model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init.")})
    with torch.no_grad():
        probs, _ = model(torch.from_numpy(g["states"]))
    assert np.array_equal(probs.numpy(), g["probs0"])
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)
    pp.ppo_update(model, opt, torch.from_numpy(g["states"]), torch.from_numpy(g["actions"]),
                  torch.from_numpy(g["rewards"]), torch.from_numpy(g["values"]), torch.from_numpy(g["log_probs"]),
                  torch.from_numpy(g["dones"]), cfg, perms=list(g["perms"]))
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), g["final." + k]), k


def test_update_large_minibatch_golden():
    """The reference's _update_model with BATCH_SIZE = 2048 on 8192 transitions (the shape class the tcgen05
    gradient kernel handles): the oracle reproduces the reference's parameters bit for bit."""
    import dataclasses
    g = load_golden("update_large_s6.npz")
    cfg = dataclasses.replace(po.config_for("2.1"), batch_size=int(g["batch_size"]))
    states, actions, rewards, dones = golden_update_inputs(g)
    model = pp.OracleActorCritic()
    model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init.")})
    with torch.no_grad():
        _, values = model(torch.from_numpy(states))
    assert np.array_equal(values.squeeze(-1).numpy(), g["values"])           # the regenerated inputs are the fixture's
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)
    pp.ppo_update(model, opt, torch.from_numpy(states), torch.from_numpy(actions), torch.from_numpy(rewards),
                  torch.from_numpy(g["values"]), torch.from_numpy(g["log_probs"]), torch.from_numpy(dones), cfg,
                  perms=list(g["perms"].astype(np.int64)))
    # bit-equal with the thread count the fixture was written with (torch's CPU reductions over 2048 samples split
    # by thread); otherwise equal up to what Adam makes of fp32 summation order
    exact = all(np.array_equal(v.numpy(), g["final." + k]) for k, v in model.state_dict().items())
    if not exact:
        for k, v in model.state_dict().items():
            assert np.allclose(v.numpy(), g["final." + k], rtol=0, atol=2 * cfg.learning_rate), k
        num = den = 0.0
        for k, v in model.state_dict().items():
            d_ref = (torch.from_numpy(g["final." + k]) - torch.from_numpy(g["init." + k])).flatten().double()
            d_got = (v - torch.from_numpy(g["init." + k])).flatten().double()
            num += float((d_ref * d_got).sum())
            den += float(d_ref.norm() * d_got.norm())
        assert num / den > 0.999


def test_lstm_golden():
    g = load_golden("lstm_s7.npz")
    m = pp.OraclePeakAndStop()
    m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
    with torch.no_grad():
        peak, stop = m(torch.from_numpy(g["windows"]).unsqueeze(-1))
    assert np.array_equal(peak.numpy(), g["peak"]) and np.array_equal(stop.numpy(), g["stop_prob"])
    flags = g["stop_prob"] > 0.8
    assert 0 < flags.sum() < len(flags)       # the fixture straddles the threshold


def test_curriculum_golden():
    g = load_golden("curriculum_s3.npz")

    class E:
        current_radius = 50.0
        explore_bonus = 0.6
    env = E()
    cur = pp.OracleCurriculum(env, po.config_for("2.1"))
    for i, s in enumerate(g["success"]):
        cur.update(bool(s))
        assert cur.current_radius == g["radius"][i] and cur.explore_bonus == g["explore_bonus"][i]
        assert env.current_radius == g["env_radius"][i]


def test_trend_golden():
    g = load_golden("trend_s9.npz")
    for i in range(len(g["label"])):
        assert pp.trend_label(g["conc"][i], g["pos"][i], g["src"][i])[0] == g["label"][i]


def test_gae_quirks():
    # hand-checked 3-step case: self-bootstrap at the end, dones[t+1] masking
    r = torch.tensor([1.0, 2.0, 3.0])
    v = torch.tensor([0.5, 0.25, 0.125])
    d = torch.tensor([0.0, 1.0, 0.0])
    a = pp.gae_quirk(r, v, d, 0.99, 0.95)
    a2 = 3.0 + 0.99 * 0.125 - 0.125
    a1 = (2.0 + 0.99 * 0.125 - 0.25) + 0.99 * 0.95 * a2          # masks with dones[2] = 0
    a0 = (1.0 + 0.0 - 0.5) + 0.0                                 # masks with dones[1] = 1
    assert torch.allclose(a, torch.tensor([a0, a1, a2]), rtol=1e-6)

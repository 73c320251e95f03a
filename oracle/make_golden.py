"""Generates tests/golden/*.npz by running the UNMODIFIED reference (container only).

    python -m oracle.make_golden

The reference is unseeded; here the *global* numpy generator is seeded (``np.random.seed``)
so that the reference's own ``np.random.rand/randn`` calls (environment.py:44,58,60,108)
become reproducible: a consumer rebuilds the identical stream with
``np.random.RandomState(seed)`` (legacy generator, frozen by numpy's compatibility policy).
torch's global generator is seeded the same way for ``torch.randperm`` (train_ppo2.0.py:43).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

from .ref_harness import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
INFO_KEYS = ("concentration_reward", "explore_reward", "move_penalty", "tke_penalty", "boundary_penalty")


def env_trace(version: str, seed: int, steps: int, radius: float, explore_p: float):
    ref = load_reference(version)
    ref.environment.np = np            # the real numpy: draws come from the seeded global generator
    np.random.seed(seed)
    env = ref.environment.MethaneEnv()
    env.current_radius = radius
    pol = np.random.RandomState(seed + 1000)
    rec = {k: [] for k in ("actions", "obs", "reward", "done", "reached", "info", "pos")}
    obs0 = env._get_obs()
    for _ in range(steps):
        # walk towards the source with probability 1-explore_p (so that `reached` fires)
        dx, dy = env.source_pos - env.agent_pos
        if pol.rand() < explore_p:
            a = int(pol.randint(0, 5))
        elif abs(dx) > abs(dy):
            a = 3 if dx > 0 else 4
        else:
            a = 1 if dy > 0 else 2
        o, r, d, info = env.step(a)
        rec["actions"].append(a)
        rec["obs"].append(o)
        rec["reward"].append(float(r))
        rec["done"].append(bool(d))
        rec["reached"].append(bool(env.trajectory[-1]["reached"]))
        rec["info"].append([float(info[k]) for k in INFO_KEYS])
        rec["pos"].append(env.agent_pos.copy())
        if d:
            break
    visited = np.zeros((10, 10), dtype=np.int64)
    for (gx, gy), v in env.visited.items():
        visited[int(gx), int(gy)] = v
    return dict(version=version, seed=seed, radius=radius, source_pos=env.source_pos, obs0=obs0,
                actions=np.array(rec["actions"], dtype=np.int32), obs=np.array(rec["obs"], dtype=np.float32),
                reward=np.array(rec["reward"], dtype=np.float64), done=np.array(rec["done"]),
                reached=np.array(rec["reached"]), info=np.array(rec["info"], dtype=np.float64),
                pos=np.array(rec["pos"], dtype=np.float32), visited=visited,
                conc_probe=env.conc_field[::50, ::50].copy(), tke_probe=env.tke_field[::50, ::50].copy())


def update_trace(seed: int, m: int = 256):
    ref = load_reference("2.1")
    torch.manual_seed(seed)
    model = ref.model.PPOActorCritic(6, 5)
    init = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    rng = np.random.RandomState(seed)
    states = rng.rand(m, 6).astype(np.float32)
    actions = rng.randint(0, 5, m)
    rewards = rng.randn(m).astype(np.float32)
    dones = (rng.rand(m) < 0.04)
    with torch.no_grad():
        probs, values = model(torch.FloatTensor(states))
    logp = torch.distributions.Categorical(probs).log_prob(torch.LongTensor(actions)).numpy()
    logp = (logp + 0.05 * rng.randn(m)).astype(np.float32)
    buf = ref.model.PPOBuffer()
    for i in range(m):
        buf.store(states[i], actions[i], rewards[i], values[i].item(), logp[i], dones[i])
    opt = torch.optim.Adam(model.parameters(), lr=ref.config.LEARNING_RATE)
    torch.manual_seed(seed + 1)
    perms = [torch.randperm(m).numpy() for _ in range(ref.config.EPOCHS)]
    torch.manual_seed(seed + 1)
    ref.train._update_model(buf, model, opt)
    final = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    out = dict(states=states, actions=actions.astype(np.int64), rewards=rewards,
               values=values.squeeze(-1).numpy(), log_probs=logp, dones=dones.astype(np.float32),
               perms=np.stack(perms), probs0=probs.numpy())
    out.update({f"init.{k}": v for k, v in init.items()})
    out.update({f"final.{k}": v for k, v in final.items()})
    return out


def lstm_trace(seed: int):
    ref = load_reference("2.1")
    torch.manual_seed(seed)
    lstm = ref.evaluate_with_lstm.PeakAndStopPredictor(input_dim=1)
    with torch.no_grad():
        lstm.fc_stop[0].bias.fill_(2.6)        # spread stop_prob around the 0.8 threshold
        lstm.fc_stop[0].weight.mul_(40.0)
        lstm.lstm.weight_ih_l0.mul_(6.0)
    rng = np.random.RandomState(seed)
    windows = (rng.rand(96, 20) * rng.rand(96, 1)).astype(np.float32)
    with torch.no_grad():
        peak, stop = lstm(torch.FloatTensor(windows).unsqueeze(-1))
    out = dict(windows=windows, peak=peak.numpy(), stop_prob=stop.numpy())
    out.update({f"sd.{k}": v.numpy() for k, v in lstm.state_dict().items()})
    return out


def curriculum_trace(seed: int, episodes: int = 1500):
    ref = load_reference("2.1")

    class E:
        current_radius = 50.0
        explore_bonus = 0.6

    env = E()
    tr = ref.model.PPOTrainer(env, None, None)
    rng = np.random.RandomState(seed)
    succ = rng.rand(episodes) < np.clip(0.15 + np.arange(episodes) / 900.0, 0, 0.9)
    radius, bonus, env_radius = [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for s in succ:
            tr.update(bool(s))
            radius.append(tr.current_radius)
            bonus.append(tr.explore_bonus)
            env_radius.append(env.current_radius)
    return dict(success=succ, radius=np.array(radius), explore_bonus=np.array(bonus), env_radius=np.array(env_radius))


def trend_trace(seed: int):
    ref = load_reference("2.1")
    rng = np.random.RandomState(seed)
    conc = rng.rand(32, 20) * 100
    pos = rng.rand(32, 2) * 499
    src = rng.rand(32, 2) * 400 + 50
    label = [ref.model.calculate_dynamic_label({"concentrations": conc[i], "positions": pos[i][None], "source_pos": src[i]})
             for i in range(32)]
    return dict(conc=conc, pos=pos, src=src, label=np.array(label))


def main():
    os.makedirs(OUT, exist_ok=True)
    for version, seed, steps, radius, p in (("2.1", 11, 400, 8.0, 0.6), ("2.1", 12, 400, 50.0, 1.0),
                                            ("2.0", 21, 300, 6.0, 0.5), ("1.1", 31, 300, 10.0, 0.6)):
        t = env_trace(version, seed, steps, radius, p)
        np.savez_compressed(os.path.join(OUT, f"env_v{version.replace('.', '')}_s{seed}.npz"), **t)
        print(version, seed, "steps", len(t["actions"]), "reached", bool(t["reached"][-1]))
    np.savez_compressed(os.path.join(OUT, "update_s5.npz"), **update_trace(5))
    np.savez_compressed(os.path.join(OUT, "lstm_s7.npz"), **lstm_trace(7))
    np.savez_compressed(os.path.join(OUT, "curriculum_s3.npz"), **curriculum_trace(3))
    np.savez_compressed(os.path.join(OUT, "trend_s9.npz"), **trend_trace(9))
    make_lstm_train(13)



def make_lstm_train(seed: int = 13, n_episodes: int = 90, epochs: int = 8):
    """N3 fixture: a synthetic training_data.nc record, the selected episodes, the initial weights, the epoch
    orders and what the torch-CPU training loop (oracle/lstm_train_oracle.py, pinned bit-equal against the
    reference's train() by tests/test_lstm_train_oracle.py) produces: every minibatch loss and gradient norm, the
    first minibatch's gradient, the parameters after `epochs` epochs."""
    import torch

    from oracle import lstm_train_oracle as lo
    nc = lo.synthetic_nc(n_episodes, seed=seed)
    elig = lo.eligible_episodes(nc, 20)
    sel = [int(elig[i]) for i in lo.select_episodes(len(elig), seed)]
    f, l = lo.build_dataset(nc, sel)
    torch.manual_seed(seed)
    model = lo.PeakAndStopPredictor()
    init = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    orders = np.stack([torch.randperm(len(l), generator=g).numpy() for _ in range(epochs)])
    _, grad0 = lo.loss_and_grad(model, torch.from_numpy(f[orders[0][:64]]), torch.from_numpy(l[orders[0][:64]]))
    res = lo.train(model, f, l, epochs=epochs, orders=orders)
    out = {"selected": np.array(sel, np.int32), "features": f, "labels": l, "orders": orders,
           "batch_losses": res["batch_losses"], "grad_norms": res["grad_norms"], "grad0": grad0.numpy(),
           "epoch_means": np.array([h[0] for h in res["history"]])}
    out.update({"nc_" + k: v for k, v in nc.items()})
    out.update({"init_" + k: v for k, v in init.items()})
    out.update({"final_" + k: v.numpy() for k, v in res["final"].items()})
    np.savez_compressed(os.path.join(OUT, f"lstm_train_s{seed}.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "lstm_train":
        make_lstm_train(13)
    else:
        main()

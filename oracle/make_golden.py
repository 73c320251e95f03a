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


def update_trace(seed: int, m: int = 256, batch_size: int | None = None, compact: bool = False):
    """One full ``_update_model`` of the reference on ``m`` synthetic transitions.  ``batch_size`` patches the
    train module's ``BATCH_SIZE`` (train_ppo2.0.py:47 reads the module global) so that a minibatch is large enough
    for the tcgen05 gradient kernel; ``compact`` drops what a consumer can rebuild from ``RandomState(seed)``
    (states, actions, rewards, dones -- drawn in exactly this order)."""
    ref = load_reference("2.1")
    saved_bs = ref.train.BATCH_SIZE
    if batch_size is not None:
        ref.train.BATCH_SIZE = int(batch_size)
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
    try:
        ref.train._update_model(buf, model, opt)
    finally:
        ref.train.BATCH_SIZE = saved_bs
    final = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    if compact:
        out = dict(seed=seed, m=m, batch_size=int(batch_size or saved_bs), values=values.squeeze(-1).numpy(),
                   log_probs=logp, perms=np.stack(perms).astype(np.int16 if m <= 32767 else np.int32))
    else:
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


def driver_loop_trace(seed: int, episodes: int = 3, actor_bias=(-1.5, 1.0, -1.5, 1.0, -1.5)):
    """The loop body of ``train_ppo()`` (PPOV2.1/train_ppo2.0.py:137-251) on the reference's own objects --
    ``MethaneEnv``, ``PPOActorCritic``, ``PPOBuffer``, ``_update_model``, ``PPOTrainer``, ``RadiusTracker``,
    ``NetCDFWriter`` (over an in-memory Dataset) -- for a few episodes.  Recorded: every step's state / action /
    reward / value / log-prob / done, the ``torch.randperm`` draws of every ``_update_model``, the parameters before
    and after, the curriculum after every episode, the CSV rows and the NetCDF variables.  The env draws come from
    the seeded global numpy generator in the reference's own order: constructor reset, then per episode
    ``rand(2), randn(G,G), rand(G,G)`` and ``randn(2)`` per step."""
    from .ref_harness import MemDataset
    ref = load_reference("2.1")
    ref.environment.np = np
    ref.model.Dataset = MemDataset                     # `from netCDF4 import Dataset` (model.py:6), stubbed
    cfgm = ref.config
    GRID_SIZE, MAX_STEPS, BATCH_SIZE = cfgm.GRID_SIZE, cfgm.MAX_STEPS, cfgm.BATCH_SIZE
    np.random.seed(seed)
    torch.manual_seed(seed)
    env = ref.environment.MethaneEnv()
    model = ref.model.PPOActorCritic(6, 5)
    with torch.no_grad():                               # a policy that heads for the source quadrant, so that
        model.actor.bias.copy_(torch.tensor(actor_bias))  # some episodes reach it inside the initial radius
    init = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    optimizer = torch.optim.Adam(model.parameters(), lr=cfgm.LEARNING_RATE)
    buffer = ref.model.PPOBuffer()
    trainer = ref.model.PPOTrainer(env, model, optimizer)
    radius_tracker = ref.train.RadiusTracker()
    nc_writer = ref.model.NetCDFWriter(filename="mem", grid_size=GRID_SIZE, max_episodes=episodes, max_steps=MAX_STEPS)
    perms, real_randperm = [], torch.randperm

    def recording_randperm(n, *a, **k):
        p = real_randperm(n, *a, **k)
        perms.append(p.numpy().copy())
        return p

    rec = {k: [] for k in ("state", "action", "reward", "value", "logp", "done", "pos", "conc", "update_after")}
    training_data, curriculum, sources = [], [], []
    n_updates = 0
    torch.randperm = recording_randperm
    try:
        for episode in range(episodes):
            state = env.reset()
            sources.append(env.source_pos.copy())
            done = False
            episode_data = {'total_reward': 0, 'steps': 0, 'x': [], 'y': [], 'conc': [], 'success': False,
                            'source_conc': 0.0, 'source_x': 0.0, 'source_y': 0.0,
                            'current_radius': trainer.current_radius, 'Conc_Reward': 0.0, 'Explore_Reward': 0.0,
                            'Move_Penalty': 0.0, 'TKE_Penalty': 0.0, 'Boundary_Penalty': 0.0}
            while not done:
                state_t = torch.FloatTensor(state).unsqueeze(0)
                with torch.no_grad():
                    probs, value = model(state_t)
                action_dist = torch.distributions.Categorical(probs)
                action = action_dist.sample().item()
                next_state, reward, done, info = env.step(action)
                x, y = env.agent_pos
                current_conc = env.conc_field[np.clip(int(x), 0, GRID_SIZE - 1), np.clip(int(y), 0, GRID_SIZE - 1)]
                episode_data['x'].append(float(x))
                episode_data['y'].append(float(y))
                episode_data['conc'].append(float(current_conc))
                episode_data['total_reward'] += reward
                episode_data['steps'] += 1
                episode_data['Conc_Reward'] += info['concentration_reward']
                episode_data['Explore_Reward'] += info['explore_reward']
                episode_data['Move_Penalty'] += info['move_penalty']
                episode_data['TKE_Penalty'] += info['tke_penalty']
                episode_data['Boundary_Penalty'] += info['boundary_penalty']
                logp = action_dist.log_prob(torch.tensor(action)).item()
                buffer.store(state, action, reward, value.item(), logp, done)
                rec["state"].append(np.array(state, dtype=np.float32))
                rec["action"].append(action)
                rec["reward"].append(float(reward))
                rec["value"].append(value.item())
                rec["logp"].append(logp)
                rec["done"].append(bool(done))
                rec["pos"].append(env.agent_pos.copy())
                rec["conc"].append(float(current_conc))
                updated = False
                if len(buffer.states) >= BATCH_SIZE:
                    ref.train._update_model(buffer, model, optimizer)
                    buffer.clear()
                    n_updates += 1
                    updated = True
                rec["update_after"].append(updated)
                state = next_state
            if env.trajectory[-1]['reached']:
                final_pos = np.clip(env.agent_pos.astype(int), 0, GRID_SIZE - 1)
                episode_data['source_conc'] = env.conc_field[final_pos[0], final_pos[1]]
                episode_data['source_x'] = float(env.agent_pos[0])
                episode_data['source_y'] = float(env.agent_pos[1])
                episode_data['success'] = True
            radius_tracker.update(current_radius=trainer.current_radius, episode_data=episode_data,
                                  is_success=episode_data['success'])
            if trainer.current_radius in radius_tracker.radius_history and episode_data['success']:
                nc_writer.write_episode_data(episode_idx=episode, steps=episode_data['steps'],
                                             x=np.array(episode_data['x']), y=np.array(episode_data['y']),
                                             conc=np.array(episode_data['conc']), source_x=episode_data['source_x'],
                                             source_y=episode_data['source_y'], source_conc=episode_data['source_conc'],
                                             sigma=env.gaussian_params['sigma'], peak=env.gaussian_params['peak'])
            nc_writer.write_episode_data(episode_idx=episode, steps=len(episode_data['x']),
                                         x=np.array(episode_data['x']), y=np.array(episode_data['y']),
                                         conc=np.array(episode_data['conc']), source_x=env.gaussian_params['mu_x'],
                                         source_y=env.gaussian_params['mu_y'], source_conc=env.gaussian_params['peak'],
                                         sigma=env.gaussian_params['sigma'], peak=env.gaussian_params['peak'])
            training_data.append([episode + 1, episode_data['total_reward'], int(episode_data['success']),
                                  episode_data['Conc_Reward'], episode_data['Explore_Reward'],
                                  episode_data['Move_Penalty'], episode_data['TKE_Penalty'],
                                  episode_data['Boundary_Penalty'], episode_data['steps'], episode_data['source_conc'],
                                  trainer.current_radius])
            with contextlib.redirect_stdout(io.StringIO()):
                trainer.update(episode_data['success'])
            curriculum.append([trainer.current_radius, trainer.explore_bonus, env.current_radius, env.explore_bonus])
    finally:
        torch.randperm = real_randperm
    final = {k: v.clone().numpy() for k, v in model.state_dict().items()}
    out = dict(seed=seed, episodes=episodes, n_updates=n_updates, actor_bias=np.array(actor_bias, dtype=np.float32),
               state=np.stack(rec["state"]), action=np.array(rec["action"], dtype=np.int8),
               reward=np.array(rec["reward"], dtype=np.float64), value=np.array(rec["value"], dtype=np.float32),
               logp=np.array(rec["logp"], dtype=np.float32), done=np.array(rec["done"]),
               pos=np.stack(rec["pos"]).astype(np.float32), conc=np.array(rec["conc"], dtype=np.float64),
               update_after=np.array(rec["update_after"]), perms=np.stack(perms).astype(np.int16),
               csv=np.array(training_data, dtype=np.float64), curriculum=np.array(curriculum, dtype=np.float64),
               sources=np.stack(sources))
    for name, var in nc_writer.ncfile.variables.items():
        out["nc_" + name] = var.data
    out.update({f"init.{k}": v for k, v in init.items()})
    out.update({f"final.{k}": v for k, v in final.items()})
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for version, seed, steps, radius, p in (("2.1", 11, 400, 8.0, 0.6), ("2.1", 12, 400, 50.0, 1.0),
                                            ("2.0", 21, 300, 6.0, 0.5), ("1.1", 31, 300, 10.0, 0.6)):
        t = env_trace(version, seed, steps, radius, p)
        np.savez_compressed(os.path.join(OUT, f"env_v{version.replace('.', '')}_s{seed}.npz"), **t)
        print(version, seed, "steps", len(t["actions"]), "reached", bool(t["reached"][-1]))
    np.savez_compressed(os.path.join(OUT, "update_s5.npz"), **update_trace(5))
    make_update_large()
    np.savez_compressed(os.path.join(OUT, "lstm_s7.npz"), **lstm_trace(7))
    np.savez_compressed(os.path.join(OUT, "curriculum_s3.npz"), **curriculum_trace(3))
    np.savez_compressed(os.path.join(OUT, "trend_s9.npz"), **trend_trace(9))
    make_lstm_train(13)
    make_driver_loop()


def make_update_large(seed: int = 6, m: int = 8192, batch_size: int = 2048):
    """The reference's _update_model with 2048-sample minibatches: pins the tensor-core gradient kernel."""
    np.savez_compressed(os.path.join(OUT, f"update_large_s{seed}.npz"), **update_trace(seed, m, batch_size, compact=True))


def make_driver_loop(seed: int = 45, episodes: int = 3):
    t = driver_loop_trace(seed, episodes)
    np.savez_compressed(os.path.join(OUT, f"driver_loop_s{seed}.npz"), **t)
    print("driver loop", seed, "steps", len(t["action"]), "updates", int(t["n_updates"]), "success", t["csv"][:, 2])


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
    elif len(sys.argv) > 1 and sys.argv[1] == "update_large":
        make_update_large()
    elif len(sys.argv) > 1 and sys.argv[1] == "driver_loop":
        make_driver_loop(*(int(a) for a in sys.argv[2:]))
    else:
        main()

"""Shared test helpers (test infrastructure)."""
import os

import numpy as np

from oracle import plume_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INFO_KEYS = ("concentration_reward", "explore_reward", "move_penalty", "tke_penalty", "boundary_penalty")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_env_stream(g):
    """Rebuilds the draws the reference consumed for an env trace: (u_src, z_field, u_field,
    z_steps[T,2]) from the legacy RandomState the fixture was generated with."""
    cfg = po.config_for(str(g["version"]))
    G = cfg.grid_size
    rs = np.random.RandomState(int(g["seed"]))
    u_src = rs.rand(2)
    z_field = rs.randn(G, G)
    u_field = rs.rand(G, G)
    T = len(g["actions"])
    z_steps = np.stack([rs.randn(2) for _ in range(T)])
    return cfg, u_src, z_field, u_field, z_steps


def golden_oracle_env(g):
    cfg, u_src, z_field, u_field, z_steps = golden_env_stream(g)
    env = po.OracleVecEnv(cfg, 1)
    env.reset_env(0, u_src, z_field, u_field)
    env.current_radius[:] = float(g["radius"])
    return cfg, env, z_steps


def golden_update_inputs(g):
    """Rebuilds what oracle/make_golden.py::update_trace drew from ``RandomState(seed)`` for a compact update
    fixture (states, actions, rewards, dones -- in exactly that order); values / log-probs / permutations / the
    parameters before and after come from the fixture itself."""
    m = int(g["m"])
    rng = np.random.RandomState(int(g["seed"]))
    states = rng.rand(m, 6).astype(np.float32)
    actions = rng.randint(0, 5, m).astype(np.int64)
    rewards = rng.randn(m).astype(np.float32)
    dones = (rng.rand(m) < 0.04).astype(np.float32)
    return states, actions, rewards, dones

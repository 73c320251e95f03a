"""Times the UNMODIFIED reference's own training loop body on the CPU -- TEST / BENCH INFRASTRUCTURE ONLY.

``bench.py --impl reference`` and ``cpu_baseline`` call this when a copy of the reference's PPOV2.1 folder is
available (``oracle/_ref/PPOV2.1``: written by ``__graft_entry__.build()`` from ``/root/reference`` in the build
container, git-ignored, shipped to the GPU box with the snapshot like the built ``.so`` files).  The loop is the body of
``train_ppo()`` (PPOV2.1/train_ppo2.0.py:137-192,251) on the reference's own ``MethaneEnv``, ``PPOActorCritic``,
``PPOBuffer``, ``_update_model`` and ``PPOTrainer`` -- without the NetCDF / CSV writers (file I/O, netCDF4 absent) and
cut off after ``n_steps`` env steps instead of 2000 episodes (``train_ppo()`` itself takes no arguments)."""
from __future__ import annotations

import contextlib
import io
import os
import time

HERE = os.path.dirname(os.path.abspath(__file__))
SHIPPED = os.path.join(HERE, "_ref")


def shipped_reference_available() -> bool:
    return os.path.isdir(os.path.join(SHIPPED, "PPOV2.1"))


def reference_loop(seed: int, n_steps: int):
    """Returns (env steps done, seconds).  Single process, ``torch.set_num_threads(1)`` is the caller's business."""
    if not os.path.isdir("/root/reference/PPOV2.1") and shipped_reference_available():
        os.environ.setdefault("PLUME_REFERENCE_ROOT", SHIPPED)
    import numpy as np
    import torch

    from . import ref_harness as rh
    if not rh.reference_available() and shipped_reference_available():
        rh.REFERENCE_ROOT = SHIPPED
    ref = rh.load_reference("2.1")
    ref.environment.np = np
    np.random.seed(seed)
    torch.manual_seed(seed)
    BATCH_SIZE = ref.config.BATCH_SIZE
    env = ref.environment.MethaneEnv()
    model = ref.model.PPOActorCritic(6, 5)
    optimizer = torch.optim.Adam(model.parameters(), lr=ref.config.LEARNING_RATE)
    buffer = ref.model.PPOBuffer()
    trainer = ref.model.PPOTrainer(env, model, optimizer)
    steps = 0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        while steps < n_steps:
            state = env.reset()
            done = False
            while not done and steps < n_steps:
                state_t = torch.FloatTensor(state).unsqueeze(0)
                with torch.no_grad():
                    probs, value = model(state_t)
                action_dist = torch.distributions.Categorical(probs)
                action = action_dist.sample().item()
                next_state, reward, done, info = env.step(action)
                buffer.store(state, action, reward, value.item(), action_dist.log_prob(torch.tensor(action)).item(), done)
                if len(buffer.states) >= BATCH_SIZE:
                    ref.train._update_model(buffer, model, optimizer)
                    buffer.clear()
                state = next_state
                steps += 1
            if done:
                trainer.update(bool(env.trajectory[-1]["reached"]))
    return steps, time.perf_counter() - t0

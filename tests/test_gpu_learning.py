"""Statistical regression (SURVEY section 4 iii): does the batched trainer LEARN like the reference?

The reference's own 2000-episode run (PPOV2.1/training_results2_0.csv:2-2001) ends with 63.6 % successful episodes
and the curriculum radius at 8.28 (from 50).  The same schedule here -- an update of 5 epochs x one 256-sample
minibatch every 256 transitions (config.py: BATCH_SIZE 256, EPOCHS 5, lr 3e-5) -- on one env with horizon 256 (the
reference's temporal structure) and on 8 envs x 32 steps.  Three seeds measured on B200
(profiles/r2_learning_curve_n1.jsonl): 66.0 / 65.0 / 66.0 % and radius 7.7 / 10.1 / 9.3; the bands below are wide
enough for seed noise and narrow enough to catch a trainer that does not move the curriculum."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))


@pytest.mark.parametrize("envs,horizon", [(1, 256), (8, 32)])
def test_reference_schedule_reproduces_the_reference_training_statistics(envs, horizon):
    import learning_curve as lc
    r = lc.run(envs, horizon, episodes=2000, seed=3)
    assert 2000 <= r["episodes"] <= 2300
    assert 0.55 <= r["success_rate"] <= 0.75, r["success_rate"]           # reference: 0.636
    assert 5.0 <= r["final_radius"] <= 15.0, r["final_radius"]            # reference: 8.28
    radii = [h["radius"] for h in r["history"]]
    assert radii[0] == 50.0 and all(b <= a + 1e-9 or b <= 50.0 for a, b in zip(radii, radii[1:]))
    assert min(radii) < 20.0

"""GPU parity of the batched evaluators (N1: PPOV2.1/evaluate_with_lstm.py, PPOV2.0/evaluate_with_lstm.py,
PPOV1.1/evaluate_model.py): the stop decision of every env is replayed through the oracle's restatement of
the reference controllers on the env's own recorded trace."""
import numpy as np
import pytest
import torch

from oracle import plume_oracle as po
from oracle import ppo_oracle as pp

pytestmark = pytest.mark.gpu


def pb():
    import uav_wrf_les_ppo_lstm_b200 as m
    return m


def _policy(m, seed, stay_bias=0.0):
    torch.manual_seed(seed)
    model = m.PPOActorCritic(device="cuda")
    with torch.no_grad():
        model.actor.weight.mul_(40.0)
        model.actor.bias[0] += stay_bias
    return model


def _stack(trace, key):
    return torch.stack(trace[key]).cpu().numpy()


def test_eval_v21_lstm_stop():
    m = pb()
    model = _policy(m, 1)
    head = m.PeakAndStopPredictor(device="cuda")
    with torch.no_grad():
        head.fc_stop[0].bias.fill_(1.386)      # sigmoid = 0.8: decisions straddle the threshold
        head.fc_stop[0].weight.mul_(40.0)
        head.lstm.weight_ih_l0.mul_(6.0)
    trace = {}
    res = m.evaluate_policy(model, stop="lstm", head=head, num_envs=96, seed=3, trace=trace)
    conc = _stack(trace, "conc")                      # [T, N]
    ora = pp.OraclePeakAndStop()
    ora.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    T, N = conc.shape
    stop_step = res.stop_step.cpu().numpy()
    done = _stack(trace, "done")
    checked = 0
    for i in range(N):
        want, margin = 0, np.inf
        for t in range(19, T):
            _, prob, _ = pp.lstm_window_stop(ora, conc[None, : t + 1, i], 20, 0.8)
            margin = min(margin, abs(float(prob[0]) - 0.8))
            if prob[0] > 0.8:
                want = t + 1
                break
            if done[t, i]:
                break
        if margin > 1e-4:
            assert stop_step[i] == want, (i, stop_step[i], want)
            checked += 1
    assert checked > N // 2
    assert res.stopped_early.any() and len(set(stop_step.tolist())) >= 5      # decisions at many different steps
    s = res.summary()
    assert 0 <= s["success_rate"] <= 1 and s["episodes"] == 96 and s["mean_steps"] >= 20
    # success = deviation <= 50 (evaluate_with_lstm.py:88)
    assert torch.equal(res.success, res.deviations <= 50.0)


def test_eval_v20_threshold_controller():
    m = pb()
    model = _policy(m, 2)
    head = m.ConcentrationThresholdPredictor(device="cuda")
    with torch.no_grad():
        head.fc[4].bias.fill_(11.0)                  # threshold ~ 10 ppb: inside the background range
    trace = {}
    res = m.evaluate_policy(model, stop="threshold", head=head, num_envs=64, seed=5, scaler=(2.0, 2.0), trace=trace)
    conc = _stack(trace, "conc")
    done = _stack(trace, "done")
    T, N = conc.shape
    ora = pp.OracleThresholdPredictor().eval()
    ora.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    stop_step = res.stop_step.cpu().numpy()
    checked = 0
    for i in range(N):
        ctrl = pp.OracleThresholdController(ora, 2.0, 2.0, window_size=10)
        traj, want, margin = [], 0, np.inf
        for t in range(T):
            traj.append(conc[t, i])
            step = t + 1
            if step % 10 == 0:
                ctrl.update_threshold(traj)
            stop = ctrl.should_stop(conc[t, i], step)
            if ctrl.current_threshold is not None and step >= 20:
                margin = min(margin, abs(conc[t, i] - ctrl.current_threshold),
                             abs(np.mean(ctrl.conc_buffer) - ctrl.current_threshold))
            if stop:
                want = step
                break
            if done[t, i]:
                break
        if margin > 1e-3:
            assert stop_step[i] == want, (i, stop_step[i], want)
            checked += 1
    assert checked > N // 2 and (stop_step > 0).any()
    assert torch.equal(res.success, res.deviations <= 40.0)          # PPOV2.0/config.py:43


def test_eval_v11_fixed_threshold():
    m = pb()
    cfg = po.config_for("1.1")
    model = _policy(m, 3, stay_bias=50.0)             # a policy that mostly hovers: the stability test can fire
    trace = {}
    res = m.evaluate_policy(model, stop="fixed", num_envs=64, seed=7, trace=trace)
    pos = _stack(trace, "pos")                        # [T, N, 2]
    cr = _stack(trace, "conc_reward")
    done = _stack(trace, "done")
    T, N = cr.shape
    stop_step = res.stop_step.cpu().numpy()
    checked = 0
    for i in range(N):
        want, margin = 0, np.inf
        for t in range(T):
            positions = [pos[k, i] for k in range(t + 1)]
            if t + 1 >= 10:
                margin = min(margin, abs(float(np.std(np.asarray(positions[-10:]), axis=0).mean()) - 2.0))
            if pp.fixed_threshold_stop(positions, float(cr[t, i]), cfg):
                want = t + 1
                break
            if done[t, i]:
                break
        if margin > 1e-3:
            assert stop_step[i] == want, (i, stop_step[i], want)
            checked += 1
    assert checked > N // 2 and (stop_step > 0).any()
    assert int(res.steps.max()) <= 2000

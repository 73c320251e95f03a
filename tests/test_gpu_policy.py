"""GPU parity of the actor-critic forward / sampling (P4, P4s, K3) and the LSTM stop heads
(P4L, K4) and trend features (P4t) against the torch-CPU oracle and the reference's golden
vectors.  Tolerance: fp32 rel 1e-5 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as pp
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def pb():
    import uav_wrf_les_ppo_lstm_b200 as m
    return m


@pytest.mark.parametrize("batch", [1, 31, 32, 33, 1000, 4096])
def test_policy_forward_vs_oracle(batch):
    torch.manual_seed(batch)
    ora = pp.OracleActorCritic()
    with torch.no_grad():          # non-trivial LayerNorm affine and biases
        for p in ora.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    model = pb().PPOActorCritic(device="cuda")
    model.load_state_dict(ora.state_dict())
    x = torch.rand(batch, 6) * torch.tensor([1, 1, 1, 1.5, 1, 1])
    with torch.no_grad():
        p_ref, v_ref = ora(x)
    p, v = model(x.cuda())
    assert p.shape == (batch, 5) and v.shape == (batch, 1)
    assert torch.allclose(p.cpu(), p_ref, rtol=1e-5, atol=1e-7)
    assert torch.allclose(v.cpu(), v_ref, rtol=1e-5, atol=2e-6)


def test_policy_forward_golden_reference_weights():
    g = load_golden("update_s5.npz")
    model = pb().PPOActorCritic(device="cuda")
    model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init.")})
    p, v = model(torch.from_numpy(g["states"]).cuda())
    assert np.allclose(p.cpu().numpy(), g["probs0"], rtol=1e-5, atol=1e-7)
    assert np.allclose(v.cpu().numpy()[:, 0], g["values"], rtol=1e-5, atol=2e-6)
    # state_dict round trip back into the oracle module
    ora = pp.OracleActorCritic()
    ora.load_state_dict({k: t.cpu() for k, t in model.state_dict().items()})


def test_act_sampling_and_logprob():
    torch.manual_seed(3)
    model = pb().PPOActorCritic(device="cuda")
    with torch.no_grad():
        model.actor.weight.mul_(60.0)             # spread the action probabilities
    B = 2000
    x = torch.rand(B, 6, device="cuda")
    u = torch.rand(B, device="cuda")
    a, lp, v, p = model.act(x, uniforms=u)
    p_cpu = p.cpu()
    pn = p_cpu / p_cpu.sum(-1, keepdim=True)
    cdf = torch.cumsum(pn, dim=-1)
    want = torch.clamp((u.cpu()[:, None] >= cdf).sum(-1), max=4)
    close = (torch.abs(u.cpu()[:, None] - cdf) < 1e-6).any(-1)        # ties at a CDF edge may differ by rounding
    assert torch.equal(a.cpu().long()[~close], want[~close])
    assert len(set(a.cpu().tolist())) == 5
    assert torch.allclose(lp.cpu(), pp.categorical_log_prob(p_cpu, a.cpu().long()), rtol=1e-5, atol=1e-6)
    # greedy = argmax (evaluate_with_lstm.py:65); forced actions replay a trace
    ag, _, _, pg = model.act(x, greedy=True)
    assert torch.equal(ag.cpu().long(), pg.cpu().argmax(-1))
    forced = torch.randint(0, 5, (B,), dtype=torch.int32)
    af, lpf, _, pf = model.act(x, forced_actions=forced)
    assert torch.equal(af.cpu(), forced)
    assert torch.allclose(lpf.cpu(), pp.categorical_log_prob(pf.cpu(), forced.long()), rtol=1e-5, atol=1e-6)
    # Philox action stream keyed by the env state is deterministic and sharding independent
    env = pb().VecMethaneEnv(256, seed=4, field_mode="procedural")
    o = env.observe()
    a1 = model.act(o, env=env)[0].clone()
    a2 = model.act(o, env=env)[0]
    assert torch.equal(a1, a2)


def test_nan_raises_like_reference():
    model = pb().PPOActorCritic(device="cuda")
    with torch.no_grad():
        model.actor.bias[2] = float("nan")
    with pytest.raises(RuntimeError, match="NaN in model output"):
        model(torch.rand(4, 6).cuda())


def test_lstm_stop_head_golden():
    g = load_golden("lstm_s7.npz")
    head = pb().PeakAndStopPredictor(device="cuda")
    head.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
    peak, stop = head(torch.from_numpy(g["windows"]).cuda().unsqueeze(-1))
    assert np.allclose(peak.cpu().numpy(), g["peak"], rtol=1e-5, atol=2e-6)
    assert np.allclose(stop.cpu().numpy(), g["stop_prob"], rtol=1e-5, atol=1e-6)
    clear = np.abs(g["stop_prob"] - 0.8) > 1e-5
    assert np.array_equal((stop.cpu().numpy() > 0.8)[clear], (g["stop_prob"] > 0.8)[clear])
    assert clear.sum() > 90


@pytest.mark.parametrize("hidden,batch,steps", [(32, 1, 20), (32, 257, 20), (64, 100, 20), (32, 40, 7), (48, 33, 20),
                                                (256, 50, 20)])
def test_lstm_stop_head_vs_oracle(hidden, batch, steps):
    torch.manual_seed(hidden + batch)
    ora = pp.OraclePeakAndStop(hidden_dim=hidden)
    head = pb().PeakAndStopPredictor(hidden_dim=hidden, device="cuda")
    head.load_state_dict(ora.state_dict())
    x = torch.rand(batch, steps, 1)
    with torch.no_grad():
        p_ref, s_ref = ora(x)
    p, s = head(x.cuda())
    assert torch.allclose(p.cpu(), p_ref, rtol=1e-5, atol=2e-6)
    assert torch.allclose(s.cpu(), s_ref, rtol=1e-5, atol=1e-6)


def test_generic_lstm_stack_vs_oracle():
    """V2.0 ConcentrationThresholdPredictor LSTM stack (3 x 128, window 10)."""
    torch.manual_seed(0)
    ora = pp.OracleThresholdPredictor().eval()
    m = pb().ConcentrationThresholdPredictor(device="cuda")
    m.load_state_dict(ora.state_dict())
    x = torch.rand(19, 10, 1)
    with torch.no_grad():
        out, _ = ora.lstm(x)
    h = m.lstm_last_hidden(x.cuda())
    assert torch.allclose(h.cpu(), out[:, -1], rtol=1e-4, atol=2e-6)


def test_trend_features_golden():
    import ctypes as C
    g = load_golden("trend_s9.npz")
    m = pb()
    lib = m._lib.load()
    conc = torch.from_numpy(g["conc"].astype(np.float32)).cuda().contiguous()
    pos = torch.from_numpy(g["pos"].astype(np.float32)).cuda().contiguous()
    src = torch.from_numpy(g["src"]).cuda().contiguous()
    out = torch.zeros(32, 4, device="cuda")
    rc = lib.plume_trend_features(conc.data_ptr(), 32, 20, pos.data_ptr(), src.data_ptr(), 100.0, out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert np.allclose(out.cpu().numpy()[:, 0], g["label"], rtol=1e-5, atol=1e-6)

"""Debug helper: per-tensor error of the tensor-core PPO gradient against autograd (float64)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import uav_wrf_les_ppo_lstm_b200 as m
from oracle import plume_oracle as po, ppo_oracle as pp

def run(M, mb_start, mb_size, path):
    os.environ["PLUME_PPO_PATH"] = path
    cfg = po.config_for("2.1")
    torch.manual_seed(M)
    ora = pp.OracleActorCritic()
    with torch.no_grad():
        for p in ora.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        ora.actor.weight.mul_(30.0)
    model = m.PPOActorCritic(device="cuda")
    model.load_state_dict(ora.state_dict())
    rng = np.random.default_rng(M)
    S = torch.from_numpy(rng.random((M, 6)).astype(np.float32))
    with torch.no_grad():
        for _ in range(50):
            y1 = ora.feature[:2](S)
            y2 = ora.feature[:5](S)
            risky = (y1.abs().min(1).values < 2e-5) | (y2.abs().min(1).values < 2e-5)
            if not bool(risky.any()):
                break
            S[risky] = torch.from_numpy(rng.random((int(risky.sum()), 6)).astype(np.float32))
    A = torch.from_numpy(rng.integers(0, 5, M))
    with torch.no_grad():
        P, V0 = ora(S)
    LP = pp.categorical_log_prob(P, A) + torch.from_numpy((0.25 * rng.normal(size=M)).astype(np.float32))
    ADV = torch.from_numpy(rng.normal(size=M).astype(np.float32))
    V = V0.squeeze(-1) + torch.from_numpy((0.3 * rng.normal(size=M)).astype(np.float32))
    RET = V + torch.from_numpy(rng.normal(size=M).astype(np.float32))
    perm = torch.randperm(M)
    idx = perm[mb_start:mb_start + mb_size]
    ora64 = pp.OracleActorCritic().double()
    ora64.load_state_dict({k: v.double() for k, v in ora.state_dict().items()})
    ora64.zero_grad()
    tot = pp.ppo_loss(ora64, S[idx].double(), A[idx], LP[idx].double(), ADV[idx].double(), RET[idx].double(), V[idx].double(), cfg)
    tot[0].backward()
    ora.zero_grad()
    t32 = pp.ppo_loss(ora, S[idx], A[idx], LP[idx], ADV[idx], RET[idx], V[idx], cfg)
    t32[0].backward()
    named32 = dict(ora.named_parameters())
    lib = m._lib.load()
    dev = "cuda"
    t = lambda x, dt: x.to(dt).to(dev).contiguous()
    obs, act, lp, adv, ret, val = t(S, torch.float32), t(A, torch.int32), t(LP, torch.float32), t(ADV, torch.float32), t(RET, torch.float32), t(V, torch.float32)
    batch = m._lib.PpoBatch(M, obs.data_ptr(), act.data_ptr(), lp.data_ptr(), adv.data_ptr(), ret.data_ptr(), val.data_ptr())
    ws = m.UpdateWorkspace(dev, mb_size)
    loss = torch.zeros(4, dtype=torch.float64, device=dev)
    model.flat_grad.zero_()
    permd = perm.to(dev)
    rc = lib.plume_ppo_grad(model.flat.data_ptr(), C.byref(batch), permd.data_ptr(), 0, 0, mb_start, mb_size, mb_size,
                            cfg.clip_epsilon, cfg.entropy_beta, model.flat_grad.data_ptr(), loss.data_ptr(),
                            ws.nan_flag.data_ptr(), ws.ws.data_ptr(), ws.bytes, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.plume_last_error()
    torch.cuda.synchronize()
    print(f"--- M={M} mb={mb_size} path={path} loss gpu={loss.cpu().numpy()} ref={[float(x) for x in tot]}")
    named = dict(ora64.named_parameters())
    gn = float(torch.sqrt(sum((p.grad ** 2).sum() for p in named.values())))
    for name, (off, shape) in m._lib.MLP_OFFSETS.items():
        n = int(np.prod(shape))
        g_gpu = model.flat_grad[off:off + n].view(shape).cpu().double()
        g_ref = named[name].grad
        e32 = (named32[name].grad.double() - g_ref).abs().max() / g_ref.abs().max()
        if path == "tc" and name in ("feature.4.bias", "feature.3.bias", "feature.4.weight", "feature.0.bias"):
            d = (g_gpu - g_ref).abs().flatten()
            top = torch.topk(d, 6)
            print("     top errors:", [(int(i), f"{float(v):.2e}", f"ref={float(g_ref.flatten()[i]):+.2e}", f"gpu={float(g_gpu.flatten()[i]):+.2e}") for v, i in zip(top.values, top.indices)])
        print(f"     l2: |d|/|g_all| = {float((g_gpu - g_ref).norm()) / gn:.2e}  |g_t|/|g_all| = {float(g_ref.norm()) / gn:.2e}  torch32 {float((named32[name].grad.double() - g_ref).norm()) / gn:.2e}")
        print(f"  {name:18s} max|ref|={g_ref.abs().max():.3e} relerr={(g_gpu - g_ref).abs().max() / g_ref.abs().max():.2e} torch32={e32:.2e}")

for args in [(60000, 5000, 50001)]:
    for path in ("cuda", "tc"):
        run(*args, path)

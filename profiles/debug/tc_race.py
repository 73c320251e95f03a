"""Debug helper: run the tensor-core PPO gradient twice on the same minibatch and compare bitwise."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["PLUME_PPO_PATH"] = "tc"
import uav_wrf_les_ppo_lstm_b200 as m
from oracle import plume_oracle as po
cfg = po.config_for("2.1")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(1)
model = m.PPOActorCritic(device="cuda")
rng = np.random.default_rng(1)
dev = "cuda"
obs = torch.from_numpy(rng.random((M, 6)).astype(np.float32)).to(dev)
act = torch.from_numpy(rng.integers(0, 5, M).astype(np.int32)).to(dev)
lp = torch.from_numpy((-1.6 + 0.1 * rng.normal(size=M)).astype(np.float32)).to(dev)
adv = torch.from_numpy(rng.normal(size=M).astype(np.float32)).to(dev)
ret = torch.from_numpy(rng.normal(size=M).astype(np.float32)).to(dev)
val = torch.from_numpy(rng.normal(size=M).astype(np.float32)).to(dev)
batch = m._lib.PpoBatch(M, obs.data_ptr(), act.data_ptr(), lp.data_ptr(), adv.data_ptr(), ret.data_ptr(), val.data_ptr())
ws = m.UpdateWorkspace(dev, M)
lib = m._lib.load()
outs = []
for r in range(reps):
    loss = torch.zeros(4, dtype=torch.float64, device=dev)
    model.flat_grad.zero_()
    rc = lib.plume_ppo_grad(model.flat.data_ptr(), C.byref(batch), None, 0, 0, 0, M, M, cfg.clip_epsilon, cfg.entropy_beta,
                            model.flat_grad.data_ptr(), loss.data_ptr(), ws.nan_flag.data_ptr(), ws.ws.data_ptr(), ws.bytes,
                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    outs.append(model.flat_grad.clone())
for r in range(1, reps):
    d = (outs[r] - outs[0]).abs()
    print(f"run {r} vs 0: max abs diff {d.max().item():.3e} (max |g| {outs[0].abs().max().item():.3e}) at {int(d.argmax())}")
    for name, (off, shape) in m._lib.MLP_OFFSETS.items():
        n = int(np.prod(shape))
        print(f"   {name:18s} {d[off:off+n].max().item():.3e} / {outs[0][off:off+n].abs().max().item():.3e}")

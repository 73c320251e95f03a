"""Run-to-run differences of ONE gradient launch (tcgen05 path): python profiles/debug/grad_probe.py [lib.so] [M] [runs]"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import uav_wrf_les_ppo_lstm_b200 as m
if len(sys.argv) > 1 and sys.argv[1] != '-':
    m._lib.LIB_PATH = sys.argv[1]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cfg = m.config_for("2.1")
lib = m._lib.load()
torch.manual_seed(3)
model = m.PPOActorCritic(device="cuda")
rng = np.random.default_rng(11)
dev = "cuda"
obs = torch.from_numpy(rng.random((M, 6)).astype(np.float32)).to(dev)
act = torch.from_numpy(rng.integers(0, 5, M).astype(np.int32)).to(dev)
lp = torch.from_numpy((-1.6 + 0.1 * rng.normal(size=M)).astype(np.float32)).to(dev)
adv = torch.from_numpy(rng.normal(size=M).astype(np.float32)).to(dev)
ret = torch.from_numpy(rng.normal(size=M).astype(np.float32)).to(dev)
val = torch.from_numpy(rng.normal(size=M).astype(np.float32)).to(dev)
batch = m._lib.PpoBatch(M, obs.data_ptr(), act.data_ptr(), lp.data_ptr(), adv.data_ptr(), ret.data_ptr(), val.data_ptr())
perm = torch.randperm(M).to(dev)
ws = m.UpdateWorkspace(dev, M)
offs = [("W1", 0), ("B1", 1536), ("G1", 1792), ("BE1", 2048), ("W2", 2304), ("B2", 35072), ("G2", 35200), ("BE2", 35328),
        ("WA", 35456), ("BA", 36096), ("WC", 36104), ("BC", 36232), ("end", 36236)]
def name(i):
    for (n, o), (_, o2) in zip(offs[:-1], offs[1:]):
        if o <= i < o2:
            j = i - o
            return f"{n}[{j // 256},{j % 256}]" if n == "W2" else f"{n}[{j}]"
res = []
loss = torch.zeros(4, dtype=torch.float64, device=dev)
for r in range(runs):
    model.flat_grad.zero_()
    loss.zero_()
    rc = lib.plume_ppo_grad(model.flat.data_ptr(), C.byref(batch), perm.data_ptr(), 0, 0, 0, M, M, cfg.clip_epsilon,
                            cfg.entropy_beta, model.flat_grad.data_ptr(), loss.data_ptr(), ws.nan_flag.data_ptr(),
                            ws.ws.data_ptr(), ws.bytes, 1, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.plume_last_error()
    # a little unrelated work between launches varies the timing
    if r % 3 == 1:
        torch.empty(1 << 22, device=dev).normal_()
    res.append(model.flat_grad.detach().cpu().clone())
ref = torch.stack(res).median(dim=0).values
scale = float(ref.abs().max())
bad = 0
for k, r in enumerate(res):
    d = (r - ref).abs()
    big = (d > 1e-5 * scale).nonzero().flatten()
    if len(big):
        bad += 1
        names = [name(int(i)) for i in big[:12]]
        print(f"run {k}: {len(big)} entries off by up to {float(d.max()) / scale:.2e} of the largest gradient: {' '.join(names)}")
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'product'}: {bad} of {runs} launches deviate (largest gradient {scale:.3e})")

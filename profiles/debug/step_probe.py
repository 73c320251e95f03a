"""Per-optimiser-step gradients of repeated identical updates (python loop, one plume_ppo_grad per step): where and by how much
does a run first depart from the majority?  python profiles/debug/step_probe.py [lib.so] [runs]"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import uav_wrf_les_ppo_lstm_b200 as m
if len(sys.argv) > 1 and sys.argv[1] != '-':
    m._lib.LIB_PATH = sys.argv[1]
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
from test_gpu_learner import _fill_buffer
cfg = m.config_for("2.1")
T, N, mb = 128, 512, 16384
buf = _fill_buffer(m, T, N, 11)
torch.manual_seed(3)
init = m.PPOActorCritic(device="cuda").flat.clone()
lib = m._lib.load()
dev = "cuda"
M = T * N
def name(i):
    if 2304 <= i < 35072:
        j = i - 2304
        return f"W2[{j // 256},{j % 256}]"
    return f"p[{i}]"
all_g = []
for r in range(runs):
    model = m.PPOActorCritic(device=dev)
    model.flat.data.copy_(init)
    opt = m.FusedAdam(model, lr=cfg.learning_rate)
    ws = m.UpdateWorkspace(dev, mb)
    m.compute_advantages(buf, cfg, ws, None)
    batch = m._lib.PpoBatch(M, buf.obs.data_ptr(), buf.actions.data_ptr(), buf.log_probs.data_ptr(), buf.advantages.data_ptr(),
                            buf.returns.data_ptr(), buf.values.data_ptr(), None)
    gs = []
    loss = torch.zeros(4, dtype=torch.float64, device=dev)
    for epoch in range(cfg.epochs):
        for start in range(0, M, mb):
            opt.zero_grad()
            rc = lib.plume_ppo_grad(model.flat.data_ptr(), C.byref(batch), None, 5, epoch, start, mb, mb, cfg.clip_epsilon,
                                    cfg.entropy_beta, model.flat_grad.data_ptr(), loss.data_ptr(), ws.nan_flag.data_ptr(),
                                    ws.ws.data_ptr(), ws.bytes, 1, torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            gs.append(model.flat_grad.detach().clone())
            opt.step()
    all_g.append(torch.stack(gs).cpu())
G = torch.stack(all_g)                       # [runs, steps, params]
ref = G.median(dim=0).values
for r in range(runs):
    d = (G[r] - ref).abs()
    per_step = d.max(dim=1).values
    first = int((per_step > 1e-9).nonzero()[0]) if bool((per_step > 1e-9).any()) else -1
    if first >= 0:
        i = int(d[first].argmax())
        row = (i - 2304) // 256 if 2304 <= i < 35072 else -1
        n_row = int((d[first][2304 + row * 256: 2304 + (row + 1) * 256] > 1e-10).sum()) if row >= 0 else 0
        n_all = int((d[first] > 1e-10).sum())
        print(f"run {r}: first departs at step {first}: {name(i)} ref {float(ref[first, i]):.6e} got {float(G[r, first, i]):.6e} "
              f"(diff {float(d[first, i]):.2e}); entries off > 1e-10: {n_all}, in that W2 row: {n_row}; "
              f"B2[row] diff {float(d[first][35072 + row]) if row >= 0 else 0:.2e}")
print("largest |g| at step 0:", float(ref[0].abs().max()))

"""Which parameter differs between two runs of the same update (test_packed_records...)?  python profiles/debug/flaky_probe.py [lib.so]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import uav_wrf_les_ppo_lstm_b200 as m
if len(sys.argv) > 1:
    m._lib.LIB_PATH = sys.argv[1]
from test_gpu_learner import _fill_buffer
cfg = m.config_for("2.1")
T, N, mb = 128, 512, 16384
buf = _fill_buffer(m, T, N, 11)
torch.manual_seed(3)
init = m.PPOActorCritic(device="cuda").flat.clone()
offs = [("W1", 0), ("B1", 1536), ("G1", 1792), ("BE1", 2048), ("W2", 2304), ("B2", 35072), ("G2", 35200), ("BE2", 35328),
        ("WA", 35456), ("BA", 36096), ("WC", 36104), ("BC", 36232), ("end", 36236)]
def name(i):
    for (n, o), (_, o2) in zip(offs[:-1], offs[1:]):
        if o <= i < o2:
            return f"{n}[{i - o}]"
res = []
for rep in range(24):
    thr = m.learner.MATERIALISE_PERM_MIN if rep % 2 == 0 else 1 << 62
    saved = m.learner.MATERIALISE_PERM_MIN
    m.learner.MATERIALISE_PERM_MIN = thr
    model = m.PPOActorCritic(device="cuda")
    model.flat.data.copy_(init)
    opt = m.FusedAdam(model, lr=cfg.learning_rate)
    m.update_model(buf, model, opt, cfg=cfg, minibatch_size=mb, workspace=m.UpdateWorkspace("cuda", mb), perm_seed=5)
    m.learner.MATERIALISE_PERM_MIN = saved
    res.append(model.flat.detach().cpu().clone())
# the majority result is the reference; report every run that is further than 1e-7 from it
ref = torch.stack(res).median(dim=0).values
bad = 0
for k, r in enumerate(res):
    d = (r - ref).abs()
    if float(d.max()) > 1e-7:
        bad += 1
        top = torch.topk(d, 4)
        print(f"run {k}: " + " ".join(f"{name(int(i))}:{float(v):.2e}" for v, i in zip(top.values, top.indices)))
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'product'}: {bad} of {len(res)} runs deviate")

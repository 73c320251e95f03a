"""N3 training kernel alone (2000 windows of 20, minibatch 64) for ncu: plain run first, then under
`ncu --set full -k regex:lstm_train_kernel`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n, T = 2048, 20
    g = torch.Generator().manual_seed(0)
    feats = torch.rand(n, T, generator=g).cuda()
    labels = torch.stack([torch.rand(n, generator=g), (torch.rand(n, generator=g) < 0.3).float()], dim=1).cuda()
    head = pb.PeakAndStopPredictor(device="cuda")
    tr = pb.LstmTrainer(head, feats, labels, batch_size=B)
    tr.train_epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr.train_epoch()
    e1.record()
    torch.cuda.synchronize()
    steps = 5 * tr.n_batches
    print(f"profile_lstm_train batch={B} {1e3 * e0.elapsed_time(e1) / steps:.2f} us/optimiser step, loss {tr.history[-1][0]:.4f}")


if __name__ == "__main__":
    main()

"""Time cm_ctc_loss against torch's CTC kernels (loss + gradient) at the bench shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mamba_asr_b200.ctc import ctc_loss

dev = torch.device("cuda:0")
for (Bt, T, Cn, S) in [(64, 501, 31, 60), (64, 501, 5000, 60), (128, 251, 31, 30), (16, 1501, 31, 200)]:
    g = torch.Generator().manual_seed(0)
    lp = F.log_softmax(torch.randn(T, Bt, Cn, generator=g), -1).to(dev)
    tg = torch.randint(1, Cn, (Bt, S), generator=g).to(dev)
    il = torch.full((Bt,), T, dtype=torch.long, device=dev)
    tl = torch.full((Bt,), S, dtype=torch.long, device=dev)
    for name, fn in (("cm", ctc_loss), ("torch", F.ctc_loss)):
        def run():
            x = lp.detach().requires_grad_(True)
            loss = fn(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=True)
            loss.backward()
            return loss
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            l = run()
        b.record()
        torch.cuda.synchronize()
        print("%-6s B=%d T=%d C=%d S=%d  %.3f ms  loss=%.5f" % (name, Bt, T, Cn, S, a.elapsed_time(b) / 20, l.item()), flush=True)

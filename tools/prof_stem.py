"""Time the CNN front-end (forward + backward) with and without the one-kernel first block at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mamba_asr_b200.encoder import ConvFrontEnd

dev = torch.device("cuda:0")
torch.manual_seed(0)
fe = ConvFrontEnd(80).to(dev)
for (Bt, T) in [(64, 2001), (32, 1001)]:
    feats = torch.randn(Bt, T, 80, device=dev)
    for name, env in (("stem", None), ("cudnn+ln_act", "1")):
        if env is None:
            os.environ.pop("CM_NO_FUSE_STEM", None)
        else:
            os.environ["CM_NO_FUSE_STEM"] = env
        def run():
            fe.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = fe(feats)
            return y
        for _ in range(3):
            run().float().sum().backward()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(10):
            ev[0].record()
            y = run()
            ev[1].record()
            g = torch.ones_like(y)
            torch.cuda.synchronize()
            ev[0].synchronize()
            s = torch.cuda.Event(enable_timing=True); s.record()
            y.backward(g)
            ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]); tb += s.elapsed_time(ev[2])
        print("%-14s B=%d T=%d  fwd %.3f ms  bwd %.3f ms" % (name, Bt, T, tf / 10, tb / 10), flush=True)

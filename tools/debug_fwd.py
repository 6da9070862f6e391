import os, sys, itertools, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_asr_b200 import kernels as K
dev="cuda"
def mk(Bt,D,L,R,dt,sliced):
    N=16; P=R+2*N
    g=torch.Generator(device=dev).manual_seed(0)
    rn=lambda *s: torch.randn(*s,device=dev,generator=g)
    cl=lambda: rn(Bt,L,D).to(dt).transpose(1,2)
    dirs=[]
    for rev in (False,True):
        if sliced:
            xd=rn(Bt,L,P).to(dt); Bm=xd[...,R:R+N].transpose(1,2); Cm=xd[...,R+N:].transpose(1,2)
        else:
            Bm=rn(Bt,L,N).to(dt).transpose(1,2); Cm=rn(Bt,L,N).to(dt).transpose(1,2)
        dirs.append(dict(u=cl(),delta=(0.5*rn(Bt,L,D)).to(dt).transpose(1,2),A=-torch.exp(0.3*rn(D,N)),B=Bm,C=Cm,D=torch.ones(D,device=dev),delta_bias=torch.full((D,),-4.0,device=dev),reverse=rev))
    return dirs, cl()
for (Bt,D,L,R) in [(2,64,37,2),(2,64,45,4),(4,128,100,8)]:
  for dt in (torch.float32, torch.bfloat16):
    for sliced in (False,True):
        dirs,z=mk(Bt,D,L,R,dt,sliced)
        for lanes in (1,2,4):
            base=K.scan_forward(dirs,z=z,out_scale=0.5,delta_softplus=True,lanes=lanes)["out"].float()
            for ck,op in itertools.product((False,True),(False,True)):
                r=K.scan_forward(dirs,z=z,out_scale=0.5,delta_softplus=True,need_ckpt=ck,need_out_pre=op,lanes=lanes)
                torch.cuda.synchronize()
                e=(r["out"].float()-base).abs().max().item()
                flag = "" if e==0 else "  <<<<<< MISMATCH"
                print(Bt,D,L,R,str(dt)[6:],"sliced" if sliced else "contig","lanes",lanes,"ckpt",ck,"out_pre",op,"maxdiff",e,flag)

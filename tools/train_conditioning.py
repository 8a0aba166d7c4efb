"""How well-conditioned is the training step as a parity target?  CPU only, PyTorch only.

Runs the oracle forward in TRAIN mode (batch-statistics BatchNorm) + losses + backward twice in fp32 -- once as is, once
with every convolution / BatchNorm output rounded to bf16 in the forward (straight-through gradient) -- and prints how far
the head outputs and the parameter gradients of the two runs are apart.  Finding (profiles/r02_train_parity_conditioning.txt):
in train mode the random-weight network is chaotic -- the bf16 rounding of activations alone moves the stage-4 features by
~70 % rms and the backbone gradients by > 100 % (relative L2), at any batch size, with synthetic or default-init weights,
with noise or structured images; in EVAL mode (running statistics) the same rounding moves the outputs by < 1 % and the
gradients by ~2 %.  A whole-step gradient comparison against fp32 therefore cannot separate an implementation error from
rounding; the training parity tests compare operator by operator, check shallow sub-networks, and bound the whole-step
error by THIS emulation's error (tests/test_gpu_train.py).

  python tools/train_conditioning.py SIZE BATCH real|smooth synth|init [CALM] [struct]     (EVAL=1: eval-mode BatchNorm)
"""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, types, os
import torch.nn.functional as F
import hydranet_b200 as hb
from hydranet_b200 import losses
from hydranet_b200.config import big_cfg
from oracle import hydranet_ref, synth, train_golden
H=W=int(sys.argv[1]); B=int(sys.argv[2]); LOSS=sys.argv[3]; WTS=sys.argv[4]
cfg=big_cfg(W,H)
torch.manual_seed(0)
m=hb.HydraNet(cfg)
if WTS=='synth': m.load_state_dict(synth.synth_state_dict(m.state_dict(),seed=1,seg_logit_gain=1.0))
sd0={k:v.detach().clone() for k,v in m.state_dict().items()}
CALM=float(sys.argv[5]) if len(sys.argv)>5 else 1.0
for k in sd0:
    if k.endswith('conv_block_3.1.weight'): sd0[k]*=CALM
x=synth.synth_input(B,H,W,seed=3)
if len(sys.argv)>6 and sys.argv[6]=='struct':
    g=torch.Generator().manual_seed(11)
    x=sum(F.interpolate(torch.randn(B,3,max(H//s,1),max(W//s,1),generator=g),size=(H,W),mode='bilinear',align_corners=False)*a for s,a in ((64,1.0),(16,0.7),(4,0.5),(1,0.3)))
fh,fw=H//32,W//32
gt=train_golden.synthetic_gt(B,H,W,fh,fw,int(H/8),seed=5)
ppl=int(H/8)
def total(out):
    if LOSS=='smooth':
        g=torch.Generator().manual_seed(7)
        ts=[out['seg'],out['detection']['regression'],out['detection']['classification'],out['lane']['predict_cls'],out['lane']['predict_loc']]
        return sum((t*torch.randn(t.shape,generator=g)).sum()/t.numel()**0.5 for t in ts)
    sc=cfg["segment"]
    s=losses.seg_loss(out["seg"],gt["gt_seg"].long(),torch.tensor(sc["class_weight"]),True,0.3,False)
    c,r=losses.detection_loss(out["detection"]["classification"],out["detection"]["regression"],out["detection"]["anchors"],gt["gt_det"])
    pos,neg,pm,pn=losses.lane_cls_loss(gt["gt_cls"],out["lane"]["predict_cls"])
    loc=losses.lane_reg_loss(pm,pn,gt["gt_loc"],out["lane"]["predict_loc"],points_per_line=ppl)
    return 5*s+c.mean()+50*r.mean()+pos+neg+loc
def run(quant):
    sd={k:(v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone()) for k,v in sd0.items()}
    ste=lambda t: t+(t.to(torch.bfloat16).float()-t).detach()
    oc,ob=F.conv2d,F.batch_norm
    if quant:
        Fq=types.SimpleNamespace(**{k:getattr(F,k) for k in dir(F) if not k.startswith('__')})
        Fq.conv2d=lambda *a,**k: ste(oc(*a,**k)); Fq.batch_norm=lambda *a,**k: ste(ob(*a,**k))
        hydranet_ref.F=Fq
    try:
        out=hydranet_ref.forward(sd,cfg,x,want_feats=True,train=(os.environ.get("EVAL")!="1"))
        FE.append([t.detach() for t in out['_feats']]+[t.detach() for t in out['_fused']])
        t=total(out); t.backward()
    finally:
        hydranet_ref.F=F
    OUTS.append([out['seg'].detach(),out['detection']['regression'].detach(),out['detection']['classification'].detach(),out['lane']['predict_cls'].detach(),out['lane']['predict_loc'].detach()])
    return float(t),{k:v.grad for k,v in sd.items() if isinstance(v,torch.Tensor) and v.grad is not None}
OUTS=[];FE=[]
t0,g0=run(False); t1,g1=run(True)
for i,(a,b) in enumerate(zip(FE[0],FE[1])): print('feat',i,tuple(a.shape),'rms-rel %.4f'%float((a-b).norm()/a.norm()))
for n,a,b in zip(['seg','reg','cls','lane_cls','lane_loc'],OUTS[0],OUTS[1]): print('fwd',n,'max-rel %.4f'%float((a-b).abs().max()/a.abs().max()),'rms-rel %.4f'%float((a-b).norm()/a.norm()))
print("loss",t0,t1)
import collections
d=collections.defaultdict(list)
for k in g0:
    e=float((g0[k]-g1[k]).norm()/g0[k].norm().clamp_min(1e-12))
    d['.'.join(k.split('.')[:3])].append(e)
allv=sorted(e for v in d.values() for e in v); print("ALL median %.3f  p10 %.3f p90 %.3f"%(allv[len(allv)//2],allv[len(allv)//10],allv[len(allv)*9//10]))
for k in sorted(d):
    v=sorted(d[k]); print("%-45s n=%3d median %.3f min %.3f max %.3f"%(k,len(v),v[len(v)//2],v[0],v[-1]))

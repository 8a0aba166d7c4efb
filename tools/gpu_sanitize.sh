#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over a small forward (big cfg 128x128, batch 2) + the decoders
mkdir -p gpurun_out
cat > /tmp/san.py <<'P'
import sys, torch
sys.path.insert(0, '.')
import hydranet_b200 as hb
from hydranet_b200.config import big_cfg
from oracle import synth
cfg = big_cfg(128, 128)
m = hb.HydraNet(cfg).eval().cuda()
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1, seg_logit_gain=20.0))
x = synth.synth_input(2, 128, 128, seed=3).cuda()
codec = hb.LaneCodec(128, 128, cfg["lane"]["anchor_stride"], int(128 / cfg["lane"]["interval"]), True, 1, True)
m.fuse_postprocess(det=(0.3, 0.3), lane=(codec, 0.3, 100, False))
with torch.no_grad():
    out = m(x)
    d, l = m.postprocess_results()
torch.cuda.synchronize()
print("ok", out["seg"].shape, int(d[3].sum()), int(l[0].sum()))
frames = torch.randint(0, 256, (2, 90, 160, 3), dtype=torch.uint8, device="cuda")
y = hb.preprocess(frames, (128, 128)); torch.cuda.synchronize(); print("pre ok", y.shape)
P
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitize_memcheck.log
for tool in racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  tail -4 gpurun_out/sanitize_$tool.log
done

"""Where does the seg-argmax disagreement come from?  Runs the engine's op list in the CPU emulator (tests/emulator.py)
with bf16 storage switched on per subsystem (activations and / or weights) and compares the seg arg-max with the fp32
oracle.  CPU only; writes a table to stdout (kept under profiles/).

  python tools/precision_study.py [--size 320] [--weights synth|init] [--seed 1]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

import emulator
import hydranet_b200 as hb
from hydranet_b200.config import big_cfg
from hydranet_b200.engine import Builder
from oracle import hydranet_ref, synth

PHASES = ("backbone", "neck", "seg", "seg.t", "seg.o")
SEG = ("seg", "seg.t", "seg.o")


class StudyBuilder(Builder):
    """Builder whose storage dtype depends on the subsystem being emitted and on whether a weight or an activation is
    being allocated.  cfgmap: {(phase, 'act'|'w'): dtype}."""

    def __init__(self, cfgmap, *a, **k):
        self._cfgmap, self._phase, self._kind, self._seg_bufs = cfgmap, "backbone", "w", 0
        super().__init__(*a, **k)

    @property
    def dt(self):
        return self._cfgmap.get((self._phase, self._kind), torch.float32)

    @dt.setter
    def dt(self, v):
        pass

    def neck(self, feats):
        self._phase = "neck"
        return super().neck(feats)

    def seg_head(self, *a):
        self._phase = "seg"
        return super().seg_head(*a)

    # seg sub-phases: "seg" = decoder.0-5, "seg.t" = decoder.6-7 (160^2 / 320^2 maps), "seg.o" = decoder.8 (logits).
    # A layer's phase decides its weights and its OUTPUT buffer, so "seg.o" in fp32 still reads decoder.7's rounded output.
    def conv3x3_plain(self, name, *a, **k):
        self._phase = "seg.t" if name in ("seg.d6",) else "seg"
        return super().conv3x3_plain(name, *a, **k)

    def conv3x3_up(self, name, *a, **k):
        self._phase = "seg.t" if name in ("seg.d7",) else "seg"
        return super().conv3x3_up(name, *a, **k)

    def seg_out(self, *a, **k):
        self._phase = "seg.o"
        return super().seg_out(*a, **k)

    def buf(self, *a, **k):
        self._kind = "act"
        try:
            ph = self._phase
            if ph.startswith("seg"):  # seg_head allocates its 8 activation buffers in layer order: decoder.0 .. decoder.7
                self._phase = "seg.t" if self._seg_bufs >= 6 else "seg"
                self._seg_bufs += 1
            return super().buf(*a, **k)
        finally:
            self._kind = "w"
            self._phase = ph

    def squeeze_excite(self, name, g, se):
        self._kind = "se"  # mean / hidden / gate vectors
        try:
            return super().squeeze_excite(name, g, se)
        finally:
            self._kind = "w"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--weights", default="synth")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--configs", default="")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    S = args.size
    cfg = big_cfg(S, S)
    cfg["train"]["train_detect"] = False
    cfg["train"]["train_lane"] = False
    torch.manual_seed(args.seed)
    m = hb.HydraNet(cfg).eval()
    if args.weights == "synth":
        m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=args.seed, seg_logit_gain=20.0))
    sd = m.state_dict()
    x = synth.synth_input(1, S, S, seed=2)
    with torch.no_grad():
        ref = hydranet_ref.forward(sd, cfg, x)["seg"]
    ref_cls = ref.argmax(1)
    top2 = torch.topk(ref, 2, dim=1).values
    gap = (top2[:, 0] - top2[:, 1]) / ref.abs().max()
    print("# %s weights seed %d, %dx%d, max|logit| %.4f; top1-top2 gap / max|logit| quantiles 0.1%%/1%%/10%%: %.2e %.2e %.2e" % (
        args.weights, args.seed, S, S, float(ref.abs().max()), *[float(torch.quantile(gap.flatten()[:4000000], q)) for q in (0.001, 0.01, 0.1)]))
    bf, f32 = torch.bfloat16, torch.float32

    def mk(act=(), w=(), se=None):
        se = ("backbone" in act) if se is None else se
        d = {}
        for p in act:
            d[(p, "act")] = bf
        for p in w:
            d[(p, "w")] = bf
        if se:
            d[("backbone", "se")] = bf
        return d

    configs = [
        ("all fp32", mk()),
        ("all bf16 (round-1 engine)", mk(PHASES, PHASES)),
        ("all bf16, SE mean/gate fp32", mk(PHASES, PHASES, se=False)),
        ("bf16 weights only", mk((), PHASES)),
        ("bf16 activations only", mk(PHASES, ())),
        ("bf16 backbone only", mk(("backbone",), ("backbone",))),
        ("bf16 neck only", mk(("neck",), ("neck",))),
        ("bf16 seg only", mk(SEG, SEG)),
        ("bf16 backbone+neck, fp32 seg", mk(("backbone", "neck"), ("backbone", "neck"))),
        ("bf16 seg weights only", mk((), SEG)),
        ("bf16 seg activations only", mk(SEG, ())),
        ("bf16 seg decoder.0-5 only", mk(("seg",), ("seg",))),
        ("bf16 seg decoder.6-7 only", mk(("seg.t",), ("seg.t",))),
        ("bf16 seg decoder.8 weights only", mk((), ("seg.o",))),
        ("all bf16, decoder.8 w fp32", mk(PHASES, ("backbone", "neck", "seg", "seg.t"))),
        ("all bf16, decoder.7 out + 8 w fp32", mk(("backbone", "neck", "seg"), ("backbone", "neck", "seg", "seg.t"))),
        ("all bf16, decoder.6-8 fp32", mk(("backbone", "neck", "seg"), ("backbone", "neck", "seg"))),
    ]
    if args.configs:
        keep = set(int(i) for i in args.configs.split(","))
        configs = [c for i, c in enumerate(configs) if i in keep]
    print("%-34s %10s %12s %12s" % ("configuration", "agreement", "max rel err", "mean rel err"))
    for name, cm in configs:
        t0 = time.time()
        with torch.no_grad():
            b = StudyBuilder(cm, m, 1, S, S, torch.device("cpu")).build(x.clone())
            emulator.run_ops(b.ops)
        out = b.out["seg"].float()
        agree = float((out.argmax(1) == ref_cls).float().mean())
        err = (out - ref).abs()
        print("%-34s %10.5f %12.3e %12.3e   (%.0f s)" % (name, agree, float(err.max() / ref.abs().max()), float(err.mean() / ref.abs().max()), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()

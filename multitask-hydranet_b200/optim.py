"""Optimizer of the training step: Adam over every parameter tensor in ONE kernel launch (``hn_adam_step``).

Semantics are ``torch.optim.Adam``'s as the reference configures it (model/train.py:147: ``Adam(params, lr, weight_decay=wd)``,
betas (0.9, 0.999), eps 1e-8, L2 decay folded into the gradient, bias-corrected moments, no amsgrad); parameters whose
``.grad`` is None are skipped exactly like the stock optimizer does (the big cfg leaves ``neck.bifpn.0.p5_to_p6.*`` unused).
It is a ``torch.optim.Optimizer``: ``zero_grad`` / ``param_groups`` / LR schedulers / ``state_dict`` keep working, so it drops
into train.py:147 in place of ``torch.optim.Adam``.

CUDA-graph safe: the learning rate and the step count live in a two-float device tensor per group (the kernel derives the
bias corrections from it and a one-thread kernel advances the count), and the table of (param, grad, m, v) pointers is only
re-uploaded when a pointer changed -- inside a captured step the gradients sit at fixed addresses.
"""
import torch

from . import _native as nv


class FusedAdam(torch.optim.Optimizer):
    CHUNK = 16384  # elements per CTA

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale))
        self._tables = {}

    def _table(self, gi, params):
        key = tuple(id(p) for p in params)
        t = self._tables.get(gi)
        if t is not None and t["key"] == key:
            return t
        dev = params[0].device
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        n = len(params)
        host = [torch.empty((n, 5), dtype=torch.int64).pin_memory() for _ in range(3)]
        for h in host:
            a = h.numpy()
            a[:, 0] = [p.data_ptr() for p in params]
            a[:, 1] = 0
            a[:, 2] = [self.state[p]["exp_avg"].data_ptr() for p in params]
            a[:, 3] = [self.state[p]["exp_avg_sq"].data_ptr() for p in params]
            a[:, 4] = [p.numel() for p in params]
        ct, ci = [], []
        for i, p in enumerate(params):
            k = (p.numel() + self.CHUNK - 1) // self.CHUNK
            ct += [i] * k
            ci += list(range(k))
        step0 = float(self.state[params[0]]["step"])
        t = dict(key=key, host=host, events=[None, None, None], turn=0, grads=None, dev_table=torch.empty((n, 5), dtype=torch.int64, device=dev),
                 chunk_tensor=torch.tensor(ct, dtype=torch.int32, device=dev), chunk_index=torch.tensor(ci, dtype=torch.int32, device=dev), n_chunks=len(ct),
                 dyn=torch.tensor([0.0, step0], dtype=torch.float32, device=dev), lr_on_device=None)
        self._tables[gi] = t
        return t

    def sync_hyper(self):
        """Push the groups' current learning rates to the device (call before replaying a CUDA graph that captured step())."""
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is not None and t["lr_on_device"] != float(group["lr"]):
                t["dyn"][0:1].fill_(float(group["lr"]))
                t["lr_on_device"] = float(group["lr"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdam: contiguous fp32 CUDA parameters only (no CPU fallback)")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
            t = self._table(gi, params)
            dev = params[0].device
            with torch.cuda.device(dev):
                grads = [p.grad.data_ptr() for p in params]
                if grads != t["grads"]:  # autograd allocates fresh gradient tensors every eager step; a captured step does not
                    k = t["turn"] = (t["turn"] + 1) % 3
                    if t["events"][k] is not None and not capturing:
                        t["events"][k].synchronize()  # the upload that last used this pinned buffer has completed
                    host = t["host"][k]
                    host.numpy()[:, 1] = grads
                    t["dev_table"].copy_(host, non_blocking=True)
                    if not capturing:
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream(dev))
                        t["events"][k] = ev
                    t["grads"] = grads
                if not capturing:
                    self.sync_hyper()
                    if t["lr_on_device"] is None:
                        t["dyn"][0:1].fill_(float(group["lr"]))
                        t["lr_on_device"] = float(group["lr"])
                step = self.state[params[0]]["step"] + 1
                for p in params:
                    self.state[p]["step"] = step
                b1, b2 = group["betas"]
                nv.check(nv.lib.hn_adam_step(t["dev_table"].data_ptr(), t["chunk_tensor"].data_ptr(), t["chunk_index"].data_ptr(), t["n_chunks"], self.CHUNK,
                                             float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), int(step),
                                             float(group["grad_scale"]), t["dyn"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return loss

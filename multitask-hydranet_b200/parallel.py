"""Data-parallel training: the gradient exchange of the training step (reference:
model/train.py:129-137 wraps the model in DistributedDataParallel(find_unused_parameters=True)).

One process per GPU, replicated weights, local (non-synchronised) BatchNorm as in the reference, and ONE exchange per step:
the mean of the parameter gradients over the ranks.  ``GradAllReduce`` does what DDP's reducer does, without its copies:

  * parameters are bucketed (~25 MB) in REVERSE registration order -- the order in which backward produces gradients;
  * a post-accumulate-grad hook per parameter counts a bucket's gradients in; when the bucket is complete its tensors are
    all-reduced as ONE coalesced NCCL group on a side stream, in place, no flat copy (NVLink 5 / NVSwitch: the cost is
    launch latency and overlap, not link count), while backward keeps running on the compute stream;
  * ``finish()`` flushes the buckets that never completed -- the big cfg leaves ``neck.bifpn.0.p5_to_p6.*`` without gradients
    (bifpn.py:158-165; what DDP needs ``find_unused_parameters`` for) -- and makes the compute stream wait for the side stream.

Backends without coalesced in-place collectives (gloo, used by the CPU tests) take a flatten -> all_reduce -> unflatten path.
"""
import torch
import torch.distributed as dist


def shard_batch(global_batch, world_size, rank):
    """Contiguous [begin, end) image slice of ``rank`` (sizes differ by at most one)."""
    from .sharding import shard_bounds
    return shard_bounds(global_batch, world_size, rank)


class GradAllReduce:
    def __init__(self, params, bucket_mb=25.0, process_group=None, model=None):
        # the hooks read a gradient the moment autograd accumulates it: weight-gradient kernels must then stay on the main
        # stream (train.py puts them on a side stream otherwise)
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_mb * 1024 * 1024:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.works = []
        self.comm_stream = None
        self.exposed_ms = None
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params] if self.world > 1 else []

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    # -- hooks ------------------------------------------------------------------------------------
    def _hook(self, p):
        bi = self.bucket_of[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0 and not self.launched[bi]:
            self._launch(bi)

    def _launch(self, bi):
        self.launched[bi] = True
        st = getattr(self.model, "_train_state", None) if self.model is not None else None
        if st is not None and st.side_dirty:  # gradients produced on the training state's side stream
            from .train import join_side_stream
            join_side_stream(st, st.device)
        grads = [p.grad for p in self.buckets[bi] if p.grad is not None]
        if not grads:
            return
        if grads[0].is_cuda:
            dev = grads[0].device
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(dev)
            self.comm_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.comm_stream):
                for g in grads:
                    g.record_stream(self.comm_stream)
                with dist._coalescing_manager(group=self.group, device=dev, async_ops=True) as cm:
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                self.works.append(cm)
        else:
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)

    def finish(self, measure=False):
        """Call after ``loss.backward()``: flush incomplete buckets, wait for the exchange, reset for the next step."""
        if self.world > 1:
            for bi in range(len(self.buckets)):
                if not self.launched[bi]:
                    self._launch(bi)
            if self.comm_stream is not None:
                cur = torch.cuda.current_stream(self.comm_stream.device)
                if measure:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(cur)
                for w in self.works:
                    w.wait()
                cur.wait_stream(self.comm_stream)
                if measure:  # time the compute stream sat waiting for the exchange = the exposed (non-overlapped) part
                    e1.record(cur)
                    e1.synchronize()
                    self.exposed_ms = e0.elapsed_time(e1)
        self.works = []
        self.pending = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

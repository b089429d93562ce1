"""Data-parallel plumbing: same entry points as /root/reference/distributed.py (init_distributed :28-44,
apply_gradient_allreduce :81-133, reduce_tensor :22-26), redesigned for one NVSwitch box.

Reference behaviour: after backward finishes, flatten ALL grads (one torch.cat), one 233 MB all-reduce, divide,
copy back -- zero overlap and two extra passes over the gradients.  Here:
  * gradients live permanently inside a few flat fp32 buckets (``p.grad`` is a view), so there is no flatten /
    unflatten traffic at all;
  * buckets follow the order gradients become final (last flow first, encoder/embeddings last); a bucket's
    all-reduce (NCCL ``AVG``: the 1/N scaling happens inside the collective) is launched from a
    post-accumulate hook on a side stream as soon as its last gradient lands, so it overlaps the BPTT of the
    earlier flows; the autograd engine's final callback joins the streams;
  * utterances are sharded across ranks (one process per GPU); there is no other collective on the path.
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist
from torch.autograd import Variable


def flat_offsets(tensors, align: int = 64):
    """(offsets, total): tensors packed back to back, each start rounded up to `align` elements (== radam.flat_offsets;
    duplicated here so this module keeps working without the CUDA library)."""
    offs, off = [], 0
    for t in tensors:
        offs.append(off)
        off += (t.numel() + align - 1) // align * align
    return offs, off


def reduce_tensor(tensor, num_gpus):
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= num_gpus
    return rt


def init_distributed(rank, num_gpus, dist_backend="nccl", dist_url=None):
    """MASTER_ADDR/MASTER_PORT rendezvous like the reference; backend honoured (gloo for CPU tests)."""
    if dist_backend == "nccl":
        assert torch.cuda.is_available(), "Distributed mode requires CUDA."
        torch.cuda.set_device(rank % torch.cuda.device_count())
    master_ip = os.getenv('MASTER_ADDR', '127.0.0.1')
    master_port = os.getenv('MASTER_PORT', '6000')
    init_method = dist_url or ('tcp://' + master_ip + ':' + master_port)
    if not dist.is_initialized():
        kw = {}
        if dist_backend == "nccl":
            kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend=dist_backend, world_size=num_gpus, rank=rank, init_method=init_method, **kw)


def _default_buckets(module) -> List[List[torch.nn.Parameter]]:
    """Order in which gradients become final during backward: flows in reverse index order, then everything else."""
    flows = getattr(module, "flows", None)
    seen, buckets = set(), []
    if flows is not None:
        for flow in reversed(list(flows)):
            ps = [p for p in flow.parameters() if p.requires_grad]
            if ps:
                buckets.append(ps)
                seen.update(id(p) for p in ps)
    rest = [p for p in module.parameters() if p.requires_grad and id(p) not in seen]
    if rest:
        buckets.append(rest)
    return buckets


class _Bucket:
    def __init__(self, params):
        self.params = params
        dev = params[0].device
        # same 256-byte-aligned packing as the fused optimizer's flat parameter buffer (radam.flat_offsets), so a bucket
        # is one contiguous run for it; the padding stays zero (AVG of zeros)
        self.offsets, n = flat_offsets(params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(params, self.offsets):
            old = p.grad
            p.grad = self.flat[off: off + p.numel()].view_as(p)     # autograd accumulates in place into the bucket
            if old is not None:
                p.grad.copy_(old)
        self.ready = 0
        self.work = None


def apply_gradient_allreduce(module):
    """Modifies ``module`` in place (class unchanged: no ``.module`` prefix in state_dict keys, like the reference).
    After this call keep gradients allocated: use ``module.zero_grad_buckets()`` (or ``zero_grad(set_to_none=False)``)."""
    world = dist.get_world_size()
    for p in module.state_dict().values():                           # rank-0 broadcast (distributed.py:91-94)
        if torch.is_tensor(p):
            dist.broadcast(p, 0)
    buckets = [_Bucket(ps) for ps in _default_buckets(module)]
    use_cuda = buckets and buckets[0].flat.is_cuda
    comm_stream = torch.cuda.Stream() if use_cuda else None
    avg_ok = dist.get_backend() == "nccl"
    state = {"callback_queued": False}

    def reset(*_):
        """Forward pre-hook (the reference re-arms `needs_reduction` in a forward hook, distributed.py:126-131): a backward
        that raised never ran finalize(), which would leave callback_queued set (every later backward would then skip
        the join and the optimizer would read un-averaged gradients) and stale ready counts / work handles."""
        state["callback_queued"] = False
        for b in buckets:
            if b.work is not None:
                try:
                    b.work.wait()
                except Exception:
                    pass
                b.work = None
            b.ready = 0

    def finalize():
        state["callback_queued"] = False
        for b in buckets:
            # every bucket is reduced on every rank every step, whether or not all (or any) of its parameters received a
            # gradient on THIS rank (untouched gradients are the zeros of zero_grad_buckets): the set and order of
            # collectives is then identical everywhere, so ranks cannot issue mismatched all-reduces and hang
            if b.work is None:
                launch(b)
        for b in buckets:
            if b.work is not None:
                b.work.wait()
                if not avg_ok:
                    b.flat.div_(world)
                b.work = None
            b.ready = 0
        if use_cuda:
            torch.cuda.current_stream().wait_stream(comm_stream)

    def launch(b):
        if use_cuda:
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG if avg_ok else dist.ReduceOp.SUM, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, async_op=True)

    def make_hook(b):
        def hook(param):
            if not state["callback_queued"]:
                state["callback_queued"] = True
                Variable._execution_engine.queue_callback(finalize)
            if param.grad.data_ptr() < b.flat.data_ptr() or param.grad.data_ptr() >= b.flat.data_ptr() + b.flat.numel() * 4:
                # someone replaced .grad (e.g. zero_grad(set_to_none=True)): fold it back into the bucket
                for q, off in zip(b.params, b.offsets):
                    if q is param:
                        b.flat[off: off + q.numel()].view_as(q).copy_(param.grad)
                        param.grad = b.flat[off: off + q.numel()].view_as(q)
            b.ready += 1
            if b.ready == len(b.params):
                launch(b)
        return hook

    for b in buckets:
        for p in b.params:
            p.register_post_accumulate_grad_hook(make_hook(b))
    module.register_forward_pre_hook(reset)

    def zero_grad_buckets():
        for b in buckets:
            b.flat.zero_()
    module.zero_grad_buckets = zero_grad_buckets
    module._grad_buckets = buckets
    return module

"""RAdam with the reference's constructor, update rule and state layout (/root/reference/radam.py:25-122), executed as
fused CUDA launches over flat buffers (csrc/optim.cu) instead of a per-tensor Python loop.

Differences a user of the reference sees:
  * parameters must already live on the GPU when the optimizer is built (``model.cuda()`` first, as train.py:226
    does); they are re-homed into one flat fp32 buffer per param group (``p.data`` becomes a view, values unchanged);
  * ``exp_avg`` / ``exp_avg_sq`` in ``state`` are views into flat buffers (``state_dict()`` has the reference's layout,
    so checkpoints interchange, train.py:110-139);
  * gradients are kept allocated inside a flat buffer: use ``optimizer.zero_grad()`` (one memset).  Gradients that
    live elsewhere (e.g. the all-reduce buckets of ``distributed.apply_gradient_allreduce``, or fresh tensors after
    ``model.zero_grad()``) are handled too -- one launch per contiguous run instead of one for everything;
  * ``optimizer.clip_grad_norm_(max_norm)`` replaces ``torch.nn.utils.clip_grad_norm_`` (train.py:326): the norm and
    the clip coefficient stay on the device (no host sync) and the coefficient is applied inside the next ``step()``
    instead of rewriting every gradient.

There is no CPU path: construction fails loudly without CUDA tensors / the native library.
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch.optim.optimizer import Optimizer

from . import _lib
from ._lib import FlowtronB200Error

FLAT_ALIGN = 64          # floats: every tensor starts 256-byte aligned inside a flat buffer


def flat_offsets(tensors, align: int = FLAT_ALIGN):
    """(offsets, total) for packing ``tensors`` back to back with each start rounded up to ``align`` elements.
    Shared with distributed._Bucket so optimizer segments and all-reduce buckets agree on the layout."""
    offs, off = [], 0
    for t in tensors:
        offs.append(off)
        off += (t.numel() + align - 1) // align * align
    return offs, off


def radam_scalars(step: int, lr: float, beta1: float, beta2: float):
    """(N_sma, step_size): radam.py:87-107 (the reference caches these per step % 10; same values)."""
    beta2_t = beta2 ** step
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma >= 5:
        step_size = lr * math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma
                                   * n_sma_max / (n_sma_max - 2)) / (1 - beta1 ** step)
    else:
        step_size = lr / (1 - beta1 ** step)
    return n_sma, step_size


class _Flat:
    """Flat storage of one param group."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        dev = params[0].device
        self.offsets, self.total = flat_offsets(params)
        self.p = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.g = None                                   # own flat gradient buffer (allocated on demand)
        for p, off in zip(params, self.offsets):
            view = self.p[off: off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.seg_key, self.segs = None, None

    def view(self, buf, i):
        p, off = self.params[i], self.offsets[i]
        return buf[off: off + p.numel()].view_as(p)

    def install_grads(self):
        """Point the .grad of every TRAINABLE parameter at the own flat buffer (keeps values of existing gradients).
        Frozen parameters (requires_grad=False, train.py's finetune_layers) keep .grad = None, so step() skips them like
        the reference does (radam.py:53-54): no weight decay, no step count."""
        if self.g is None:
            self.g = torch.zeros(self.total, dtype=torch.float32, device=self.p.device)
        for i, p in enumerate(self.params):
            if not p.requires_grad:
                continue
            gv = self.view(self.g, i)
            if p.grad is not None and p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
            p.grad = gv

    def owns_grads(self) -> bool:
        if self.g is None:
            return False
        base = self.g.data_ptr()
        return all((not p.requires_grad and p.grad is None) or (p.grad is not None and p.grad.data_ptr() == base + 4 * off)
                   for p, off in zip(self.params, self.offsets))

    def rehome(self):
        """Re-attach parameters whose storage was moved out of the flat buffer after construction -- cuDNN's
        ``nn.LSTM.flatten_parameters()`` (``param.set_``), ``module.to()`` / ``.half()`` -- which would otherwise make
        step() update a dead copy while the live weights silently stop training.  The live values win."""
        base, moved = self.p.data_ptr(), 0
        for p, off in zip(self.params, self.offsets):
            if p.data_ptr() != base + 4 * off:
                if p.dtype != torch.float32 or p.device != self.p.device:
                    raise FlowtronB200Error("flowtron_b200.RAdam: a parameter changed dtype/device after the optimizer was "
                                            "built; rebuild the optimizer")
                view = self.p[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                moved += 1
        return moved


class RAdam(Optimizer):
    """RAdam optimizer (fused, flat-buffer).  Signature: radam.py:28-29."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, capturable=False):
        """``capturable=True`` (extension, like torch.optim's): the step count lives on the device and the kernels derive
        N_sma / step_size from it, so ``step()`` can be captured in a CUDA graph and replayed; requires that every
        trainable parameter of a group receives a gradient every step (one step count per group)."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.capturable = bool(capturable)
        self._flats = []
        for group in self.param_groups:
            ps = list(group['params'])
            for p in ps:
                self._require_cuda(p)
                if p.grad is not None and p.grad.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')      # radam.py:57-60
            _lib.lib()                                   # fail loudly now if the native library is missing
            flat = _Flat(ps)
            if all(p.grad is None for p in ps) and any(p.requires_grad for p in ps):
                flat.install_grads()
            self._flats.append(flat)
            for i, p in enumerate(ps):
                self.state[p] = {'step': 0, 'exp_avg': flat.view(flat.m, i), 'exp_avg_sq': flat.view(flat.v, i)}
        self._partials = None
        self._norm_coef = None
        self._pending_coef = None
        self._step_dev = [torch.zeros(1, dtype=torch.int32, device=f.p.device) for f in self._flats] if self.capturable else None
        self._built = True

    def add_param_group(self, param_group):
        """Groups are fixed at construction (each owns flat buffers); adding one later would be silently ignored by
        step(), so it is refused."""
        if getattr(self, "_built", False):
            raise FlowtronB200Error("flowtron_b200.RAdam: add_param_group after construction is not supported; "
                                    "build a new optimizer with all groups")
        super().add_param_group(param_group)

    @staticmethod
    def _require_cuda(p):
        if not p.is_cuda or p.dtype != torch.float32:
            raise FlowtronB200Error("flowtron_b200.RAdam needs fp32 CUDA parameters (move the model to the GPU first); "
                                    "there is no CPU path")

    # ------------------------------------------------------------------------------------------ gradient runs
    def _segments(self, flat: _Flat):
        """Maximal runs of parameters whose values (flat.p) AND gradients are contiguous in memory and share a step
        count: [(element offset into flat.p/m/v, gradient device pointer, n elements, step)].  Inside a run the
        alignment padding between tensors is part of the run (zeros on both sides: own flat buffer or an all-reduce
        bucket laid out with the same flat_offsets); two gradients merge only if they are views of one base tensor or
        there is no padding between them, and the trailing padding of a run is never touched.  Cached on the tuple of
        (gradient pointer, step)."""
        # the parameter pointer is part of the key: a parameter that left the flat buffer invalidates the cache and is
        # re-homed before any kernel sees a stale pointer
        key = tuple((p.grad.data_ptr() if p.grad is not None else 0, self.state[p]['step'], p.data_ptr()) for p in flat.params)
        if key == flat.seg_key:
            return flat.segs
        if flat.rehome():
            key = tuple((p.grad.data_ptr() if p.grad is not None else 0, self.state[p]['step'], p.data_ptr()) for p in flat.params)
        if flat.g is not None:                            # own gradient buffer no longer referenced (buckets took over): free it
            lo, hi = flat.g.data_ptr(), flat.g.data_ptr() + 4 * flat.total
            if not any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in flat.params):
                flat.g = None
        runs, run = [], None
        n_par = len(flat.params)
        for i, p in enumerate(flat.params):
            g = p.grad
            if g is None:                                 # radam.py:53-54: parameters without a gradient are skipped
                run = None
                continue
            if g.is_sparse:
                raise RuntimeError('RAdam does not support sparse gradients')
            if g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                raise FlowtronB200Error("flowtron_b200.RAdam: gradients must be contiguous fp32 tensors on the parameter's device")
            off, st = flat.offsets[i], self.state[p]['step']
            span = (flat.offsets[i + 1] if i + 1 < n_par else flat.total) - off      # numel + alignment padding
            base = g._base if g._base is not None else g
            if (run is not None and run['step'] == st and run['off'] + run['len'] == off
                    and run['gptr'] + 4 * run['len'] == g.data_ptr() and (run['pad'] == 0 or run['base'] == base.data_ptr())):
                run['len'] += span
            else:
                run = {'off': off, 'gptr': g.data_ptr(), 'len': span, 'step': st, 'base': base.data_ptr()}
                runs.append(run)
            run['pad'] = span - p.numel()
        segs = [(r['off'], r['gptr'], r['len'] - r['pad'], r['step']) for r in runs]
        flat.seg_key, flat.segs = key, segs
        return segs

    # ------------------------------------------------------------------------------------------ public API
    def zero_grad(self, set_to_none: bool = False):
        """Default keeps gradients allocated in the flat buffer (one memset).  ``set_to_none=True`` behaves like
        torch's and drops them (the next backward then allocates per-tensor gradients: slower multi-launch path)."""
        if set_to_none:
            for flat in self._flats:
                for p in flat.params:
                    p.grad = None
            return
        for flat in self._flats:
            if flat.owns_grads():
                flat.g.zero_()
            elif all(p.grad is None for p in flat.params):
                flat.install_grads()
                flat.g.zero_()
            else:
                for p in flat.params:
                    if p.grad is not None:
                        p.grad.zero_()

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float):
        """torch.nn.utils.clip_grad_norm_(parameters, max_norm) (train.py:326) without touching the gradients: returns the
        total norm (0-dim device tensor, no sync); the coefficient min(1, max_norm/(norm+1e-6)) multiplies the gradients
        inside the next step()."""
        runs = [(gptr, n) for flat in self._flats for (_, gptr, n, _) in self._segments(flat)]
        if not runs:
            return torch.zeros((), device=self._flats[0].p.device)
        dev = self._flats[0].p.device
        need = _lib.SUMSQ_PARTIALS * len(runs)
        if self._partials is None or self._partials.numel() < need:
            self._partials = torch.empty(need, dtype=torch.float32, device=dev)
        if self._norm_coef is None:
            self._norm_coef = torch.empty(2, dtype=torch.float32, device=dev)
        for k, (gptr, n) in enumerate(runs):
            _lib.sumsq_partials_raw(gptr, n, self._partials.data_ptr() + 4 * _lib.SUMSQ_PARTIALS * k)
        _lib.clip_coef(self._partials, need, max_norm, self._norm_coef)
        self._pending_coef = self._norm_coef
        return self._norm_coef[0]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        coef_ptr = self._pending_coef.data_ptr() + 4 if self._pending_coef is not None else 0
        if self.capturable:
            for gi, (group, flat) in enumerate(zip(self.param_groups, self._flats)):
                segs = self._segments(flat)
                if not segs:
                    continue
                if len({st for _, _, _, st in segs}) != 1 or any(p.grad is None for p in flat.params if p.requires_grad):
                    raise FlowtronB200Error("flowtron_b200.RAdam(capturable=True) needs one step count per group: every "
                                            "trainable parameter must receive a gradient every step")
                beta1, beta2 = group['betas']
                for off, gptr, n, _ in segs:
                    _lib.radam_step_dev_raw(flat.p.data_ptr() + 4 * off, gptr, flat.m.data_ptr() + 4 * off,
                                            flat.v.data_ptr() + 4 * off, n, beta1, beta2, group['eps'], group['weight_decay'],
                                            group['lr'], self._step_dev[gi], coef_ptr)
                _lib.step_increment(self._step_dev[gi])
            self._pending_coef = None
            return loss
        for group, flat in zip(self.param_groups, self._flats):
            segs = self._segments(flat)
            if not segs:
                continue
            beta1, beta2 = group['betas']
            for off, gptr, n, st in segs:
                n_sma, step_size = radam_scalars(st + 1, group['lr'], beta1, beta2)
                _lib.radam_step_raw(flat.p.data_ptr() + 4 * off, gptr, flat.m.data_ptr() + 4 * off, flat.v.data_ptr() + 4 * off,
                                    n, beta1, beta2, group['eps'], group['weight_decay'] * group['lr'], step_size, n_sma >= 5,
                                    coef_ptr)
            for p in flat.params:
                if p.grad is not None:
                    self.state[p]['step'] += 1
            if flat.seg_key is not None:                 # keep the cache valid across the uniform step increment
                flat.seg_key = tuple((gp, s + 1 if gp else s, pp) for gp, s, pp in flat.seg_key)
                flat.segs = [(o, gp, n, s + 1) for o, gp, n, s in flat.segs]
        self._pending_coef = None
        return loss

    def _sync_steps(self):
        """capturable: bring the Python-side ``state[p]['step']`` in line with the device counters (one host sync; only
        when the state is inspected or saved -- graph replays advance the device counter without running Python)."""
        if not self.capturable:
            return
        for sd, flat in zip(self._step_dev, self._flats):
            n = int(sd.item())
            for p in flat.params:
                if p.requires_grad:
                    self.state[p]['step'] = n
            flat.seg_key = None

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Accepts the reference's optimizer checkpoints (train.py:123): moments are copied into the flat buffers."""
        super().load_state_dict(state_dict)
        for flat in self._flats:
            for i, p in enumerate(flat.params):
                st = self.state.get(p)
                if not st:
                    self.state[p] = {'step': 0, 'exp_avg': flat.view(flat.m, i), 'exp_avg_sq': flat.view(flat.v, i)}
                    continue
                mv, vv = flat.view(flat.m, i), flat.view(flat.v, i)
                mv.copy_(st['exp_avg'])
                vv.copy_(st['exp_avg_sq'])
                st['exp_avg'], st['exp_avg_sq'] = mv, vv
                st['step'] = int(st['step'])
            flat.seg_key = None
        if self.capturable:
            for sd, flat in zip(self._step_dev, self._flats):
                steps = {self.state[p]['step'] for p in flat.params if p.requires_grad}
                sd.fill_(max(steps) if steps else 0)

"""Data parallelism for adapter training: one flat gradient arena, bucketed all-reduce overlapped with backward.

The reference has no collective; kohya users get DDP from accelerate.  Here the batch
is sharded across ranks (one process per GPU), base weights are frozen replicas that never travel,
and only the adapter gradients — 1.6 M … 185 M elements for the SDXL configs — are averaged:

* every trainable parameter's ``.grad`` is a view into ONE contiguous buffer per dtype, laid out in
  REVERSE parameter order (the order backward produces them), cut into buckets of ``bucket_bytes``;
* ``overlap="backward"`` (default on CUDA): a post-accumulate-grad hook counts a bucket's gradients
  in; when the last one lands the bucket's ``ncclAllReduce`` is issued on a side stream forked from
  the backward stream at that point, so only the LAST bucket (the first layers' gradients) is exposed
  after backward.  The fork/join are plain stream waits, so the whole step — collectives included —
  is CUDA-graph capturable (NCCL kernels are captured like any other);
* ``overlap=True`` / ``"step"``: one all-reduce per dtype issued after backward on the side stream
  (round-1 behaviour); ``overlap=False``: same on the compute stream (CPU / gloo tests);
* averaging uses NCCL's AVG reduction (no extra pass); gloo (CPU tests) falls back to SUM + scale.

Usage contract: clear gradients with ``dp.zero_grad()`` (keeps every ``.grad`` aliased to the arena).
``optimizer.zero_grad(set_to_none=True)`` — what kohya / accelerate call — detaches the aliases;
``allreduce()`` detects that and re-packs the fresh gradients into the arena (one extra copy per
parameter, and no overlap for that step) instead of silently reducing a stale buffer.  Parameters are
broadcast from rank 0 at construction, so adapter initialisation need not be seeded identically by hand.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("arena", "lo", "hi", "params", "pending", "fired")

    def __init__(self, arena, lo, hi, params):
        self.arena, self.lo, self.hi, self.params = arena, lo, hi, params
        self.pending = len(params)
        self.fired = False

    @property
    def view(self):
        return self.arena[self.lo : self.hi]


class FlatGradAllReduce:
    def __init__(self, params, process_group=None, bucket_dtype=None, overlap=True, broadcast_params=True,
                 bucket_bytes=32 << 20):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradAllReduce: no trainable parameters")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if broadcast_params and self.world > 1:
            self._broadcast_params()
        dev = self.params[0].device
        on_cuda = dev.type == "cuda"
        if overlap is True:
            overlap = "backward" if bucket_dtype is None else "step"
        self.mode = overlap if overlap else "none"  # "backward" | "step" | "none"
        if self.mode == "backward" and bucket_dtype is not None:
            raise ValueError("FlatGradAllReduce: overlap='backward' needs the arena in the parameter dtype")

        self.arenas = {}
        self.buckets = []
        self._bucket_of = {}
        by_dtype = {}
        for p in self.params:
            by_dtype.setdefault((p.dtype, p.device), []).append(p)
        for (dtype, device), plist in by_dtype.items():
            plist = plist[::-1]  # backward order: the last layer's gradient is ready first
            total = sum(p.numel() for p in plist)
            arena = torch.zeros(total, dtype=bucket_dtype or dtype, device=device)
            off = 0
            cur, cur_lo = [], 0
            limit = max(1, bucket_bytes // arena.element_size())
            for p in plist:
                n = p.numel()
                if arena.dtype == p.dtype:
                    p.grad = arena[off : off + n].view_as(p)  # autograd accumulates in place into the arena
                cur.append(p)
                off += n
                if off - cur_lo >= limit:
                    self._add_bucket(arena, cur_lo, off, cur)
                    cur, cur_lo = [], off
            if cur:
                self._add_bucket(arena, cur_lo, off, cur)
            self.arenas[(dtype, device)] = (arena, plist)
        self._stream = torch.cuda.Stream(device=dev) if (self.mode != "none" and on_cuda) else None
        self._compute_stream = None
        self._pending = False
        self._hooks = []
        self._armed = False
        self.realiased = 0  # parameters whose .grad had to be re-packed (diagnostic: 0 with dp.zero_grad())
        self.buckets_overlapped = 0  # buckets issued from inside backward during the last step
        if self.mode == "backward" and self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self._armed = True

    # ------------------------------------------------------------------ layout
    def _add_bucket(self, arena, lo, hi, params):
        b = _Bucket(arena, lo, hi, list(params))
        self.buckets.append(b)
        for p in params:
            self._bucket_of[p] = b

    def _broadcast_params(self):
        """Rank 0's adapter parameters become everyone's (one flat broadcast per dtype)."""
        by_dtype = {}
        for p in self.params:
            by_dtype.setdefault((p.dtype, p.device), []).append(p)
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        with torch.no_grad():
            for plist in by_dtype.values():
                flat = torch.cat([p.detach().reshape(-1) for p in plist])
                dist.broadcast(flat, src=src, group=self.group)
                off = 0
                for p in plist:
                    n = p.numel()
                    p.copy_(flat[off : off + n].view_as(p))
                    off += n

    def _realias(self, arena, plist):
        """Re-establish ``p.grad is a view of the arena``.  A gradient autograd allocated elsewhere (after
        ``zero_grad(set_to_none=True)``) is copied in; a missing one contributes zeros.  Returns how many
        parameters had lost their alias."""
        off, lost = 0, 0
        esz = arena.element_size()
        base = arena.data_ptr()
        for p in plist:
            n = p.numel()
            g = p.grad
            if g is None or g.data_ptr() != base + off * esz or g.dtype != arena.dtype:
                lost += 1
                view = arena[off : off + n].view_as(p)
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
                p.grad = view
            off += n
        return lost

    @property
    def num_elements(self):
        return sum(a.numel() for a, _ in self.arenas.values())

    def zero_grad(self):
        for arena, _ in self.arenas.values():
            arena.zero_()
        self._rearm()

    def _rearm(self):
        for b in self.buckets:
            b.pending = len(b.params)
            b.fired = False
        self.buckets_overlapped = 0

    # ---------------------------------------------------------------- reduce
    def _reduce(self, t):
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def _on_grad(self, p):
        """post-accumulate-grad hook (autograd worker thread, backward's stream is current)."""
        b = self._bucket_of.get(p)
        if b is None or b.fired or not self._armed:
            return
        g = p.grad
        if g is None or g.data_ptr() < b.arena.data_ptr() or g.data_ptr() >= b.arena.data_ptr() + b.arena.numel() * b.arena.element_size():
            # the alias was lost (zero_grad(set_to_none=True)): allreduce() will re-pack and reduce at the end
            b.pending = -1
            return
        b.pending -= 1
        if b.pending == 0:
            self._issue(b)
            self.buckets_overlapped += 1

    def _issue(self, b):
        if self._stream is None:  # CPU / gloo: synchronous
            self._reduce(b.view)
            b.fired = True
            return
        cur = torch.cuda.current_stream()
        self._compute_stream = cur
        self._stream.wait_stream(cur)  # fork: everything that produced this bucket's gradients
        with torch.cuda.stream(self._stream):
            self._reduce(b.view)
        b.fired = True
        self._pending = True

    def _gather_foreign(self, arena, plist):
        """bucket_dtype != param dtype: pack grads into the arena before the reduce."""
        off = 0
        for p in plist:
            n = p.numel()
            if p.grad is not None:
                arena[off : off + n].copy_(p.grad.reshape(-1))
            else:
                arena[off : off + n].zero_()
            off += n

    def _scatter_foreign(self, arena, plist):
        off = 0
        for p in plist:
            n = p.numel()
            g = arena[off : off + n].view_as(p).to(p.dtype)
            if self._stream is not None:
                # allocated on the side stream, consumed on the compute stream after wait()
                g.record_stream(self._compute_stream)
            p.grad = g
            off += n

    def allreduce(self):
        """Average the adapter gradients over the data-parallel group (call after backward).  In
        ``overlap="backward"`` mode most buckets are already in flight; this issues whatever is left."""
        if self.world == 1:
            return
        cur = torch.cuda.current_stream() if self._stream is not None else None
        self._compute_stream = cur
        lost = 0
        for (dtype, _), (arena, plist) in self.arenas.items():
            if arena.dtype == dtype:
                lost += self._realias(arena, plist)  # on the compute stream, before the fork
        self.realiased += lost
        if self.mode == "backward" and lost == 0:
            for b in self.buckets:
                if not b.fired:
                    self._issue(b)
            return
        if self.mode == "backward" and self._pending:
            # some buckets were reduced from inside backward before the alias loss was seen for others:
            # finish them, then reduce only the unfired buckets (their re-packed content)
            cur.wait_stream(self._stream)
            for b in self.buckets:
                if not b.fired:
                    self._issue(b)
            return
        if self._stream is not None:
            self._stream.wait_stream(cur)
        ctx = torch.cuda.stream(self._stream) if self._stream is not None else _NullCtx()
        with ctx:
            for (dtype, _), (arena, plist) in self.arenas.items():
                foreign = arena.dtype != dtype
                if foreign:
                    self._gather_foreign(arena, plist)
                self._reduce(arena)
                if foreign:
                    self._scatter_foreign(arena, plist)
        for b in self.buckets:
            b.fired = True
        self._pending = self._stream is not None

    def wait(self):
        """Make the compute stream wait for the reduced gradients (before optimizer.step); re-arms the
        bucket counters for the next backward."""
        if self._pending:
            torch.cuda.current_stream().wait_stream(self._stream)
            self._pending = False
        n = self.buckets_overlapped
        self._rearm()
        self.buckets_overlapped = n

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self._armed = False


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

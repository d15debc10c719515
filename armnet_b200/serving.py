"""Pipelined scoring of HOST batches with a drop-in ARMNetModel (eval mode): what a serving process calls.

The reference scores a batch with `y = model(batch)` after `.cuda(non_blocking=True)` copies and then syncs on
`.item()` / `.cpu()` every batch (train.py:104-122).  BatchScorer keeps the same inputs and outputs
(pinned host `id` [B,F] int64 / `value` [B,F] float32 in, host `y` [B] out) but overlaps the three phases of
consecutive batches on the device:

    copy stream        H2D(i+1)            H2D(i+2)            H2D(i+3)
    compute stream 0   forward(i) D2H(i)                       forward(i+2) ...
    compute stream 1            forward(i+1) D2H(i+1)

With more than one compute stream consecutive batches are independent streams of work: the narrow end of batch i's
forward (tcgen05 GEMM and MLP tail: 128 CTAs on 148 SMs, pipeline fill / drain) overlaps the persistent fused kernel of
batch i+1.  Measured on B200 at the Criteo shape: depth 2 / 1 stream 238 us per batch, depth 4 / 2 streams 212 us,
depth 6 / 3 streams 207 us.

Each slot owns static device buffers, so the whole forward of a slot (attention pre-contraction, fused
lookup+interaction kernel, tcgen05 GEMM, MLP tail kernel) is captured once in a CUDA graph and replayed: one graph
launch per batch instead of ~6 kernel launches plus allocator traffic.  Nothing here computes: the arithmetic is the
same kernels ARMNetModel.forward launches.

Unlike ARMNetModel.forward, the caller's host `value` tensor is NOT clamped in place (the clamp of armnet.py:82 is
applied to the device copy).

The captured graphs hold pointers to parameter-derived caches (padded table, pre-contracted attention matrix, split MLP
weights).  submit() compares the parameters' version counters with the ones seen at capture time and re-captures when
they changed (in-place updates such as optimizer steps or load_state_dict bump the counters).
"""
import torch

__all__ = ['BatchScorer']


class _Slot:
    def __init__(self, B, F, dev, ids_dtype):
        self.ids = torch.zeros(B, F, dtype=ids_dtype, device=dev)
        self.vals = torch.ones(B, F, dtype=torch.float32, device=dev)
        self.y_host = None
        self.y_dev = None
        self.flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()   # copy of the model's bad-id flag, per batch
        self.graph = None
        self.ev_in = torch.cuda.Event()
        self.ev_done = torch.cuda.Event()
        self.used = False


class BatchScorer:
    def __init__(self, model, batch_size, nfield, depth=4, use_graph=True, ids_dtype=torch.int64, compute_streams=2):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError('BatchScorer needs the model on a CUDA device (armnet_b200 has no CPU path)')
        if model.training:
            raise RuntimeError('BatchScorer scores in eval mode: call model.eval() first')
        if getattr(model, 'validate_ids', False) is True:
            raise RuntimeError("validate_ids=True synchronises every batch; use the default 'lazy' (checked in result())")
        self.model, self.dev = model, p.device
        self.B, self.F, self.depth, self.use_graph = batch_size, nfield, depth, use_graph
        self.copy_stream = torch.cuda.Stream(self.dev)
        if depth < 1 or compute_streams < 1 or (compute_streams > 1 and depth < 2 * compute_streams):
            raise ValueError('depth >= 2 * compute_streams (a slot is reused only after the other streams made progress)')
        self.compute_streams = [torch.cuda.Stream(self.dev) for _ in range(max(1, compute_streams))]
        self.compute_stream = self.compute_streams[0]
        self.slots = [_Slot(batch_size, nfield, self.dev, ids_dtype) for _ in range(depth)]
        for i, s in enumerate(self.slots):
            s.stream = self.compute_streams[i % len(self.compute_streams)]
        self.n_submitted = 0
        self._tensors = list(model.parameters()) + [b for b in model.buffers() if b.dtype.is_floating_point]
        self._prepare()

    def _forward(self, slot):
        with torch.no_grad():
            return self.model({'id': slot.ids, 'value': slot.vals}).reshape(-1)

    def _version_key(self):
        return tuple(t._version for t in self._tensors)

    def _prepare(self):
        """Warm every lazily built cache (padded table, folded BatchNorm, split weights), then capture one graph per slot."""
        self._key = self._version_key()
        for s in self.slots:
            cs = s.stream
            cs.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(cs):
                for _ in range(2):
                    y = self._forward(s)
                s.y_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            cs.synchronize()
        if not self.use_graph:
            return
        for s in self.slots:
            s.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(s.graph, stream=s.stream):
                s.y_dev = self._forward(s)
            s.stream.synchronize()

    def submit(self, ids_host, values_host):
        """Enqueue one batch (pinned host tensors [B,F]); returns a ticket for result(). At most `depth` tickets may be
        outstanding: result(t) must have been called before submit() reuses its slot."""
        if tuple(ids_host.shape) != (self.B, self.F) or tuple(values_host.shape) != (self.B, self.F):
            raise ValueError(f'BatchScorer was built for batches of shape {(self.B, self.F)}')
        if self.use_graph and self._version_key() != self._key:     # parameters changed since capture: rebuild
            torch.cuda.synchronize(self.dev)
            for sl in self.slots:
                sl.graph = None
            self._prepare()
        t = self.n_submitted
        s = self.slots[t % self.depth]
        if s.used:
            self.copy_stream.wait_event(s.ev_done)       # the slot's previous forward + D2H are done with its buffers
        with torch.cuda.stream(self.copy_stream):
            s.ids.copy_(ids_host, non_blocking=True)
            s.vals.copy_(values_host, non_blocking=True)
            s.ev_in.record(self.copy_stream)
        with torch.cuda.stream(s.stream):
            s.stream.wait_event(s.ev_in)
            if s.graph is not None:
                s.graph.replay()
                y = s.y_dev
            else:
                y = self._forward(s)
            s.y_host.copy_(y, non_blocking=True)
            flag = getattr(self.model, '_err_flag', None)
            if flag is not None and getattr(self.model, 'validate_ids', False):
                s.flag_host.copy_(flag, non_blocking=True)     # 4 bytes behind y: the id check costs no synchronisation
            s.ev_done.record(s.stream)
        s.used = True
        self.n_submitted += 1
        return t

    def result(self, ticket):
        """Host tensor y [B] of a submitted batch (blocks until its D2H copy has landed). The tensor is the slot's pinned
        buffer: consume or clone it before `depth` further batches are submitted."""
        if not (self.n_submitted - self.depth <= ticket < self.n_submitted):
            raise ValueError('ticket is not outstanding')
        s = self.slots[ticket % self.depth]
        s.ev_done.synchronize()
        if int(s.flag_host[0]) & 1:        # the reference raises IndexError for an id outside [0, nfeat) (layers.py:20)
            s.flag_host.zero_()
            self.model._err_flag.zero_()
            raise IndexError('index out of range in self (armnet_b200: id outside [0, nfeat) in a scored batch)')
        return s.y_host

    def score(self, batches):
        """Convenience: iterate over (ids_host, values_host) pairs, yield cloned host results in order."""
        pending = []
        for ids_h, vals_h in batches:
            if len(pending) == self.depth:
                yield self.result(pending.pop(0)).clone()
            pending.append(self.submit(ids_h, vals_h))
        for t in pending:
            yield self.result(t).clone()

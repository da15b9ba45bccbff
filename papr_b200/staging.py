"""Host <-> device staging around a training step or a frame render (the loops of reference train.py:155-179 and
test.py:76-104, which upload one batch of rays, run the model and read the loss / the rendered image back every iteration).

Done naively -- upload, run, ``loss.item()`` -- the GPU idles at the start of every iteration while the host, which has just
been blocked on the read-back, enqueues the first kernels of the next one, and the PCIe copies sit on the compute stream.
``StepPipeline`` keeps the same per-iteration traffic but software-pipelines it:

  * inputs are copied from (pinned) host memory on a copy stream; the compute stream waits for that event only;
  * results are copied into pinned host buffers on the copy stream once the step's event has fired;
  * ``submit(batch i)`` returns the host results of iteration i-1, i.e. the host blocks on iteration i-1 only after
    iteration i has been enqueued.  ``flush()`` returns the last one.

Results are identical to the plain loop (same kernels, same order on the compute stream); only the host-side read-back is
one iteration late, which is what a training loop that logs its loss tolerates and what a frame writer does anyway.
"""
import torch


class StepPipeline:
    def __init__(self, fn, device, depth=2):
        """fn(device_batch: dict) -> tensor or tuple of tensors (all on `device`); `depth` result slots (>= 2)."""
        assert depth >= 2
        self.fn, self.device, self.depth = fn, torch.device(device), depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * depth          # (event, [pinned host tensors], single)
        self.n = 0
        self.h2d_bytes = self.d2h_bytes = 0

    def _upload(self, host_batch):
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            dev = {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        main.wait_event(ready)
        for v in dev.values():
            if isinstance(v, torch.Tensor):
                v.record_stream(main)        # allocated on the copy stream, consumed on the compute stream
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in host_batch.values() if isinstance(v, torch.Tensor))
        return dev

    def submit(self, host_batch):
        """Enqueue one iteration; returns the host results of the PREVIOUS iteration (None for the first call)."""
        main = torch.cuda.current_stream(self.device)
        outs = self.fn(self._upload(host_batch))
        single = isinstance(outs, torch.Tensor)
        outs = [outs] if single else list(outs)
        slot = self.n % self.depth
        old = self.slots[slot]
        bufs = old[1] if old is not None and all(b.shape == o.shape and b.dtype == o.dtype for b, o in zip(old[1], outs)) \
            and len(old[1]) == len(outs) else [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        done_compute = torch.cuda.Event()
        done_compute.record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(done_compute)
            for b, o in zip(bufs, outs):
                b.copy_(o.detach(), non_blocking=True)
                o.record_stream(self.copy_stream)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.d2h_bytes = sum(b.numel() * b.element_size() for b in bufs)
        self.slots[slot] = (done, bufs, single)
        self.n += 1
        return self._collect(self.n - 2) if self.n >= 2 else None

    def _collect(self, i):
        done, bufs, single = self.slots[i % self.depth]
        done.synchronize()
        return bufs[0] if single else tuple(bufs)

    def flush(self):
        """Host results of the last submitted iteration."""
        return self._collect(self.n - 1) if self.n else None


class GraphedCall:
    """A forward-only call with fixed shapes captured once into a CUDA graph and replayed per frame.

    A frame render launches ~30 of the library's kernels and ~250 small torch kernels (the grid build of the selection, the
    query-side elementwise work, the composite); enqueued from Python that is 5-15 ms of host time per call, which is what
    a rank of an N-GPU render (a stripe of 100-150 rows, 3-5 ms of GPU work) would otherwise be bound by.  Replaying the
    captured graph costs one launch.  `fn(*static_inputs)` must be free of host synchronisation and of data-dependent
    host control flow (the render path of papr_b200.model.PAPR is); inputs are copied into the static tensors on replay.
    """

    def __init__(self, fn, example_inputs, warmup=3):
        self.fn = fn
        self.static_in = [t.clone() for t in example_inputs]
        dev = self.static_in[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():          # warm-up off the capture: lazy one-time work (function
            for _ in range(warmup):                             # attributes, allocator pools, weight images) happens here
                self.fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            out = self.fn(*self.static_in)
        self.static_out = out

    def __call__(self, *inputs):
        """Replays the graph on new inputs; returns the STATIC output tensor(s) (overwritten by the next call)."""
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out

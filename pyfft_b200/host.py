"""Host-buffer convenience path: transform data that lives in (pinned) host memory.

The reference's users do ``gpu = toGpu(host); plan.execute(gpu); host = fromGpu(gpu)``
(test/helpers.py:37-53).  ``HostPipeline`` is that sequence, chunked along the batch axis and
spread over a few CUDA streams so the H2D copy of chunk k+1, the FFT of chunk k and the D2H copy
of chunk k-1 overlap (two copy engines + SMs).  All compute still goes through ``Plan.execute``.
"""
import numpy

from .cuda import Plan


class HostPipeline(object):
    def __init__(self, shape, dtype=numpy.complex64, batch=1, chunks=8, slots=3, device=None, **plan_kwargs):
        import torch
        self._torch = torch
        dt = numpy.dtype(dtype)
        if dt.kind != "c":
            raise ValueError("HostPipeline handles interleaved complex data")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
        self.size = int(numpy.prod(self.shape))
        self.batch = int(batch)
        chunks = max(1, min(int(chunks), self.batch))
        while self.batch % chunks:
            chunks -= 1
        self.chunks = chunks
        self.chunk_batch = self.batch // chunks
        self.tdtype = torch.complex64 if dt == numpy.complex64 else torch.complex128
        self.slots = max(1, min(int(slots), chunks))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.slots)]
        self.plans = [Plan(self.shape, dtype=dt, stream=s, **plan_kwargs) for s in self.streams]
        n = self.chunk_batch * self.size
        self.bufs = [torch.empty(n, dtype=self.tdtype, device=self.device) for _ in range(self.slots)]
        self.h2d_bytes = self.batch * self.size * dt.itemsize
        self.d2h_bytes = self.h2d_bytes

    def run(self, host_in, host_out, inverse=False):
        """host_in / host_out: pinned CPU tensors of batch*size complex elements.  Asynchronous with
        respect to the host until ``synchronize()``."""
        torch = self._torch
        n = self.chunk_batch * self.size
        hin = host_in.view(-1)
        hout = host_out.view(-1)
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for c in range(self.chunks):
            k = c % self.slots
            s, buf, plan = self.streams[k], self.bufs[k], self.plans[k]
            with torch.cuda.stream(s):
                buf.copy_(hin[c * n:(c + 1) * n], non_blocking=True)
                plan.execute(buf, batch=self.chunk_batch, inverse=inverse)
                hout[c * n:(c + 1) * n].copy_(buf, non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()

    @property
    def launch_count(self):
        return sum(p.launch_count for p in self.plans)

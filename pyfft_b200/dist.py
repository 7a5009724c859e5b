"""Multi-GPU execution: one process per GPU, ``torch.distributed`` for the plumbing.

The reference is single-device; both paths here are new functionality:

* ``shard_batch`` / ``BatchShardedPlan`` -- independent transforms: the batch is split contiguously
  over the ranks, every rank runs an ordinary ``Plan`` on its shard, no data-path collective.
* ``SlabPlan`` -- one large 3D transform of a (Z, Y, X) array distributed as z-slabs.  Forward:
  local X and Y passes, ONE exchange, local Z pass; the result is left y-slab distributed
  ("transposed out": rank h holds ``[Z][Y/G][X]`` for its y range).  The exchange is fused into the
  Y pass: its stores are destination-blocked (``b2fft_plan_set_output_blocks``) and, with
  ``exchange="p2p"``, go straight into the peers' receive buffers over NVLink (CUDA IPC mapped), so
  the transfer overlaps the butterflies tile by tile and no pack/unpack kernel exists.
  ``exchange="nccl"`` is the baseline: blocked stores into a local send buffer followed by
  ``all_to_all_single``.
  ``exchange="xslab"`` is the NVLink-efficient fused form: local Y pass first, then the X pass (whose
  lines are contiguous) scatters every output row as G contiguous X/G-element pieces straight into
  the peers' buffers, leaving the result x-slab distributed as ``[Y][Z][X/G]``; the X pass runs in
  y-chunks so the Z pass of chunk c overlaps the NVLink stores of chunk c+1 (see ``SlabPlan``).
"""
import ctypes

import numpy

from . import _lib
from .plan import _NP_DTYPES, _resolve_dtype, _stream_handle


# ----------------------------------------------------------------------------- batch sharding
def shard_batch(batch, world_size, rank):
    """Contiguous split of ``batch`` transforms: returns (first, count) for ``rank``."""
    batch, world_size, rank = int(batch), int(world_size), int(rank)
    if world_size < 1 or not 0 <= rank < world_size or batch < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(batch, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class BatchShardedPlan(object):
    """``Plan`` for a batch split over the ranks of a process group (no collective on the data path)."""

    def __init__(self, shape, dtype=numpy.complex64, group=None, **plan_kwargs):
        import torch.distributed as dist
        from .cuda import Plan
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = Plan(shape, dtype=dtype, **plan_kwargs)

    def local_range(self, global_batch):
        return shard_batch(global_batch, self.world, self.rank)

    def execute(self, *buffers, **kw):
        """Buffers hold this rank's shard; ``batch`` is the GLOBAL batch."""
        global_batch = kw.pop("batch", 1)
        _, count = self.local_range(global_batch)
        if count == 0:
            return None
        return self.plan.execute(*buffers, batch=count, **kw)


# ----------------------------------------------------------------------------- slab layout (pure host logic)
def slab_layout(shape, world_size, rank):
    """Index bookkeeping of the slab decomposition, in complex elements.

    Rank g owns z in [g*Zl, (g+1)*Zl) as ``[Zl][Y][X]``; after the exchange rank h owns
    y in [h*Yb, (h+1)*Yb) as ``[Z][Yb][X]``.
    """
    Z, Y, X = (int(s) for s in shape)
    G = int(world_size)
    if G < 1 or (G & (G - 1)):
        raise ValueError("number of ranks must be a power of two")
    if Z % G or Y % G:
        raise ValueError("Z and Y must be divisible by the number of ranks")
    Zl, Yb = Z // G, Y // G
    return {
        "Z": Z, "Y": Y, "X": X, "G": G, "Zl": Zl, "Yb": Yb,
        "slab_elems": Zl * Y * X, "yslab_elems": Z * Yb * X, "block_elems": Zl * Yb * X,
        # forward: Y pass of rank g writes y-block h to  recv_h + fwd_peer_offset  (P2P) or
        #          send + h * block_elems (NCCL), rows of length X, planes Yb*X apart
        "fwd_peer_offset": rank * Zl * Yb * X,
        "fwd_out_inner": X, "fwd_out_outer_stride": Yb * X,
        # inverse: Z pass of rank h writes z-block g to  slab_g + inv_peer_offset, z planes Y*X apart
        "inv_peer_offset": rank * Yb * X,
        "inv_out_inner": Y * X, "inv_out_outer_stride": 0,
        # "xslab" exchange: the X pass of rank g writes x-block h of row (z, y) to  xslab_h + (y*Z + g*Zl + z)*Xb
        "Xb": X // G if X % G == 0 else 0, "xslab_elems": Y * Z * (X // G),
        "xs_peer_offset": rank * Zl * (X // G), "xs_out_stride_y": Z * (X // G), "xs_out_stride_z": X // G,
        # "yzx" y-slab layout [Yb][Z][X]: the Z pass then strides by X only (a Y-pass-like access pattern)
        # instead of Yb*X; the Y pass's stores stride by Z*X between consecutive y instead of X.
        "fwd_peer_offset_yzx": rank * Zl * X, "fwd_out_inner_yzx": Z * X, "fwd_out_outer_stride_yzx": X,
        "inv_out_outer_stride_yzx": X,
    }


class _DeviceBuffer(object):
    """cudaMalloc'ed buffer whose IPC handle can be shared with the other ranks."""

    def __init__(self, nbytes, device):
        self._lib = _lib.load()
        self.nbytes = int(nbytes)
        self.device = int(device)
        p = ctypes.c_void_p()
        _lib.check(self._lib.b2fft_mem_alloc(self.nbytes, self.device, ctypes.byref(p)))
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def handle(self):
        buf = ctypes.create_string_buffer(64)
        _lib.check(self._lib.b2fft_ipc_export(self.ptr, buf))
        return buf.raw

    def tensor(self, tdtype):
        import torch
        return torch.as_tensor(self, device="cuda:%d" % self.device).view(tdtype)

    def free(self):
        if self.ptr:
            self._lib.b2fft_mem_free(self.ptr)
            self.ptr = None


def torch_real(t):
    """(re, im) view of a complex tensor: NCCL collectives have no complex dtypes."""
    import torch
    return torch.view_as_real(t).view(-1)


def _open_peers(handles, my_rank, my_ptr, device):
    lib = _lib.load()
    ptrs = []
    for r, h in enumerate(handles):
        if r == my_rank:
            ptrs.append(my_ptr)
        else:
            p = ctypes.c_void_p()
            _lib.check(lib.b2fft_ipc_import(h, device, ctypes.byref(p)))
            ptrs.append(p.value)
    return ptrs


class SlabPlan(object):
    """Slab-decomposed 3D C2C FFT over the ranks of a torch.distributed (NCCL) group.

    ``plan.slab``  : this rank's ``[Zl, Y, X]`` z-slab (input of ``forward``, output of ``inverse``)
    ``plan.yslab`` : this rank's ``[Z, Yb, X]`` y-slab (output of ``forward``, input of ``inverse``)
    Both are torch views of plan-owned, peer-visible device buffers; fill ``plan.slab`` in place.
    """

    def __init__(self, shape, dtype=numpy.complex64, group=None, normalize=True, scale=1.0, fast_math=True,
                 exchange="xslab", device=None, yslab_layout="zyx", chunks=1, exchange_ctas_per_sm=3, z_chunks=0,
                 overlap_sms=None, overlap_columns=0):
        import torch
        import torch.distributed as dist
        if len(shape) != 3:
            raise ValueError("SlabPlan needs a 3D shape (Z, Y, X)")
        if exchange not in ("p2p", "nccl", "xslab"):
            raise ValueError("exchange must be 'p2p', 'nccl' or 'xslab'")
        if yslab_layout not in ("zyx", "yzx"):
            raise ValueError("yslab_layout must be 'zyx' ([Z][Yb][X]) or 'yzx' ([Yb][Z][X])")
        if yslab_layout == "yzx" and exchange != "p2p":
            raise ValueError("the 'yzx' y-slab layout needs the fused p2p exchange")
        self.yslab_layout = yslab_layout
        # exchange="nccl" only: the slab is processed in `chunks` groups of z planes; the all-to-all of
        # chunk c runs on a side stream while the X/Y passes of chunk c+1 compute
        self.chunks = max(1, int(chunks))
        self.exchange_ctas_per_sm = int(exchange_ctas_per_sm)
        # x-slab mode: z_chunks = 0 lets the native plan choose (8 with the Y pass hidden under the exchange where that is
        # possible, else 1); overlap_sms: SMs left to the exchange while the Y pass runs (None = default, 0 = off)
        self.z_chunks = max(0, int(z_chunks))
        self.overlap_sms = overlap_sms
        self.overlap_columns = int(overlap_columns)
        self._normalize, self._scale, self._fast_math = bool(normalize), float(scale), bool(fast_math)
        self._native = None
        self._torch, self._dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.exchange = exchange
        self.dtype = _resolve_dtype(dtype)
        if self.dtype.kind != "c":
            raise ValueError("SlabPlan handles interleaved complex data")
        self.prec, self.layout = _NP_DTYPES[self.dtype]
        self.tdtype = torch.complex64 if self.dtype == numpy.complex64 else torch.complex128
        self.esz = self.dtype.itemsize
        L = self.L = slab_layout(shape, self.world, self.rank)
        lib = self._lib = _lib.load()

        def make(dims, axes, apply_scale):
            h = ctypes.c_void_p()
            d = (ctypes.c_int64 * 3)(*dims)
            _lib.check(lib.b2fft_plan_create_ex(ctypes.byref(h), d, axes, self.prec, self.layout, int(bool(normalize)),
                                                float(scale), int(bool(fast_math)), self.device,
                                                float(L["X"]) * L["Y"] * L["Z"], int(apply_scale)))
            return h

        X, Y, Z, Zl, Yb = L["X"], L["Y"], L["Z"], L["Zl"], L["Yb"]
        self._make = make
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda:%d" % self.device)
        self._send = None
        self._fwd_xy = self._fwd_z = self._inv_z = self._inv_xy = None
        if exchange == "xslab":
            self._init_xslab()
            return
        self._fwd_xy = make((X, Y, Zl), _lib.AXIS_X | _lib.AXIS_Y, 0)     # forward: scale applied by the Z pass
        yzx = yslab_layout == "yzx"
        if yzx:      # y-slab stored [Yb][Z][X]: the z axis is the middle ("y") axis of that array
            self._fwd_z = make((X, Z, Yb), _lib.AXIS_Y, 1)
            self._inv_z = make((X, Z, Yb), _lib.AXIS_Y, 0)
        else:
            self._fwd_z = make((X, Yb, Z), _lib.AXIS_Z, 1)
            self._inv_z = make((X, Yb, Z), _lib.AXIS_Z, 0)                # inverse: scale applied by the X/Y passes
        self._inv_xy = make((X, Y, Zl), _lib.AXIS_X | _lib.AXIS_Y, 1)

        self._slab_buf = _DeviceBuffer(L["slab_elems"] * self.esz, self.device)
        self._yslab_buf = _DeviceBuffer(L["yslab_elems"] * self.esz, self.device)
        self.slab = self._slab_buf.tensor(self.tdtype).view(Zl, Y, X)
        self.yslab = self._yslab_buf.tensor(self.tdtype).view(*((Yb, Z, X) if yzx else (Z, Yb, X)))
        G = self.world
        if exchange == "p2p" and G > 1:
            hs = [None] * G
            dist.all_gather_object(hs, (self._slab_buf.handle(), self._yslab_buf.handle()), group=group)
            self._peer_slab = _open_peers([h[0] for h in hs], self.rank, self._slab_buf.ptr, self.device)
            self._peer_yslab = _open_peers([h[1] for h in hs], self.rank, self._yslab_buf.ptr, self.device)
            sfx = "_yzx" if yzx else ""
            fwd_ptrs = [p + L["fwd_peer_offset" + sfx] * self.esz for p in self._peer_yslab]
            inv_ptrs = [p + L["inv_peer_offset"] * self.esz for p in self._peer_slab]
            self._set_blocks(self._fwd_xy, fwd_ptrs, L["fwd_out_inner" + sfx], L["fwd_out_outer_stride" + sfx])
            self._set_blocks(self._inv_z, inv_ptrs, L["inv_out_inner"], L["inv_out_outer_stride" + sfx])
        elif G > 1:
            self._send = torch.empty(L["slab_elems"], dtype=self.tdtype, device="cuda:%d" % self.device)
            C = self.chunks
            while Zl % C:
                C -= 1
            self.chunks = C
            Zc = Zl // C
            cb = Zc * Yb * X                                  # elements one rank receives from me per chunk
            self._chunk_plans = []
            for c in range(C):
                pl = make((X, Y, Zc), _lib.AXIS_X | _lib.AXIS_Y, 0)
                ptrs = [self._send.data_ptr() + (c * G + h) * cb * self.esz for h in range(G)]
                self._set_blocks(pl, ptrs, L["fwd_out_inner"], L["fwd_out_outer_stride"])
                self._chunk_plans.append(pl)
            self._comm_stream = torch.cuda.Stream(device=self.device)
            sendr = torch.view_as_real(self._send).view(C, G, cb * 2)
            recvr = torch.view_as_real(self.yslab.view(-1)).view(G, C, cb * 2)   # z = (g, c, z') -> natural z order
            self._send_lists = [[sendr[c, h] for h in range(G)] for c in range(C)]
            self._recv_lists = [[recvr[g, c] for g in range(G)] for c in range(C)]
        else:
            # single rank: the "exchange" is the Y pass writing into the y-slab buffer directly
            sfx = "_yzx" if yzx else ""
            self._set_blocks(self._fwd_xy, [self._yslab_buf.ptr], L["fwd_out_inner" + sfx], L["fwd_out_outer_stride" + sfx])
            self._set_blocks(self._inv_z, [self._slab_buf.ptr], L["inv_out_inner"], L["inv_out_outer_stride" + sfx])

    # ------------------------------------------------------------------ x-slab mode (native: csrc/slab.cu)
    def _init_xslab(self):
        """The whole schedule lives behind the C ABI (``b2fft_slab_*``, csrc/slab.cu): local Y pass per z-chunk, X pass
        per (z-chunk, y-chunk) whose destination-blocked stores put x-block h of every row straight into rank h's
        x-slab ``[Y][Z][Xb]`` over NVLink, one epoch word per (rank, y-chunk) in peer-mapped memory as the cross-rank
        signal, Z pass per y-chunk beside the stores of the following chunks; the inverse pulls the pieces back inside
        its X pass.  This class allocates the peer-visible buffers, exchanges their CUDA IPC handles through
        torch.distributed and attaches them."""
        torch, dist, L, lib = self._torch, self._dist, self.L, self._lib
        X, Y, Z, Zl, G = L["X"], L["Y"], L["Z"], L["Zl"], self.world
        if X % G:
            raise ValueError("X must be divisible by the number of ranks for the x-slab exchange")
        h = ctypes.c_void_p()
        dims = (ctypes.c_int64 * 3)(X, Y, Z)
        self._check(lib.b2fft_slab_plan_create(ctypes.byref(h), dims, self.prec, int(self._normalize), float(self._scale),
                                               int(self._fast_math), self.device, self.rank, G,
                                               self.chunks if self.chunks > 1 else 0, self.z_chunks,
                                               self.exchange_ctas_per_sm if G > 1 else 0))
        self._native = h
        if self.overlap_sms is not None:
            rc = lib.b2fft_slab_plan_set_overlap(h, int(self.overlap_sms))
            if rc not in (_lib.OK, _lib.E_UNSUPPORTED):
                self._check(rc)
        if self.overlap_columns:
            self._check(lib.b2fft_slab_plan_set_option(h, b"overlap_columns", float(self.overlap_columns)))
        geo = (ctypes.c_int64 * 8)()
        self._check(lib.b2fft_slab_plan_geometry(h, geo))
        self.chunks, self.z_chunks = int(geo[2]), int(geo[3])
        sb, xb, fb = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        self._check(lib.b2fft_slab_plan_sizes(h, ctypes.byref(sb), ctypes.byref(xb), ctypes.byref(fb)))
        Xb = L["Xb"]
        self._slab_buf = _DeviceBuffer(sb.value, self.device)
        self._yslab_buf = _DeviceBuffer(xb.value, self.device)
        self._flag_buf = _DeviceBuffer(max(fb.value, 256), self.device)
        self.slab = self._slab_buf.tensor(self.tdtype).view(Zl, Y, X)
        self.xslab = self._yslab_buf.tensor(self.tdtype).view(Y, Z, Xb)
        self.yslab = self.xslab                    # the distributed output, whatever its layout
        if G > 1:
            hs = [None] * G
            dist.all_gather_object(hs, (self._yslab_buf.handle(), self._flag_buf.handle()), group=self.group)
            self._peer_yslab = _open_peers([x[0] for x in hs], self.rank, self._yslab_buf.ptr, self.device)
            self._peer_flags = _open_peers([x[1] for x in hs], self.rank, self._flag_buf.ptr, self.device)
        else:
            self._peer_yslab, self._peer_flags = [self._yslab_buf.ptr], [self._flag_buf.ptr]
        self._peer_slab = None
        xs = (ctypes.c_void_p * G)(*self._peer_yslab)
        fl = (ctypes.c_void_p * G)(*self._peer_flags)
        self._check(lib.b2fft_slab_plan_attach(h, self._slab_buf.ptr, xs, fl))
        if G > 1:
            dist.barrier(group=self.group)         # every rank has cleared its flag words before anyone signals

    def _check(self, rc):
        if rc == _lib.OK:
            return
        msg = self._lib.b2fft_slab_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        if rc == _lib.E_INVALID:
            raise ValueError(msg)
        if rc == _lib.E_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise RuntimeError("b2fft: " + msg)

    def _forward_xslab(self):
        stream = self._torch.cuda.current_stream(self.device)
        self._check(self._lib.b2fft_slab_forward(self._native, _stream_handle(stream)))
        return self.xslab

    def _inverse_xslab(self):
        """x-slabs -> z-slabs: inverse Z pass per y-chunk, then the X pass pulls every row's pieces from the ranks'
        x-slabs over NVLink (source-blocked loads), then the local Y pass with the scale (csrc/slab.cu)."""
        stream = self._torch.cuda.current_stream(self.device)
        self._check(self._lib.b2fft_slab_inverse(self._native, _stream_handle(stream)))
        return self.slab

    def status(self):
        """0 = healthy; non-zero = a cross-rank wait timed out on this rank (synchronises the device)."""
        if getattr(self, "_native", None) is None:
            return 0
        out = ctypes.c_int(0)
        self._check(self._lib.b2fft_slab_plan_status(self._native, ctypes.byref(out)))
        return out.value

    def set_trace(self, on=True):
        if getattr(self, "_native", None) is not None:
            self._check(self._lib.b2fft_slab_plan_set_trace(self._native, int(bool(on))))

    def trace(self):
        """[(phase, ms since the start of the last traced forward)] -- synchronises the device."""
        if getattr(self, "_native", None) is None:
            return []
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self._lib.b2fft_slab_plan_trace(self._native, buf, len(buf)))
        return [(k, float(v)) for k, v in (item.split(":") for item in buf.value.decode().split(";") if item)]

    def describe(self):
        if getattr(self, "_native", None) is None:
            return "exchange=%s" % self.exchange
        buf = ctypes.create_string_buffer(2048)
        self._check(self._lib.b2fft_slab_plan_describe(self._native, buf, len(buf)))
        return buf.value.decode()

    # ------------------------------------------------------------------ helpers
    def _set_blocks(self, plan, ptrs, out_inner, out_outer_stride):
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        _lib.check(self._lib.b2fft_plan_set_output_blocks(plan, len(ptrs), arr, None, int(out_inner),
                                                          int(out_outer_stride)))

    def _exec(self, plan, src_ptr, dst_ptr, inverse):
        stream = self._torch.cuda.current_stream(self.device)
        _lib.check(self._lib.b2fft_execute(plan, src_ptr, None, dst_ptr, None, int(inverse), 1, _stream_handle(stream)))

    def _sync_ranks(self):
        """Stream-ordered cross-rank barrier (tiny all-reduce; does not block the host)."""
        if self.world > 1:
            self._dist.all_reduce(self._flag, group=self.group)

    @property
    def launch_count(self):
        if self._native is not None:
            return int(self._lib.b2fft_slab_plan_launch_count(self._native))
        plans = [self._fwd_xy, self._fwd_z, self._inv_z, self._inv_xy] + list(getattr(self, "_chunk_plans", []))
        return sum(int(self._lib.b2fft_plan_launch_count(p)) for p in plans if p is not None)

    # ------------------------------------------------------------------ transforms
    def forward(self):
        """``plan.slab`` (destroyed) -> ``plan.yslab``.  Asynchronous on the current stream."""
        L, dist = self.L, self._dist
        if self.exchange == "xslab":
            return self._forward_xslab()
        self._sync_ranks()                  # every rank is done with its previous y-slab contents
        if self.world > 1 and self.exchange == "nccl":
            # per z-chunk: X pass in place, Y pass with destination-blocked stores into the send buffer,
            # then that chunk's all-to-all on the side stream while the next chunk computes
            torch = self._torch
            cur = torch.cuda.current_stream(self.device)
            chunk_bytes = (L["slab_elems"] // self.chunks) * self.esz
            works = []
            for c, pl in enumerate(self._chunk_plans):
                ptr = self._slab_buf.ptr + c * chunk_bytes
                self._exec(pl, ptr, ptr, 0)
                self._comm_stream.wait_stream(cur)
                with torch.cuda.stream(self._comm_stream):
                    works.append(dist.all_to_all(self._recv_lists[c], self._send_lists[c], group=self.group,
                                                 async_op=True))
            for w in works:
                w.wait()                    # current stream waits for the exchange
            cur.wait_stream(self._comm_stream)
        else:
            # X pass in place on the slab, then the Y pass whose stores are the exchange
            self._exec(self._fwd_xy, self._slab_buf.ptr, self._slab_buf.ptr, 0)
            self._sync_ranks()              # all peers' blocks have landed in my y-slab
        self._exec(self._fwd_z, self._yslab_buf.ptr, self._yslab_buf.ptr, 0)
        return self.yslab

    def inverse(self):
        """``plan.yslab`` (destroyed) -> ``plan.slab``."""
        L, dist, torch = self.L, self._dist, self._torch
        G = self.world
        if self.exchange == "xslab":
            return self._inverse_xslab()
        self._sync_ranks()
        if G > 1 and self.exchange == "nccl":
            self._exec(self._inv_z, self._yslab_buf.ptr, self._yslab_buf.ptr, 1)
            tmp = self._send
            dist.all_to_all_single(torch_real(tmp), torch_real(self.yslab.view(-1)), group=self.group)
            # tmp = [h][Zl][Yb][X] -> slab [Zl][h*Yb + yl][X]
            self.slab.view(L["Zl"], G, L["Yb"], L["X"]).copy_(
                tmp.view(G, L["Zl"], L["Yb"], L["X"]).permute(1, 0, 2, 3))
        else:
            self._exec(self._inv_z, self._yslab_buf.ptr, self._yslab_buf.ptr, 1)   # stores land in the peers' slabs
            self._sync_ranks()
        self._exec(self._inv_xy, self._slab_buf.ptr, self._slab_buf.ptr, 1)
        return self.slab

    def close(self):
        if getattr(self, "_native", None) is not None:
            self._torch.cuda.synchronize(self.device)
            if self.world > 1:
                self._dist.barrier(group=self.group)       # nobody still reads or writes a peer's buffers
            self._lib.b2fft_slab_plan_destroy(self._native)
            self._native = None
            if self.world > 1:
                for table in (self._peer_yslab, getattr(self, "_peer_flags", [])):
                    for r, p in enumerate(table):
                        if r != self.rank:
                            self._lib.b2fft_ipc_release(p)
                self._dist.barrier(group=self.group)       # peers have unmapped my buffers before I free them
            self._peer_yslab = self._peer_flags = None
            self.slab = self.yslab = self.xslab = None
            for b in ("_slab_buf", "_yslab_buf", "_flag_buf"):
                buf = getattr(self, b, None)
                if buf is not None:
                    buf.free()
                    setattr(self, b, None)
            return
        for p in ("_fwd_xy", "_fwd_z", "_inv_z", "_inv_xy"):
            h = getattr(self, p, None)
            if h is not None:
                self._lib.b2fft_plan_destroy(h)
                setattr(self, p, None)
        for h in getattr(self, "_chunk_plans", []):
            self._lib.b2fft_plan_destroy(h)
        self._chunk_plans = []
        self._send_lists = self._recv_lists = None
        self.slab = self.yslab = self.xslab = None
        if self.exchange == "xslab" and getattr(self, "_peer_yslab", None) and self.world > 1:
            for r, p in enumerate(self._peer_yslab):
                if r != self.rank:
                    self._lib.b2fft_ipc_release(p)
            self._peer_yslab = None
        if getattr(self, "_peer_slab", None):
            for r, p in enumerate(self._peer_slab):
                if r != self.rank:
                    self._lib.b2fft_ipc_release(p)
            for r, p in enumerate(self._peer_yslab):
                if r != self.rank:
                    self._lib.b2fft_ipc_release(p)
            self._peer_slab = self._peer_yslab = None
        for b in ("_slab_buf", "_yslab_buf"):
            buf = getattr(self, b, None)
            if buf is not None:
                buf.free()

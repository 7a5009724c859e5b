"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: batch sharding and the slab
decomposition's layout arithmetic.  The FFT compute is stood in for by numpy (test infrastructure);
what is checked is that the block offsets / strides the CUDA Y pass is given (slab_layout) route
every element to the right rank and position, in both exchange flavours."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyfft_b200.dist import shard_batch, slab_layout


def test_shard_batch_covers_everything():
    for batch in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(batch, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == batch
            pos = 0
            for first, count in spans:
                assert first == pos
                pos += count
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_batch(4, 2, 2)


def test_slab_layout_rejects_bad_shapes():
    with pytest.raises(ValueError):
        slab_layout((6, 8, 8), 4, 0)
    with pytest.raises(ValueError):
        slab_layout((8, 8, 8), 3, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _blocked_store(L, yfft, send_or_peers, p2p):
    """Emulates the Y pass's destination-blocked stores with the layout numbers the kernel gets."""
    Zl, Y, X, Yb, G = L["Zl"], L["Y"], L["X"], L["Yb"], L["G"]
    flat = yfft.reshape(Zl, Y, X)
    for z in range(Zl):
        for y in range(Y):
            h, r = divmod(y, Yb)
            base = z * L["fwd_out_outer_stride"] + r * L["fwd_out_inner"]
            if p2p:
                send_or_peers[h][L["fwd_peer_offset"] + base:L["fwd_peer_offset"] + base + X] = flat[z, y]
            else:
                send_or_peers[h * L["block_elems"] + base:h * L["block_elems"] + base + X] = flat[z, y]


def _xslab_blocked_store(L, xfft, peers, chunks):
    """Emulates the x-slab exchange pass (pyfft_b200/dist.py _init_xslab): the X pass walks the rows
    {all local z} x {y chunk c} through the two-level outer index given to
    b2fft_plan_set_outer_split and writes x-block h of each row to peers[h]."""
    Zl, Y, X, Z, Xb, G = L["Zl"], L["Y"], L["X"], L["Z"], L["Xb"], L["G"]
    Yc = Y // chunks
    src = xfft.reshape(-1)
    for c in range(chunks):
        in0 = c * Yc * X                                             # pointer offsets the host adds per chunk
        blk0 = c * Yc * Z * Xb + L["xs_peer_offset"]
        outer_div, in_lo, in_hi, out_lo, out_hi = Yc, X, Y * X, L["xs_out_stride_y"], L["xs_out_stride_z"]
        for o in range(Yc * Zl):                                     # the pass's dense row index
            hi, lo = divmod(o, outer_div)
            row = src[in0 + hi * in_hi + lo * in_lo:in0 + hi * in_hi + lo * in_lo + X]
            for h in range(G):
                off = blk0 + hi * out_hi + lo * out_lo
                peers[h][off:off + Xb] = row[h * Xb:(h + 1) * Xb]


def _xslab_pull(L, xslabs, chunks):
    """Emulates the inverse x-slab exchange pass (csrc/slab.cu b2fft_slab_inverse): per y-chunk the X pass walks the rows
    {all local z} x {y in chunk} through the two-level outer index and LOADS piece h of every row from rank h's x-slab
    (b2fft_plan_set_input_blocks + b2fft_plan_set_outer_split, same numbers as slab.cu passes); returns the z-slab rows."""
    Zl, Y, X, Z, Xb, G = L["Zl"], L["Y"], L["X"], L["Z"], L["Xb"], L["G"]
    Yc = Y // chunks
    rank_off = L["xs_peer_offset"]                                  # rank * Zl * Xb
    out = np.zeros(Zl * Y * X, dtype=np.complex128)
    for c in range(chunks):
        blk0 = c * Yc * Z * Xb + rank_off                            # base of every source block for this chunk
        out0 = c * Yc * X
        outer_div, in_lo, in_hi, out_lo, out_hi = Yc, Z * Xb, Xb, X, Y * X
        for o in range(Yc * Zl):
            hi, lo = divmod(o, outer_div)                            # (local z, y in chunk)
            for h in range(G):
                src = blk0 + hi * in_hi + lo * in_lo
                dst = out0 + hi * out_hi + lo * out_lo + h * Xb
                out[dst:dst + Xb] = xslabs[h][src:src + Xb]
    return out.reshape(Zl, Y, X)


def _worker(rank, world, port, shape, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Z, Y, X = shape
        L = slab_layout(shape, world, rank)
        rng = np.random.default_rng(5)
        full = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        want = np.fft.fftn(full)
        slab = full[rank * L["Zl"]:(rank + 1) * L["Zl"]]
        xy = np.fft.fft(np.fft.fft(slab, axis=2), axis=1)                 # local X and Y passes
        # NCCL-flavoured exchange: blocked stores into a send buffer, then all_to_all
        send = np.zeros(L["slab_elems"], dtype=np.complex128)
        _blocked_store(L, xy, send, p2p=False)
        st = torch.from_numpy(np.stack([send.real, send.imag], axis=1).copy())
        rt = torch.empty_like(st)
        dist.all_to_all_single(rt, st)
        recv = (rt[:, 0] + 1j * rt[:, 1]).numpy().reshape(Z, L["Yb"], X)
        got = np.fft.fft(recv, axis=0)                                    # local Z pass
        ref = want[:, rank * L["Yb"]:(rank + 1) * L["Yb"], :]
        err_nccl = float(np.abs(got - ref).max() / np.abs(ref).max())
        # P2P-flavoured exchange: every rank's blocked stores applied to the peers' buffers; emulate by
        # gathering all slabs and replaying each source rank's stores into this rank's buffer
        gathered = [None] * world
        dist.all_gather_object(gathered, xy)
        mine = np.zeros(L["yslab_elems"], dtype=np.complex128)
        for src in range(world):
            Ls = slab_layout(shape, world, src)
            bufs = [np.zeros(L["yslab_elems"], dtype=np.complex128) if h != rank else mine for h in range(world)]
            _blocked_store(Ls, gathered[src], bufs, p2p=True)
        got2 = np.fft.fft(mine.reshape(Z, L["Yb"], X), axis=0)
        err_p2p = float(np.abs(got2 - ref).max() / np.abs(ref).max())
        # inverse layout: z-block g of rank h's y-slab goes to slab_g + inv_peer_offset with stride Y*X
        zinv = np.fft.ifft(want[:, rank * L["Yb"]:(rank + 1) * L["Yb"], :], axis=0)
        gathered = [None] * world
        dist.all_gather_object(gathered, zinv)
        myslab = np.zeros(L["slab_elems"], dtype=np.complex128)
        for src in range(world):
            Ls = slab_layout(shape, world, src)
            blk = gathered[src][rank * L["Zl"]:(rank + 1) * L["Zl"]]      # [Zl][Yb][X] destined for me
            for z in range(L["Zl"]):
                for yl in range(L["Yb"]):
                    off = Ls["inv_peer_offset"] + z * Ls["inv_out_inner"] + yl * X
                    myslab[off:off + X] = blk[z, yl]
        back = np.fft.ifft(np.fft.ifft(myslab.reshape(L["Zl"], Y, X), axis=1), axis=2)
        err_inv = float(np.abs(back - slab).max())
        # x-slab exchange: Y pass, then the X pass scatters rows into the ranks' [Y][Z][Xb] x-slabs
        yx = np.fft.fft(np.fft.fft(slab, axis=1), axis=2)
        gathered = [None] * world
        dist.all_gather_object(gathered, yx)
        minex = np.zeros(L["xslab_elems"], dtype=np.complex128)
        for src in range(world):
            Ls = slab_layout(shape, world, src)
            bufs = [np.zeros(L["xslab_elems"], dtype=np.complex128) if h != rank else minex for h in range(world)]
            _xslab_blocked_store(Ls, gathered[src], bufs, chunks=2)
        got3 = np.fft.fft(minex.reshape(Y, Z, L["Xb"]), axis=1)          # Z pass on [Y][Z][Xb]
        ref3 = want[:, :, rank * L["Xb"]:(rank + 1) * L["Xb"]].transpose(1, 0, 2)
        err_xs = float(np.abs(got3 - ref3).max() / np.abs(ref3).max())
        # fused inverse of the x-slab mode: inverse Z pass in place on every rank's x-slab, then this rank's X pass pulls
        # the pieces of its rows from all x-slabs, then the local Y pass
        zinv_x = np.fft.ifft(got3, axis=1)                                # [Y][Z][Xb], z back in the space domain
        gathered = [None] * world
        dist.all_gather_object(gathered, zinv_x.reshape(-1))
        rows = _xslab_pull(L, gathered, chunks=2)                          # [Zl][Y][X], still transformed along y and x
        back_x = np.fft.ifft(np.fft.ifft(rows, axis=2), axis=1)
        err_pull = float(np.abs(back_x - slab).max())
        q.put((rank, err_nccl, err_p2p, err_inv, max(err_xs, err_pull)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 4, 16), (4, 8, 8)])
def test_slab_exchange_layout_world2(shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, e3, e4 in res:
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12 and e4 < 1e-12, (rank, e1, e2, e3, e4)


def _schedule(built_lib, ranks, C, K, hidden, cols=0):
    import ctypes
    buf = ctypes.create_string_buffer(1 << 14)
    assert built_lib.b2fft_slab_schedule_preview(ranks, C, K, int(hidden), cols, buf, len(buf)) == 0
    cells = []
    for item in buf.value.decode().strip(";").split(";"):
        k, c = item.split(":")
        ks = list(range(K)) if k == "all" else [int(k)]
        cs = list(range(int(c.split("-")[0]), int(c.split("-")[1]) + 1)) if "-" in c else [int(c)]
        cells.append((k == "all", ks, cs))
    return cells


@pytest.mark.parametrize("ranks,C,K,hidden,cols", [(8, 8, 8, 1, 0), (4, 8, 8, 1, 0), (8, 16, 4, 1, 0), (8, 8, 4, 1, 8), (2, 8, 1, 0, 0),
                                                  (8, 8, 4, 0, 0), (16, 8, 8, 1, 0), (8, 8, 8, 1, 1)])
def test_slab_launch_order(built_lib, ranks, C, K, hidden, cols):
    """csrc/slab.cu's order of the X-pass launches (b2fft_slab_schedule_preview, no device): every (z-chunk, y-chunk) cell is
    sent exactly once; a launch never needs a z-chunk later than the ones the exchange stream has already waited for plus the
    next one (the Y launch completes its chunks in order); with the hidden Y pass the columns that leave z-chunk by z-chunk
    come first and ALL of them complete before the first whole-column launch, whose columns then complete one at a time (so
    their Z passes spread out); the default number of early columns follows the rank count (3 of 8 on 8 ranks, 4 of 8 on 4)."""
    cells = _schedule(built_lib, ranks, C, K, hidden, cols)
    seen = {}
    waited = -1                      # highest z-chunk whose Y pass the exchange stream has waited for
    completed = []                   # columns in the order they complete
    for whole, ks, cs in cells:
        need = max(ks)
        assert need <= waited + K if whole else need <= waited + 1
        waited = max(waited, need)
        for k in ks:
            for c in cs:
                assert (k, c) not in seen
                seen[(k, c)] = True
        for c in cs:
            if all((k, c) in seen for k in range(K)) and c not in completed:
                completed.append(c)
    assert len(seen) == K * C and sorted(completed) == list(range(C))
    if hidden:
        early = cells[0][2]
        want = cols if cols else max(1, min(C, int(C * 0.364 * ranks / (ranks - 1) + 0.5)))      # 8 ranks: 3 of 8; 4 ranks: 4 of 8
        assert cols or C != 8 or want == {8: 3, 4: 4, 16: 3}[ranks]
        assert len(early) == min(want, C) and early == list(range(len(early)))
        assert [c[1] for c in cells[:K]] == [[k] for k in range(K)] and all(c[2] == early for c in cells[:K])
        late = cells[K:]
        assert all(whole for whole, _, _ in late) and [cs for _, _, cs in late] == [[c] for c in range(len(early), C)]
        assert completed == list(range(C))
    else:
        assert all(not whole and len(cs) == 1 for whole, _, cs in cells) and len(cells) == K * C

"""Host feeding helpers.

* ``decode_jpeg_batch``: compressed tiles -> ``torch.uint8 [B,H,W,3]`` on the GPU through nvJPEG (``sb_decode_jpeg``);
  the B200-native replacement of the PIL loading step of the reference's callers (stainlib_normalization.ipynb:61-74).
* ``stream_host_batches``: runs ANY device-batch operator of this package (a normaliser's ``transform``, an augmenter,
  a composition) over a pinned host batch in overlapped chunks -- H2D, compute and D2H on three streams with three
  slots, the Python counterpart of ``sb_normalize_host`` for the operators that have no fused host entry point.
"""
import ctypes

import torch

from stainlib_b200 import _native as nv


def decode_jpeg_batch(jpegs, H, W, device=None, out=None, sync=True):
    """jpegs: sequence of ``bytes`` / ``bytearray`` / 1-D uint8 tensors holding baseline JPEG streams, every one H x W.
    Returns (or fills ``out``) a uint8 [B,H,W,3] RGB CUDA tensor, stream-ordered on the current stream.  ``sync=False``
    returns ``(out, keepalive)`` without waiting: the caller keeps ``keepalive`` (the host buffers nvJPEG still reads) until
    the stream has passed this point."""
    h, idx = nv.get_handle(device)
    B = len(jpegs)
    bufs = [bytes(j) if not isinstance(j, torch.Tensor) else bytes(j.cpu().numpy().tobytes()) for j in jpegs]
    ptrs = (ctypes.c_void_p * B)(*[ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p) for b in bufs])
    sizes = (ctypes.c_size_t * B)(*[len(b) for b in bufs])
    if out is None:
        out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=f"cuda:{idx}")
    assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (B, H, W, 3) and out.is_contiguous()
    nv.check(nv.load_library().sb_decode_jpeg(h, ptrs, sizes, B, int(H), int(W), nv.ptr(out), nv.stream_ptr(idx)))
    if not sync:
        return out, bufs
    torch.cuda.current_stream(idx).synchronize()          # nvJPEG reads the host buffers asynchronously: keep them alive
    return out


def stream_jpeg_batches(op, jpegs, H, W, host_out=None, chunk_tiles=512, device=None, keep_on_device=False):
    """``host_out[i] = op(decode(jpegs[i]))`` with the three stages overlapped: nvJPEG decode of chunk k+1 (its host-side parsing
    and its GPU Huffman / IDCT kernels on a decode stream), ``op`` on chunk k and the copy back of chunk k-1, on three streams
    with three device slots.  The compressed counterpart of ``stream_host_batches``; ``keep_on_device=True`` returns one
    [B,H,W,3] CUDA tensor instead of copying back.

    Chunk size: nvJPEG's batched decode has a large fixed cost per call and decodes Huffman streams on the GPU only for
    batches above ~100 images (measured on B200, 512^2 tiles: 0.4 Gpx/s at 16-32 tiles per call, 2.2 at 128, 4.4 at 256, 7.2 at
    512, 9.5 at 1024, 12.3 at 2048 -- profiles/r02_jpeg_probe.txt), so chunks should hold many hundreds of tiles; a batch that fits
    the device in one piece is fastest through ``decode_jpeg_batch`` + ``op`` directly, and this pipeline is for streams that do
    not fit (its 3 slots bound the device memory at 3 x chunk)."""
    _, idx = nv.get_handle(device)
    dev = torch.device("cuda", idx)
    B = len(jpegs)
    chunk_tiles = max(1, min(int(chunk_tiles), B))
    dev_out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) if keep_on_device else None
    if not keep_on_device and host_out is None:
        host_out = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
    if idx not in _STREAMS:
        _STREAMS[idx] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    s_dec, s_comp, s_out = _STREAMS[idx]
    n_slot = 3
    d_in = [torch.empty((chunk_tiles, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(n_slot)]
    d_out = [None] * n_slot
    ev_in = [torch.cuda.Event() for _ in range(n_slot)]
    ev_comp = [torch.cuda.Event() for _ in range(n_slot)]
    ev_out = [torch.cuda.Event() for _ in range(n_slot)]
    keep = []
    cur = torch.cuda.current_stream(dev)
    for s in (s_dec, s_comp, s_out):
        s.wait_stream(cur)
    for c, t0 in enumerate(range(0, B, chunk_tiles)):
        slot, nt = c % n_slot, min(chunk_tiles, B - t0)
        if c >= n_slot:
            s_dec.wait_event(ev_comp[slot])             # the slot's previous input has been consumed
            if not keep_on_device:
                s_comp.wait_event(ev_out[slot])
        with torch.cuda.stream(s_dec):
            _, bufs = decode_jpeg_batch(jpegs[t0:t0 + nt], H, W, device=idx, out=d_in[slot][:nt], sync=False)
            keep.append(bufs)
            ev_in[slot].record(s_dec)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in[slot])
            d_out[slot] = op(d_in[slot][:nt])
            if keep_on_device:
                dev_out[t0:t0 + nt].copy_(d_out[slot])
            ev_comp[slot].record(s_comp)
        if keep_on_device:
            continue
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[slot])
            host_out[t0:t0 + nt].copy_(d_out[slot], non_blocking=True)
            d_out[slot].record_stream(s_out)
            ev_out[slot].record(s_out)
    s_out.synchronize()
    s_comp.synchronize()
    s_dec.synchronize()
    del keep
    if keep_on_device:
        cur.wait_stream(s_comp)
        return dev_out
    return host_out


_STREAMS = {}      # device index -> (s_in, s_comp, s_out): kept across calls (stream-ordered allocations stay on one stream)
_SLOTS = {}        # (device index, chunk shape) -> three device input slots


def stream_host_batches(op, host_in, host_out=None, chunk_tiles=0, device=None, keep_on_device=False):
    """``host_out[i] = op(host_in[i])`` tile batch by tile batch with copies and compute overlapped.

    op: callable taking and returning a uint8 [b,H,W,3] CUDA tensor (e.g. ``lambda x: rein.transform(hed.transform(x))``).
    host_in / host_out: uint8 [B,H,W,3] CPU tensors, pinned for full speed.  chunk_tiles = 0 picks ~96 MB chunks: the
    per-chunk cost of the Python-level operator calls (~0.2 ms) wants chunks of at least tens of MB.  Returns host_out.
    ``keep_on_device=True``: the results stay in HBM (one [B,H,W,3] CUDA tensor is returned, nothing is copied back) --
    the case of a training loop that consumes the normalised tiles on the GPU."""
    _, idx = nv.get_handle(device)
    dev = torch.device("cuda", idx)
    dev_out = None
    if keep_on_device:
        dev_out = torch.empty(host_in.shape, dtype=torch.uint8, device=dev)
    elif host_out is None:
        host_out = torch.empty_like(host_in, pin_memory=host_in.is_pinned())
    B = host_in.shape[0]
    tile_bytes = int(host_in[0].numel())
    if chunk_tiles <= 0:
        chunk_tiles = max(1, (96 << 20) // tile_bytes)
    chunk_tiles = min(chunk_tiles, B)
    if idx not in _STREAMS:
        _STREAMS[idx] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    s_in, s_comp, s_out = _STREAMS[idx]
    n_slot = 3
    key = (idx, (chunk_tiles,) + tuple(host_in.shape[1:]))
    if key not in _SLOTS:
        _SLOTS.clear()                                  # one geometry at a time: do not hoard device memory
        _SLOTS[key] = [torch.empty(key[1], dtype=torch.uint8, device=dev) for _ in range(n_slot)]
    d_in = _SLOTS[key]
    d_out = [None] * n_slot
    ev_in = [torch.cuda.Event() for _ in range(n_slot)]
    ev_comp = [torch.cuda.Event() for _ in range(n_slot)]
    ev_out = [torch.cuda.Event() for _ in range(n_slot)]
    cur = torch.cuda.current_stream(dev)
    for s in (s_in, s_comp, s_out):
        s.wait_stream(cur)                              # order behind whatever the caller queued (and earlier calls)
    for c, t0 in enumerate(range(0, B, chunk_tiles)):
        slot, nt = c % n_slot, min(chunk_tiles, B - t0)
        if c >= n_slot:
            s_in.wait_event(ev_comp[slot])              # the slot's previous input has been consumed
            if not keep_on_device:
                s_comp.wait_event(ev_out[slot])         # ... and its previous output copied out
        with torch.cuda.stream(s_in):
            d_in[slot][:nt].copy_(host_in[t0:t0 + nt], non_blocking=True)
            ev_in[slot].record(s_in)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in[slot])
            d_out[slot] = op(d_in[slot][:nt])
            if keep_on_device:
                dev_out[t0:t0 + nt].copy_(d_out[slot])
            ev_comp[slot].record(s_comp)
        if keep_on_device:
            continue
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[slot])
            host_out[t0:t0 + nt].copy_(d_out[slot], non_blocking=True)
            d_out[slot].record_stream(s_out)
            ev_out[slot].record(s_out)
    s_out.synchronize()
    s_comp.synchronize()
    s_in.synchronize()
    if keep_on_device:
        cur.wait_stream(s_comp)
        return dev_out
    return host_out

"""Probe of io.stream_host_batches: the skeleton alone (op = device copy) and with real operators, several chunk sizes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.io import stream_host_batches
from stainlib_b200.synth import synth_batch, synth_tile
B, H, W = 1024, 512, 512
host = torch.from_numpy(synth_batch(5000, 64, H, W)).repeat(B // 64, 1, 1, 1).contiguous().pin_memory()
out = torch.empty_like(host).pin_memory()
rein = sb.ReinhardStainNormalizer(); rein.fit(synth_tile(1, H, W, kind="target"))
mac = sb.ExtractiveStainNormalizer("macenko"); mac.fit(synth_tile(1, H, W, kind="target"))
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
for name, op in (("copy", lambda x: x.clone()), ("reinhard", lambda x: rein.transform(x)), ("macenko", lambda x: mac.transform(x))):
    for chunk in (16, 64, 128):
        dt = timed(lambda: stream_host_batches(op, host, out, chunk_tiles=chunk))
        print(f"{name:9s} chunk {chunk:4d}: {dt * 1e3:7.2f} ms  {B * H * W / dt / 1e9:6.2f} Gpx/s")
dt = timed(lambda: mac.transform(host, out=out))
print(f"sb_normalize_host (macenko): {dt * 1e3:7.2f} ms  {B * H * W / dt / 1e9:6.2f} Gpx/s")

"""A/B of the chunk size of stream_host_batches (pinned host -> HED-light + Reinhard -> pinned host), one process.
python tools/stream_chunk_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.augmentation.augmenter import HedLightColorAugmenter
from stainlib_b200.io import stream_host_batches
from stainlib_b200.synth import synth_batch, synth_tile
B, H, W = 1024, 512, 512
host_in = torch.from_numpy(synth_batch(5000, 64, H, W)).repeat(B // 64, 1, 1, 1).contiguous().pin_memory()
host_out = torch.empty_like(host_in).pin_memory()
hed = HedLightColorAugmenter(); rein = sb.ReinhardStainNormalizer(); rein.fit(synth_tile(1, H, W, kind="target"))
rng = np.random.default_rng(0)
sig, bias = rng.uniform(-0.1, 0.1, (B, 3)), rng.uniform(-0.1, 0.1, (B, 3))
def run(chunk):
    pos = {"t0": 0}
    def op(x):
        t0 = pos["t0"]; pos["t0"] += x.shape[0]
        return rein.transform(hed.transform(x, sigmas=sig[t0:t0 + x.shape[0]], biases=bias[t0:t0 + x.shape[0]]))
    stream_host_batches(op, host_in, host_out, chunk_tiles=chunk)
    torch.cuda.synchronize()
for rep in range(2):
    for chunk in (128, 64, 32, 16):
        run(chunk); run(chunk)
        t0 = time.perf_counter()
        for _ in range(5): run(chunk)
        dt = (time.perf_counter() - t0) / 5
        print(f"chunk {chunk:4d} tiles ({chunk * H * W * 3 >> 20:3d} MB): {B * H * W / dt / 1e6:8.1f} Mpx/s  {dt * 1e3:6.2f} ms")

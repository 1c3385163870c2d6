"""End to end from COMPRESSED tiles: nvJPEG decode (sb_decode_jpeg) -> Macenko transform -> D2H into pinned memory, per nvJPEG
backend, beside the raw-pixel feed.  python tools/jpeg_probe.py [tiles] [H]"""
import os, sys, time
import cv2, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.io import decode_jpeg_batch
from stainlib_b200.synth import synth_batch, synth_tile

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
pool = synth_batch(5000, 32, H, H)
jp = [cv2.imencode(".jpg", cv2.cvtColor(t, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for t in pool]
jpegs = [jp[i % 32] for i in range(B)]
cbytes = sum(len(j) for j in jpegs)
norm = sb.ExtractiveStainNormalizer("macenko"); norm.fit(synth_tile(1, H, kind="target"))
host_out = torch.empty((B, H, H, 3), dtype=torch.uint8).pin_memory()
dev = torch.empty((B, H, H, 3), dtype=torch.uint8, device="cuda")
npx = B * H * H
print(f"# {B} tiles of {H}x{H}; JPEG q90 4:2:0: {cbytes / B / 1024:.0f} KB per tile ({npx * 3 / cbytes:.1f}x smaller than raw)")
for be in (("default", "gpu_hybrid", "hardware") if B <= 512 else ("gpu_hybrid",)):
    os.environ["SB_NVJPEG_BACKEND"] = be
    try:
        def step():
            decode_jpeg_batch(jpegs, H, H, out=dev)
            host_out.copy_(norm.transform(dev), non_blocking=True)
            torch.cuda.synchronize()
        step(); step()
        t0 = time.perf_counter(); n = 3
        for _ in range(n): step()
        dt = (time.perf_counter() - t0) / n
        t0 = time.perf_counter()
        for _ in range(n): decode_jpeg_batch(jpegs, H, H, out=dev)
        dd = (time.perf_counter() - t0) / n
        print(f"backend {be:10s}: decode + transform + D2H {npx / dt / 1e6:8.1f} Mpx/s ({dt * 1e3:.1f} ms); decode alone {npx / dd / 1e6:8.1f} Mpx/s, {B / dd:7.0f} tiles/s")
    except Exception as e:
        print(f"backend {be:10s}: unavailable ({type(e).__name__}: {e})")
from stainlib_b200.io import stream_jpeg_batches
os.environ["SB_NVJPEG_BACKEND"] = "gpu_hybrid"
for chunk in (128, 256, 512):
    def piped():
        stream_jpeg_batches(norm.transform, jpegs, H, H, host_out=host_out, chunk_tiles=chunk)
        torch.cuda.synchronize()
    piped(); piped()
    t0 = time.perf_counter()
    for _ in range(3): piped()
    dt = (time.perf_counter() - t0) / 3
    print(f"overlapped pipeline, chunks of {chunk:3d} tiles: {npx / dt / 1e6:8.1f} Mpx/s ({dt * 1e3:.1f} ms)")
host_in = torch.from_numpy(np.stack([pool[i % 32] for i in range(B)])).pin_memory()
def raw():
    norm.transform(host_in, out=host_out)
raw(); raw()
t0 = time.perf_counter()
for _ in range(3): raw()
dt = (time.perf_counter() - t0) / 3
print(f"raw pixels from pinned host memory (sb_normalize_host): {npx / dt / 1e6:8.1f} Mpx/s")

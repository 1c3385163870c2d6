"""Host<->device link probe for the e2e leg of bench.py: pinned H2D, D2H and simultaneous both-way copy rates, and a
sweep of sb_normalize_host's chunk size.  Run on the GPU box: python tools/pcie_probe.py > gpurun_out/pcie.txt"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = 768 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def t(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    dt = t(lambda: d_a.copy_(h_in, non_blocking=True))
    print(f"H2D pinned: {n / dt / 1e9:.1f} GB/s")
    dt = t(lambda: h_out.copy_(d_b, non_blocking=True))
    print(f"D2H pinned: {n / dt / 1e9:.1f} GB/s")

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
    dt = t(both)
    print(f"both ways at once: {n / dt / 1e9:.1f} GB/s each direction ({dt * 1e3:.2f} ms for {n >> 20} MiB each way)")
    for chunk in (4, 16, 48):
        c = chunk << 20

        def chunked():
            for o in range(0, n, c):
                with torch.cuda.stream(s1):
                    d_a[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out[o:o + c].copy_(d_b[o:o + c], non_blocking=True)
        dt = t(chunked)
        print(f"both ways, {chunk} MiB chunks: {n / dt / 1e9:.1f} GB/s each direction")
    del d_a, d_b, h_in, h_out

    import stainlib_b200 as sb
    from stainlib_b200.synth import synth_tile, synth_batch
    B, H, W = 1024, 512, 512
    norm = sb.ExtractiveStainNormalizer("macenko")
    norm.fit(synth_tile(1, H, W, kind="target"))
    pool = torch.from_numpy(synth_batch(5000, 64, H, W))
    host_in = pool.repeat(B // 64, 1, 1, 1).contiguous().pin_memory()
    for chunk_tiles in (0, 8, 16, 32, 64, 128, 256):
        for _ in range(2):
            norm.transform(host_in, chunk_tiles=chunk_tiles)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            norm.transform(host_in, chunk_tiles=chunk_tiles)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"e2e macenko512 chunk_tiles={chunk_tiles}: {dt * 1e3:.2f} ms/step, {B * H * W / dt / 1e6:.0f} Mpx/s, "
              f"{host_in.numel() / dt / 1e9:.1f} GB/s each way")


if __name__ == "__main__":
    main()

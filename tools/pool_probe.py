"""Per-rank tile pools of bench.py (seeds 5000 + 64 r) timed on ONE GPU: is the statistics kernel's time data dependent?
usage: python tools/pool_probe.py [tile=512] [pools=8]   (SB_BRACKET_SIGMAS / SB_BRACKET_PAD select the bracket width)"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200 import _native as nv
from stainlib_b200.synth import synth_batch, synth_tile
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pools = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = 1024 * 512 * 512 // (H * W)
n = sb.ExtractiveStainNormalizer("macenko"); n.fit(synth_tile(1, H, W, kind="target"))
lib = nv.load_library(); h, _ = nv.get_handle(0); p = n._params()
M = torch.empty(B, 2, 3, dtype=torch.float64, device="cuda"); C = torch.empty(B, 2, dtype=torch.float64, device="cuda")
for r in range(pools):
    x = torch.from_numpy(synth_batch(5000 + 64 * r, 64, H, W)).repeat(B // 64, 1, 1, 1).contiguous().cuda()
    f = lambda: nv.check(lib.sb_fit(h, nv.ptr(x), B, H, W, ctypes.byref(p), nv.ptr(M), nv.ptr(C), None, nv.stream_ptr(0)))
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    print(f"rank-{r} pool {H}x{W}: statistics kernel {a.elapsed_time(b) / 10:.3f} ms")

"""Stage / cluster-size timing of the fused pipeline (development tool; numbers land in profiles/)."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import stainlib_b200 as sb
from stainlib_b200 import _native as nv
from stainlib_b200.synth import synth_tile, synth_batch

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    lib = nv.load_library(); h, idx = nv.get_handle(0)
    pool = torch.from_numpy(synth_batch(5000, 64, H, H))
    x = pool.repeat(-(-B // 64), 1, 1, 1)[:B].contiguous().cuda()
    out = torch.empty_like(x)
    M = torch.empty(B, 2, 3, dtype=torch.float64, device='cuda'); C = torch.empty(B, 2, dtype=torch.float64, device='cuda')
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    n = sb.ExtractiveStainNormalizer('macenko'); n.fit(synth_tile(1, H, H, kind='target'))
    tgt = torch.as_tensor(np.concatenate([n.stain_matrix_target.reshape(6), n.maxC_target.reshape(2)])).cuda()
    mpx = B * H * H / 1e6
    for S in (1, 2, 4, 8):
        p = nv.default_params(0, cluster_size=S)
        t_ext = timeit(lambda: nv.check(lib.sb_extract(h, nv.ptr(x), B, H, H, ctypes.byref(p), nv.ptr(M), nv.ptr(st), nv.stream_ptr(0))))
        t_fit = timeit(lambda: nv.check(lib.sb_fit(h, nv.ptr(x), B, H, H, ctypes.byref(p), nv.ptr(M), nv.ptr(C), nv.ptr(st), nv.stream_ptr(0))))
        t_nrm = timeit(lambda: nv.check(lib.sb_normalize(h, nv.ptr(x), nv.ptr(out), B, H, H, ctypes.byref(p), nv.ptr(tgt), ctypes.c_void_p(tgt.data_ptr() + 48), None, None, nv.ptr(st), nv.stream_ptr(0))))
        print(f"S={S}: extract {t_ext:.3f} ms  fit {t_fit:.3f} ms  normalize {t_nrm:.3f} ms  -> {mpx / t_nrm:.1f} Gpx/s   (A+B {t_ext:.3f}, C {t_fit - t_ext:.3f}, D {t_nrm - t_fit:.3f})")
    scale = (torch.as_tensor(n.maxC_target, device='cuda') / C).contiguous(); Mt = torch.as_tensor(n.stain_matrix_target, device='cuda').contiguous()
    t = timeit(lambda: nv.check(lib.sb_recombine(h, nv.ptr(x), nv.ptr(out), B, H, H, nv.ptr(M), nv.ptr(scale), nv.ptr(Mt), 0.01, nv.stream_ptr(0))))
    print(f"K4 recombine {t:.3f} ms -> {mpx / t:.1f} Gpx/s, {mpx * 6 / t:.1f} GB/s")
    mask = torch.empty(B, H, H, dtype=torch.uint8, device='cuda')
    t = timeit(lambda: nv.check(lib.sb_tissue_mask(h, nv.ptr(x), B, H, H, 0.8, nv.ptr(mask), None, nv.stream_ptr(0))))
    print(f"mask {t:.3f} ms -> {mpx / t:.1f} Gpx/s")
    pv = nv.default_params(1, dl_iters=30)
    t = timeit(lambda: nv.check(lib.sb_extract(h, nv.ptr(x), B, H, H, ctypes.byref(pv), nv.ptr(M), nv.ptr(st), nv.stream_ptr(0))), n=2)
    print(f"vahadane extract 30 it {t:.3f} ms -> {mpx / t:.1f} Gpx/s ({t / 30:.3f} ms/iter)")

main()

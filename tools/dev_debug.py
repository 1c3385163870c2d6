import ctypes, numpy as np, torch, sys
sys.path.insert(0, '.')
from oracle import stain_oracle as so
from stainlib_b200 import _native as nv
import stainlib_b200 as sb
from stainlib_b200.synth import synth_tile
src = synth_tile(0, 256); tgt = synth_tile(1, 256, kind='target')
o = so.ExtractiveStainNormalizer('macenko'); o.fit(tgt)
Ms = so.macenko_stain_matrix(src); Cs = so.get_concentrations(src, Ms); maxCs = np.percentile(Cs, 99, axis=0)
n = sb.ExtractiveStainNormalizer('macenko'); n.fit(src)
print('src M err', np.abs(n.stain_matrix_target - Ms).max(), 'maxC', n.maxC_target, maxCs, (n.maxC_target - maxCs) / maxCs)
n.fit(tgt)
print('tgt M err', np.abs(n.stain_matrix_target - o.stain_matrix_target).max(), (n.maxC_target - o.maxC_target) / o.maxC_target)
ref = o.transform(src)
got = n.transform(src)
d = got.astype(int) - ref.astype(int)
print('transform diff hist', {k: int((d == k).sum()) for k in np.unique(d)}, 'per channel exact', [(d[..., c] == 0).mean() for c in range(3)])
# K4 with exact params
b = nv.Batch(torch.from_numpy(src).cuda()); out = b.new_like()
scale = (o.maxC_target / maxCs).reshape(1, 2)
dM, dS, dT = (torch.as_tensor(x, dtype=torch.float64).cuda() for x in (Ms[None], scale, o.stain_matrix_target))
nv.check(nv.load_library().sb_recombine(b.handle, nv.ptr(b.dev), nv.ptr(out), 1, 256, 256, nv.ptr(dM), nv.ptr(dS), nv.ptr(dT), 0.01, nv.stream_ptr(b.idx)))
d2 = out[0].cpu().numpy().astype(int) - ref.astype(int)
print('K4 exact params diff hist', {k: int((d2 == k).sum()) for k in np.unique(d2)})
from stainlib_b200.utils.stain_utils import get_concentrations
C = get_concentrations(src, Ms)
print('conc err', np.abs(C - Cs).max(), 'where', np.unravel_index(np.abs(C - Cs).argmax(), C.shape), 'frac>1e-4', (np.abs(C - Cs) > 1e-4).mean())
bad = np.abs(C - Cs).max(axis=1) > 1e-4
print('bad examples', Cs[bad][:5], C[bad][:5])

"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference package from /root/reference (build container) or from the
pip --target copy under baseline/_ref (git-ignored; travels to the GPU box for bench.py's CPU reference arm).

``import stainlib`` fails in this image because ``spams``, ``skimage`` and ``pylab`` are not installed
(``stain_utils.py:3``, ``augmenter.py:5,10``).  This loader injects three stub modules so that the reference's own,
unmodified Python runs:

* ``spams.lasso``   -> ``oracle.stain_oracle.lasso_pos2`` wrapped in a scipy CSC matrix (the reference calls
  ``.toarray()`` on the result, ``stain_utils.py:78``);
* ``spams.trainDL`` -> ``oracle.stain_oracle.train_dl_fullbatch`` (deterministic) -- the reference's real call is
  irreproducible, see the oracle header;
* ``skimage.color`` -> the skimage-0.17.2 restatements in the oracle;
* ``pylab``         -> empty module (only used by plotting helpers).

Everything else (numpy, OpenCV) is the real thing.  Used by ``oracle/gen_golden.py`` to produce ``tests/golden``;
never used on the GPU box (``/root/reference`` does not exist there).
"""
import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


def load_reference(n_iter_traindl=50, root=None):
    import scipy.sparse as sp
    from oracle import stain_oracle as so

    spams = types.ModuleType("spams")

    def lasso(X, D, mode=2, lambda1=0.0, pos=False, **kw):
        assert mode == 2 and pos, "only the call pattern of stain_utils.py:78 is shimmed"
        return sp.csc_matrix(so.lasso_pos2(X, D, lambda1))

    def trainDL(X, K=2, lambda1=0.1, mode=2, modeD=0, posAlpha=True, posD=True, verbose=False, **kw):
        assert K == 2 and mode == 2 and modeD == 0 and posAlpha and posD
        return so.train_dl_fullbatch(X, lam=lambda1, n_iter=n_iter_traindl)

    spams.lasso = lasso
    spams.trainDL = trainDL

    skimage = types.ModuleType("skimage")
    color = types.ModuleType("skimage.color")
    color.rgb2hed = lambda rgb: so.rgb2hed(rgb)
    color.hed2rgb = lambda hed: so.hed2rgb(hed)
    color.rgb2gray = so.rgb2gray
    skimage.color = color
    pylab = types.ModuleType("pylab")

    sys.modules.setdefault("spams", spams)
    sys.modules.setdefault("skimage", skimage)
    sys.modules.setdefault("skimage.color", color)
    sys.modules.setdefault("pylab", pylab)
    root = root or REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in [m for m in sys.modules if m == "stainlib" or m.startswith("stainlib.")]:
        del sys.modules[name]
    ref = importlib.import_module("stainlib")
    importlib.import_module("stainlib.augmentation.augmenter")
    assert ref.__file__.startswith(root), ref.__file__
    return ref

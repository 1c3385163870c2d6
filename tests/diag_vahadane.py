"""Diagnostic (run by hand on the GPU box: python tests/diag_vahadane.py): |GPU - oracle| and |. - fixed point| of the Vahadane stain matrix for a few synthetic tiles and pass counts."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from oracle import stain_oracle as so
from stainlib_b200.synth import synth_tile
for size, seed in [(256, 0), (256, 7), (512, 3), (128, 5), (512, 11), (1024, 2)]:
    I = synth_tile(seed, size)
    M_fix = so.vahadane_stain_matrix(I, solver="fullbatch", n_iter=200)
    for nf in (5, 6, 8):
        M = sb.VahadaneStainExtractor.get_stain_matrix(I, n_iter=nf)
        M_o = so.vahadane_stain_matrix(I, n_iter=nf)
        print(f"{size} seed {seed} nf {nf}: gpu-oracle {np.abs(M - M_o).max():.2e}  gpu-fix {np.abs(M - M_fix).max():.2e}  oracle-fix {np.abs(M_o - M_fix).max():.2e}")

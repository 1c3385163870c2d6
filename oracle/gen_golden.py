"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/reference_golden.npz by running the REAL reference package
(/root/reference, unmodified, loaded through oracle/ref_loader.py) in the build container.

Run:  python -m oracle.gen_golden            (from the repo root; needs /root/reference)

Provenance of each group of keys:
  pure reference (numpy + OpenCV only, nothing shimmed):
      mask/*  macenko_M/*  reinhard/*  lumstd/*  od/*  errors/*  labutil/*
  reference code + shimmed spams.lasso (closed form, see oracle/stain_oracle.py header):
      macenko_norm/*  stain_aug/*
  reference code + shimmed spams.trainDL (deterministic full-batch restatement, 50 iterations):
      vahadane/*
  reference code + shimmed skimage 0.17.2 colour deconvolution:
      hed/*  gray/*
Inputs are stored alongside outputs so the fixtures do not depend on the synthetic generator staying unchanged.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.ref_loader import load_reference  # noqa: E402
from stainlib_b200.synth import synth_tile, edge_case_tiles  # noqa: E402


def main(out_path):
    load_reference(n_iter_traindl=50)
    from stainlib.extraction.macenko_stain_extractor import MacenkoStainExtractor
    from stainlib.extraction.vahadane_stain_extractor import VahadaneStainExtractor
    from stainlib.normalization.normalizer import ExtractiveStainNormalizer, ReinhardStainNormalizer
    from stainlib.augmentation.augmenter import (StainAugmentor, HedLightColorAugmenter, HedLighterColorAugmenter,
                                                 HedStrongColorAugmenter, GrayscaleAugmentor)
    from stainlib.utils import stain_utils as su
    from stainlib.utils.excepts import TissueMaskException

    G = {}
    cases = {
        "c1_256": (synth_tile(0, 256), synth_tile(1, 256, kind="target")),          # BASELINE config 1
        "s_64": (synth_tile(10, 64), synth_tile(11, 64, kind="target")),
        "ragged_96x80": (synth_tile(20, 96, 80), synth_tile(21, 72, 100, kind="target")),
        "s_128": (synth_tile(30, 128), synth_tile(31, 128, kind="target")),
        "odd_67x53": (synth_tile(40, 67, 53), synth_tile(41, 67, 53, kind="target")),
    }
    edges = edge_case_tiles(64, 64)
    for name, (src, tgt) in cases.items():
        G[f"in/{name}/src"] = src
        G[f"in/{name}/tgt"] = tgt
        # ---- pure reference
        G[f"mask/{name}"] = su.LuminosityThresholdTissueLocator.get_tissue_mask(src)
        G[f"mask075/{name}"] = su.LuminosityThresholdTissueLocator.get_tissue_mask(src, luminosity_threshold=0.75)
        if name == "s_64":
            G[f"od/{name}"] = su.convert_RGB_to_OD(src)
        G[f"macenko_M/{name}/src"] = MacenkoStainExtractor.get_stain_matrix(src)
        G[f"macenko_M/{name}/tgt"] = MacenkoStainExtractor.get_stain_matrix(tgt)
        G[f"macenko_M95/{name}/src"] = MacenkoStainExtractor.get_stain_matrix(src, angular_percentile=95)
        r = ReinhardStainNormalizer()
        r.fit(tgt)
        G[f"reinhard/{name}/means"] = np.array(r.target_means).reshape(3)
        G[f"reinhard/{name}/stds"] = np.array(r.target_stds).reshape(3)
        G[f"reinhard/{name}/out"] = r.transform(src)
        G[f"reinhard/{name}/out_masked"] = r.transform(src, mask_background=True)
        G[f"lumstd/{name}"] = su.LuminosityStandardizer.standardize(src)
        if name in ("s_64", "odd_67x53"):
            # the exported pieces of the Reinhard path (stain_utils.py:146-194)
            G[f"labutil/{name}/bright"] = su.standardize_brightness(src)
            I1, I2, I3 = su.lab_split(src)
            G[f"labutil/{name}/I1"], G[f"labutil/{name}/I2"], G[f"labutil/{name}/I3"] = I1.copy(), I2.copy(), I3.copy()
            m, sd = su.get_mean_std(src)
            G[f"labutil/{name}/means"] = np.array(m).reshape(3)
            G[f"labutil/{name}/stds"] = np.array(sd).reshape(3)
            # merge_back of transformed planes, float32 and float64 (it scales its arguments in place: pass copies)
            J = [(I1 * 0.9 + 3.0), (I2 * 1.1 - 2.0), (I3 * 0.8 + 1.5)]
            G[f"labutil/{name}/merge_f32"] = su.merge_back(*[x.astype(np.float32).copy() for x in J])
            G[f"labutil/{name}/merge_f64"] = su.merge_back(*[x.astype(np.float64).copy() for x in J])
        # ---- reference + lasso shim
        n = ExtractiveStainNormalizer("macenko")
        n.fit(tgt)
        G[f"macenko_norm/{name}/M_target"] = n.stain_matrix_target
        G[f"macenko_norm/{name}/maxC_target"] = n.maxC_target
        G[f"macenko_norm/{name}/out"] = n.transform(src)
        if name in ("s_64", "ragged_96x80"):
            G[f"macenko_norm/{name}/conc_src"] = su.get_concentrations(src, G[f"macenko_M/{name}/src"])
        a = StainAugmentor("macenko")
        a.fit(src)
        np.random.seed(1234)
        G[f"stain_aug/{name}/pop0"] = a.pop()
        G[f"stain_aug/{name}/pop1"] = a.pop()
        a = StainAugmentor("macenko", sigma1=0.4, sigma2=0.3, augment_background=True)
        a.fit(src)
        np.random.seed(99)
        G[f"stain_aug/{name}/pop_bg"] = a.pop()
        # ---- reference + trainDL shim
        if name in ("s_64", "ragged_96x80", "s_128"):
            G[f"vahadane/{name}/M_src"] = VahadaneStainExtractor.get_stain_matrix(src)
            v = ExtractiveStainNormalizer("vahadane")
            v.fit(tgt)
            G[f"vahadane/{name}/M_target"] = v.stain_matrix_target
            G[f"vahadane/{name}/maxC_target"] = v.maxC_target
            G[f"vahadane/{name}/out"] = v.transform(src)
        # ---- reference + skimage shim
        for cls, tag in ((HedLighterColorAugmenter, "lighter"), (HedLightColorAugmenter, "light"),
                         (HedStrongColorAugmenter, "strong")):
            h = cls()
            G[f"hed/{name}/{tag}/default"] = h.transform(src)   # before randomize(): sigma = beta = -thresh
            np.random.seed(7)
            h.randomize()
            G[f"hed/{name}/{tag}/sigmas"] = np.array(h._sigmas)
            G[f"hed/{name}/{tag}/biases"] = np.array(h._biases)
            G[f"hed/{name}/{tag}/out"] = h.transform(src)
        g = GrayscaleAugmentor()
        g.fit(src)
        np.random.seed(5)
        G[f"gray/{name}/pop"] = g.pop()

    # edge cases: store inputs, outputs where defined, and which exception the reference raises
    for name, I in edges.items():
        G[f"in/edge_{name}"] = I
        for what, fn in (("mask", lambda x: su.LuminosityThresholdTissueLocator.get_tissue_mask(x)),
                         ("macenko_M", lambda x: MacenkoStainExtractor.get_stain_matrix(x))):
            try:
                G[f"{what}/edge_{name}"] = fn(I)
                G[f"errors/{what}/edge_{name}"] = np.array("")
            except TissueMaskException as e:
                G[f"errors/{what}/edge_{name}"] = np.array("TissueMaskException:" + str(e))
            except Exception as e:  # e.g. LinAlgError for a single tissue pixel
                G[f"errors/{what}/edge_{name}"] = np.array(type(e).__name__ + ":" + str(e))
        if G[f"errors/macenko_M/edge_{name}"] == "":
            n = ExtractiveStainNormalizer("macenko")
            n.fit(cases["s_64"][1])
            G[f"macenko_norm/edge_{name}/out"] = n.transform(I)
        r = ReinhardStainNormalizer()
        r.fit(cases["s_64"][1])
        if name not in ("all_white",):
            G[f"reinhard/edge_{name}/out"] = r.transform(I)
    # hed cutoff: nearly-white patch is returned unchanged
    h = HedLightColorAugmenter()
    G["hed/edge_all_white/light/default"] = h.transform(edges["all_white"])

    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **G)
    print(f"wrote {out_path}: {len(G)} arrays, {os.path.getsize(out_path) / 1e6:.2f} MB")
    for k in sorted(G):
        if k.startswith("errors/"):
            print(k, "->", str(G[k]))


if __name__ == "__main__":
    main(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_golden.npz"))

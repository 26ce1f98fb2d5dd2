"""Calibration of the image gates (parity layer 3) by the reference itself.

    python tests/golden/make_selfcal.py          (build container only: needs oracle/_ref)

For every (scene, integrator) image case of tests/test_gpu_render.py whose scene is deterministic
(everything but the randomly generated scenes 1 and 9) the UNMODIFIED reference renders

  * four more independent images B with the fixture's sample count: frac_<scene>_<integrator> =
    [mean, standard deviation] over them of the fraction of pixel channels of B within 3 sigma of the
    fixture's image A (sigma per pixel from A's own sample variance, parity.unpooled_3sigma_fraction) —
    what the estimator returns when both sides ARE the reference, at EQUAL sample counts;
  * a high-sample image (64x the fixture's samples): mean_<scene>_<integrator> = whole-image mean
    r, g, b and its standard error r, g, b — the target of the 1 % mean gate.
Only scenes a second reference instance reproduces (same primitives up to BVH order, same Perlin
tables, no gated spheres) can be calibrated; the random ones keep the pooled gates.

Written to tests/golden/selfcal.npz (committed).  The GPU tests gate the un-pooled 3-sigma fraction
at the reference-vs-itself value minus a stated margin, and the mean at 1 % of the high-sample mean.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402
import parity  # noqa: E402
from oracle import refbind  # noqa: E402

CASES = [(7, 0), (7, 1), (21, 3), (21, 4), (23, 2), (23, 3), (23, 4), (19, 3), (19, 4), (26, 4), (24, 4), (15, 3),
         (17, 4), (18, 3), (8, 1)] + [(c[0], c[1]) for c in make_golden.CATALOGUE]


N_SELF = 4       # independent reference re-renders behind every calibrated 3-sigma fraction
HI_FACTOR = 64   # samples of the high-sample mean, in units of the fixture's sample count


def same_scene(blob_a, blob_b):
    """The same primitives up to the order the reference's random-axis BVH build leaves them in."""
    a, b = refbind.abi.parse_blob(blob_a), refbind.abi.parse_blob(blob_b)
    if len(a["gates"]) or len(b["gates"]) or len(a["prims"]) != len(b["prims"]):
        return False

    def rows(t):
        r = np.concatenate([t["prims"]["type"][:, None].astype(np.float64), t["prims"]["d"]], axis=1)
        return r[np.lexsort(r.T[::-1])]
    return np.array_equal(rows(a), rows(b)) and all(len(a[k]) == len(b[k]) for k in ("materials", "textures", "lights")) \
        and a["perlins"].tobytes() == b["perlins"].tobytes()     # (Perlin tables are drawn per instance)


def sid_needs_assets(out):
    return any(f"frac_{c[0]}_{c[1]}" not in out for c in make_golden.CATALOGUE)


def main():
    path = os.path.join(HERE, "selfcal.npz")
    out = dict(np.load(path)) if os.path.exists(path) and "--all" not in sys.argv else {}   # (only the missing cases)
    with tempfile.TemporaryDirectory() as tmp:
        make_golden.write_hdr(os.path.join(tmp, "sky.hdr"), make_golden.synthetic_sky(64, 32, 1))
        make_golden.write_hdr(os.path.join(tmp, "rnl_probe.hdr"), make_golden.synthetic_sky(32, 32, 2))
        if sid_needs_assets(out):
            make_golden.write_assets(tmp)
        os.chdir(tmp)
        for sid, integ in CASES:
            if f"frac_{sid}_{integ}" in out or sid == 24:   # (24: the fixture is the missing-.hdr fallback)
                continue
            g = np.load(os.path.join(HERE, f"scene{sid:02d}.npz"))
            if not same_scene(refbind.RefScene(sid).blob(), g["blob"].tobytes()):
                # a randomly generated scene, or one whose random tree decides what a gated sphere can be hit
                # through: a second instance of the reference is a different scene, no calibration possible
                print(f"scene {sid}: not reproducible, skipped", flush=True)
                continue
            a_sum, a_sumsq, spp = g[f"img_{integ}_sum"], g[f"img_{integ}_sumsq"], int(g[f"img_{integ}_spp"][0])
            h, w, _ = a_sum.shape
            s = refbind.RefScene(sid)
            fr = []
            for _ in range(N_SELF):
                b_sum, b_sumsq, _ = s.render_linear(integ, w, h, spp)
                fr.append(parity.unpooled_3sigma_fraction(a_sum, a_sumsq, spp, b_sum / spp, spp))
            hi = HI_FACTOR * spp
            S, S2, _ = s.render_linear(integ, w, h, hi)
            mean = S / hi
            var = np.maximum(S2 / hi - mean ** 2, 0) / hi
            se = np.sqrt(var.sum(axis=(0, 1))) / (w * h)
            out[f"frac_{sid}_{integ}"] = np.array([np.mean(fr), np.std(fr, ddof=1)])
            out[f"mean_{sid}_{integ}"] = np.concatenate([mean.mean(axis=(0, 1)), se])
            print(f"scene {sid} int {integ}: ref-vs-ref within 3 sigma (un-pooled, {N_SELF} re-renders) {np.mean(fr):.4f} +- {np.std(fr, ddof=1):.4f}; "
                  f"mean {mean.mean(axis=(0, 1))} rel se {se / mean.mean(axis=(0, 1))}", flush=True)
        os.chdir(ROOT)
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()

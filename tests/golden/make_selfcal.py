"""Calibration of the image gates (parity layer 3) by the reference itself.

    python tests/golden/make_selfcal.py          (build container only: needs oracle/_ref)

For every (scene, integrator) image case of tests/test_gpu_render.py whose scene is deterministic
(everything but the randomly generated scenes 1 and 9) the UNMODIFIED reference renders

  * a second, independent image B with the fixture's sample count: frac_<scene>_<integrator> =
    [un-pooled, pooled] fraction of pixel channels of B within 3 sigma of the fixture's image A —
    what the estimator of tests/parity.py returns when both sides ARE the reference;
  * a high-sample image (16x the fixture's samples): mean_<scene>_<integrator> = whole-image mean
    r, g, b and its standard error r, g, b — the target of the 1 % mean gate.

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
         (17, 4), (18, 3), (8, 1)]


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        make_golden.write_hdr(os.path.join(tmp, "sky.hdr"), make_golden.synthetic_sky(64, 32, 1))
        make_golden.write_hdr(os.path.join(tmp, "rnl_probe.hdr"), make_golden.synthetic_sky(32, 32, 2))
        os.chdir(tmp)
        for sid, integ in CASES:
            g = np.load(os.path.join(HERE, f"scene{sid:02d}.npz"))
            a_sum, a_sumsq, spp = g[f"img_{integ}_sum"], g[f"img_{integ}_sumsq"], int(g[f"img_{integ}_spp"][0])
            h, w, _ = a_sum.shape
            s = refbind.RefScene(sid)
            b_sum, b_sumsq, _ = s.render_linear(integ, w, h, spp)
            rep = parity.image_report(a_sum, a_sumsq, spp, (b_sum / spp)[None], gpu_spp=spp)
            hi = 16 * spp
            S, S2, _ = s.render_linear(integ, w, h, hi)
            mean = S / hi
            var = np.maximum(S2 / hi - mean ** 2, 0) / hi
            se = np.sqrt(var.sum(axis=(0, 1))) / (w * h)
            out[f"frac_{sid}_{integ}"] = np.array([rep["frac_within_3sigma_unpooled"], rep["frac_within_3sigma"]])
            out[f"mean_{sid}_{integ}"] = np.concatenate([mean.mean(axis=(0, 1)), se])
            print(f"scene {sid} int {integ}: ref-vs-ref within 3 sigma un-pooled {rep['frac_within_3sigma_unpooled']:.4f} "
                  f"pooled {rep['frac_within_3sigma']:.4f}; mean {mean.mean(axis=(0, 1))} rel se {se / mean.mean(axis=(0, 1))}", flush=True)
        os.chdir(ROOT)
    np.savez_compressed(os.path.join(HERE, "selfcal.npz"), **out)


if __name__ == "__main__":
    main()

"""Comparison helpers shared by the CPU (hostcheck / oracle restatement) and GPU parity
tests.  Every tolerance used by a test is defined here, next to its justification."""
import numpy as np

# ---- tolerances (north_star's three correctness layers) ----------------------------------------
# Layer 1: closest-hit primitive id and t.
#   fp64 validation kernels: bit-exact (0 differences) on non-stochastic primitives.
#   fp32 production kernels: the same primitive on >= 99.99 % of the queries.
FP32_MIN_AGREEMENT = 0.9999
# Layer 2: BSDF eval/pdf, Light::sample/pdf/Le and texture values from the fp64 kernels:
# |got - ref| <= 1e-5 * |ref| + 1e-12 (libm of CUDA vs glibc differs in the last ulps of
# sin/cos/acos/atan2/pow, nothing else).
VALUE_RTOL = 1e-5
VALUE_ATOL = 1e-12
# The fp32 production evaluation of the same functions is reported against a looser bound
# (GGX at roughness 0.01 has condition number ~1e4 near the specular peak).
FP32_VALUE_RTOL = 5e-3
FP32_VALUE_ATOL = 1e-5
# Layer 3: images (the RNG streams differ, so the comparison is statistical).  Two sets of gates:
# (a) every case, image_gates() — the pooled estimator:
#   * whole-image mean within 1 % of the reference's (or within 4 standard errors of the
#     difference where the noise of the two estimates is itself above 1 %);
#   * per-pixel |GPU - reference| <= 3 sigma for >= 97 % of the pixel channels, sigma from the
#     reference's per-pixel sample variance and the spread of K independent GPU renders, both
#     pooled over 5x5 pixels (the reference against ITSELF gives 97.9-99.4 %; a Gaussian with known
#     sigma would give 99.73 %: path-tracing noise is heavy-tailed and sigma is estimated);
#   * RMSE <= 1.6x the RMSE expected from the two noise levels alone (reference vs itself: 0.9-1.5x).
# (b) every case a second reference instance reproduces (selfcal), strict_image_gates() — the contract's
#   own: mean within 1 % of a 64x-sample reference mean with no escape, and the UN-POOLED 3-sigma
#   fraction at equal sample counts held to the reference-vs-itself value of that very scene.
IMAGE_MEAN_RTOL = 0.01
IMAGE_MEAN_NSIGMA = 4.0
IMAGE_3SIGMA_MIN_FRACTION = 0.97
IMAGE_RMSE_FACTOR = 1.6

PRIM_MEDIUM = 5


def is_medium(tables, prim_ids):
    ptype = tables["prims"]["type"]
    prim_ids = np.asarray(prim_ids)
    return np.where(prim_ids >= 0, ptype[np.maximum(prim_ids, 0)] == PRIM_MEDIUM, False)


def deterministic_mask(tables, ref_hits, got_hits):
    """Queries whose answer does not depend on constant_medium's random free-flight draw
    (constant_medium.h:85): neither side reports a medium.  Media only ever ADD candidate
    hits, so when both answers are solid primitives they must be the same primitive."""
    return ~is_medium(tables, ref_hits["prim"]) & ~is_medium(tables, got_hits["prim"])


def gated_mask(tables, *hit_sets):
    """Queries none of whose answers is a gated (negative-radius) sphere (rtb200_scene.h rtb_gate).  The
    fp64 validation kernels reproduce the reference's reach of such a sphere exactly
    (trace_gated_exact, csrc/rtb_geom.cuh); the fp32 production traversals test its gate box against
    the interval they hold at that moment of THEIR visiting order, which can differ from the
    reference's on rays that meet the sphere before the box (2 of 1,200 recorded rays on scene 34)."""
    gated = set(int(i) for i in tables["gates"]["prim"]) if len(tables.get("gates", ())) else set()
    ok = np.ones(len(hit_sets[0]), bool)
    for h in hit_sets:
        ok &= ~np.isin(h["prim"], list(gated))
    return ok


def trace_mismatches(ref_hits, got_hits, mask, fields=("prim", "t", "p", "normal", "front_face", "material")):
    """Number of masked queries on which any of `fields` differs bit for bit."""
    bad = np.zeros(len(ref_hits), bool)
    for f in fields:
        a, b = ref_hits[f], got_hits[f]
        d = a != b
        if d.ndim > 1:
            d = d.any(axis=1)
        bad |= d
    # misses carry zeros on both sides
    return int((bad & mask).sum())


def to_segment_form(rays):
    """Shadow queries (finite t_max) in the form the production connect stage issues them:
    direction = light_point - origin (not normalised), t in [t_min/dist, 1 - 0.001/dist].
    Same geometric segment as the reference's (wi, [0.001, dist - 0.001])
    (direct_light_integrator.h:115-118), but well conditioned in fp32."""
    out = rays.copy()
    sh = np.isfinite(rays["t_max"])
    dist = rays["t_max"][sh] + 0.001
    # the production path holds the shading point and the light point in fp32 and forms
    # their difference there; mirror that so both ends of the segment are what fp32 sees
    o32 = rays["o"][sh].astype(np.float32).astype(np.float64)
    q32 = (rays["o"][sh] + rays["d"][sh] * dist[:, None]).astype(np.float32).astype(np.float64)
    out["o"][sh] = o32
    out["d"][sh] = q32 - o32
    out["t_min"][sh] = rays["t_min"][sh] / dist
    out["t_max"][sh] = 1.0 - 0.001 / dist
    return out


def values_close(got, ref, rtol, atol):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    both_inf = np.isinf(got) & np.isinf(ref) & (np.sign(got) == np.sign(ref))
    return both_inf | (np.abs(got - ref) <= rtol * np.abs(ref) + atol)


def box_filter(a, r=2):
    """Mean over a (2r+1)^2 window (edge-clamped), per channel."""
    a = np.asarray(a, np.float64)
    pad = np.pad(a, ((r, r), (r, r), (0, 0)), mode="edge")
    out = np.zeros_like(a)
    n = 2 * r + 1
    for dy in range(n):
        for dx in range(n):
            out += pad[dy:dy + a.shape[0], dx:dx + a.shape[1]]
    return out / (n * n)


def image_report(ref_sum, ref_sumsq, ref_spp, gpu_means):
    """ref_*: per-pixel sum / sum of squares of linear Li over ref_spp samples.
    gpu_means: (K, H, W, 3) means of K independent GPU renders.  The pooled estimator: sigma from the
    reference's sample variance and the spread of the K GPU renders, both pooled over 5x5 pixels."""
    ref_sum = np.asarray(ref_sum, np.float64)
    ref_mean = ref_sum / ref_spp
    ref_var = np.maximum(np.asarray(ref_sumsq, np.float64) / ref_spp - ref_mean ** 2, 0) / ref_spp
    gpu_means = np.asarray(gpu_means, np.float64)
    k = gpu_means.shape[0]
    gpu_mean = gpu_means.mean(axis=0)
    gpu_var = gpu_means.var(axis=0, ddof=1) / k
    var = box_filter(ref_var) + box_filter(gpu_var)
    sigma = np.sqrt(var + 1e-20)
    z = (gpu_mean - ref_mean) / sigma
    lit = sigma > 1e-9
    se = np.sqrt((ref_var + gpu_var).sum(axis=(0, 1))) / (ref_var.shape[0] * ref_var.shape[1])
    return {
        "mean_se": se,
        "ref_mean": ref_mean.mean(axis=(0, 1)),
        "gpu_mean": gpu_mean.mean(axis=(0, 1)),
        "mean_rel_err": np.abs(gpu_mean.mean(axis=(0, 1)) - ref_mean.mean(axis=(0, 1))) /
        np.maximum(ref_mean.mean(axis=(0, 1)), 1e-12),
        "frac_within_3sigma": float((np.abs(z[lit]) <= 3).mean()) if lit.any() else 1.0,
        # RMSE over the displayable range: the reference clamps to [0,1] before it stores a
        # pixel (renderer.h:136-139); a single unconverged specular highlight (values >> 1,
        # enormous variance) would otherwise decide the whole-image RMSE on its own
        "rmse": float(np.sqrt(np.mean((np.clip(gpu_mean, 0, 1) - np.clip(ref_mean, 0, 1)) ** 2))),
        "expected_rmse": float(np.sqrt(np.mean(np.where((gpu_mean > 1) | (ref_mean > 1), 0.0, var)))),
    }


def unpooled_3sigma_fraction(ref_sum, ref_sumsq, ref_spp, other_mean, other_spp):
    """Fraction of pixel channels on which |other - reference| <= 3 sigma, sigma per pixel from the
    reference's own sample variance s^2 alone: sigma^2 = s^2 (1 / ref_spp + 1 / other_spp).  No smoothing.
    Path-tracing noise is heavy-tailed and s^2 misses the fireflies the reference did not catch, so the
    value a CORRECT second renderer reaches is well below 99.7 % and depends on the scene and on both sample
    counts: it is only meaningful next to the reference-vs-itself value at the same counts (selfcal)."""
    ref_mean = np.asarray(ref_sum, np.float64) / ref_spp
    s2 = np.maximum(np.asarray(ref_sumsq, np.float64) / ref_spp - ref_mean ** 2, 0)
    sig = np.sqrt(s2 * (1.0 / ref_spp + 1.0 / other_spp) + 1e-20)
    z = (np.asarray(other_mean, np.float64) - ref_mean) / sig
    lit = sig > 1e-9
    return float((np.abs(z[lit]) <= 3).mean()) if lit.any() else 1.0


_selfcal = None


def selfcal(sid, integrator):
    """Reference-vs-itself calibration of this image case (tests/golden/make_selfcal.py), or None:
    {"frac", "frac_sd": mean / standard deviation of the un-pooled 3-sigma fraction of four reference
    re-renders at the fixture's sample count, "mean": high-sample mean rgb, "mean_se": its standard error}."""
    global _selfcal
    if _selfcal is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selfcal.npz")
        _selfcal = dict(np.load(path)) if os.path.exists(path) else {}
    kf, km = f"frac_{sid}_{integrator}", f"mean_{sid}_{integrator}"
    if kf not in _selfcal:
        return None
    return {"frac": float(_selfcal[kf][0]), "frac_sd": float(_selfcal[kf][1]), "mean": _selfcal[km][:3],
            "mean_se": _selfcal[km][3:]}


# The contract's layer-3 gates (north_star / SURVEY 8d(3)) for the calibrated cases:
#   * whole-image mean of a high-sample render within 1 % of the reference's high-sample mean — no escape;
#   * un-pooled 3-sigma fraction of renders at the reference's sample count, averaged over K of them, at
#     least the reference-vs-itself value minus a margin: 1 %, or 3 standard errors of the difference of
#     the two averaged fractions where the scene's fraction is itself noisier than that.
IMAGE_3SIGMA_MARGIN = 0.01


def strict_image_gates(ref_sum, ref_sumsq, ref_spp, equal_spp_means, hi_mean, cal):
    """equal_spp_means: (K, H, W, 3) means of K independent renders of ref_spp samples each; hi_mean: rgb
    whole-image mean of a high-sample render; cal: selfcal() of the case.  Returns failure strings."""
    bad = []
    fr = [unpooled_3sigma_fraction(ref_sum, ref_sumsq, ref_spp, m, ref_spp) for m in equal_spp_means]
    margin = max(IMAGE_3SIGMA_MARGIN, 3.0 * cal["frac_sd"] * np.sqrt(1.0 / 4 + 1.0 / len(fr)))
    if np.mean(fr) < cal["frac"] - margin:
        bad.append(f"un-pooled 3-sigma fraction {np.mean(fr):.4f} (of {len(fr)} renders) < reference vs itself "
                   f"{cal['frac']:.4f} - {margin:.4f}")
    d = np.abs(np.asarray(hi_mean) - cal["mean"])
    if not (d <= IMAGE_MEAN_RTOL * cal["mean"]).all():
        bad.append(f"whole-image mean {np.asarray(hi_mean)} vs reference high-sample mean {cal['mean']} (rel {d / cal['mean']}, "
                   f"gate {IMAGE_MEAN_RTOL})")
    return bad


def image_gates(rep, calibrated=False):
    """The pooled layer-3 gates on an image_report() (every case; the only ones for the randomly generated
    scenes, which no second reference instance reproduces); returns a list of failure strings.
    calibrated: the case also goes through strict_image_gates(), whose per-scene 3-sigma gate replaces
    the fixed 97 % one (scene 31's heavy-tailed caustics put the reference itself at 96.9-97.5 % there)."""
    bad = []
    d = np.abs(rep["gpu_mean"] - rep["ref_mean"])
    ok = (d <= IMAGE_MEAN_RTOL * rep["ref_mean"]) | (d <= IMAGE_MEAN_NSIGMA * rep["mean_se"])
    if not ok.all():
        bad.append(f"whole-image mean: gpu {rep['gpu_mean']} vs ref {rep['ref_mean']} (se {rep['mean_se']})")
    if not calibrated and rep["frac_within_3sigma"] < IMAGE_3SIGMA_MIN_FRACTION:
        bad.append(f"only {rep['frac_within_3sigma']:.4f} of pixel channels within 3 sigma")
    if rep["rmse"] > IMAGE_RMSE_FACTOR * rep["expected_rmse"] + 1e-6:
        bad.append(f"rmse {rep['rmse']:.5f} > {IMAGE_RMSE_FACTOR} x expected {rep['expected_rmse']:.5f}")
    return bad

"""CPU check of the arithmetic of the half-precision box test (csrc/trace.cuh cw_node_test_h2): a numpy emulation of the device
code, operation by operation (fp32 set-up, round-to-nearest conversions to half, fused half multiply-adds with ONE rounding,
exact half min / max / compare), against the exact geometric answer in float64.

What is pinned here is the derivation of the error bound (kH, kAbs on top of kSlabErr): whenever the ray really passes through
a child box within [0, tMax], the emulated test must report a hit - for ordinary rays, for rays far from the node (large
|origin - o| / grid unit), for anisotropic node exponents, and for directions far beyond the 2^11 component ratio at which the
device hands over to the fp32 test (the hand-over is a performance matter: the half test must stay conservative everywhere).
The device code itself is pinned by the hit-id parity tests (tests/test_gpu_trace.py); this file needs no GPU.
Replaces Bounds3::IntersectP (Core/Geometry.h:1437-1468) on the device; the reference's own test is exact-ish fp32 without any
widening, ours may only err towards "hit"."""
import numpy as np

F = np.float32
K_SLAB_ERR = F(5e-7)
K_H = F(2.3) * F(4.8828125e-4)
K_ABS = F(1.1920929e-7)
K_Q = F(255.0) * F(5.9604645e-8)


def _fma32(a, b, c):
    """fp32 fused multiply-add: exact product and sum in float64, one rounding (double rounding is possible only when the exact
    result needs more than 53 bits: irrelevant for a pivot and for slack terms)"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _fma16(q, a, b):
    """HFMA2 lane: half inputs, exact product + sum (float64 holds both exactly: 11 x 11 significand bits, exponent span < 53),
    ONE rounding to half; +-inf offsets propagate, NaN cannot arise from finite q, a"""
    with np.errstate(over="ignore", invalid="ignore"):
        return (q.astype(np.float64) * a.astype(np.float64) + b.astype(np.float64)).astype(np.float16)


def _pow2(e_biased):
    return (e_biased.astype(np.uint32) << np.uint32(23)).view(F)


def h2_box_test(n0, ebyte, qlo, qhi, o, idir, tmax):
    """n0 [N,3] f32, ebyte [N,3] biased exponent bytes, qlo/qhi [N,3] uint8 (ONE child per case), o [N,3] f32, idir [N,3] f32
    (reciprocal direction as trav_init leaves it), tmax [N] f32.  Returns hit [N] bool - the value of one bit of HSET2.LE."""
    a = _pow2(ebyte) * idir                       # exact power-of-two scaling
    b = (n0 - o).astype(F) * idir
    far_plane = _fma32(np.full_like(a, 255.0), a, b)
    tref = np.maximum(np.minimum(b, far_plane).max(axis=1), F(0)).astype(F)
    amax = np.abs(a).max(axis=1).astype(F)
    Eb = amax.view(np.uint32) & np.uint32(0x7F800000)
    S = (np.uint32(0x86000000) - Eb).view(F)
    sg = (np.uint32(0x7A000000) - Eb).view(F)
    tsg = (tref * sg).astype(F)
    sa = (a * S[:, None]).astype(F)
    sb = _fma32(b, np.broadcast_to(sg[:, None], b.shape), np.broadcast_to(-tsg[:, None], b.shape))
    c32 = (K_SLAB_ERR * sg).astype(F)[:, None]
    ca = F(K_H * K_Q)
    inner = _fma32(np.full_like(sb, K_H), np.abs(sb), np.full_like(sb, K_ABS))
    inner = _fma32(np.full_like(sa, ca), np.abs(sa), inner)
    err = _fma32(np.broadcast_to(c32, a.shape), _fma32(np.full_like(a, 512.0), np.abs(a), np.abs(b)), inner)
    z0 = _fma32(-tsg, np.full_like(tsg, F(1.0) + K_H), np.full_like(tsg, -K_ABS))
    with np.errstate(invalid="ignore", over="ignore"):
        tm0 = _fma32(tmax, sg, -tsg)
        tm = (_fma32(np.full_like(tm0, K_H), np.abs(tm0), tm0) + K_ABS).astype(F)
        hA = sa.astype(np.float16)
        hN = (sb - err).astype(F).astype(np.float16)
        hF = (sb + err).astype(F).astype(np.float16)
        hz0, htm = z0.astype(np.float16), tm.astype(np.float16)
    pos = idir >= 0
    qn = np.where(pos, qlo, qhi).astype(np.float64) * 2.0 ** -24   # byte as a half denormal
    qf = np.where(pos, qhi, qlo).astype(np.float64) * 2.0 ** -24
    tn = _fma16(qn.astype(np.float16), hA, hN)
    tf = _fma16(qf.astype(np.float16), hA, hF)
    assert not np.isnan(tn.astype(np.float64)).any() and not np.isnan(tf.astype(np.float64)).any()
    lo = np.maximum(tn.max(axis=1), hz0)
    hi = np.minimum(tf.min(axis=1), htm)
    return lo <= hi


def exact_box_test(n0, ebyte, qlo, qhi, o, d, tmax):
    """The ray o + t d, t in [0, tmax], against the closed box n0 + 2^e [qlo, qhi] in float64 (exact inputs; the divisions carry
    1e-16, nothing next to the fp32 quantities above)."""
    scale = np.ldexp(1.0, ebyte.astype(np.int64) - 127)
    blo = n0.astype(np.float64) + qlo * scale
    bhi = n0.astype(np.float64) + qhi * scale
    o64, d64 = o.astype(np.float64), d.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (blo - o64) / d64
        t2 = (bhi - o64) / d64
    par = d64 == 0
    inside = (o64 >= blo) & (o64 <= bhi)
    near = np.where(par, np.where(inside, -np.inf, np.inf), np.minimum(t1, t2))   # a zero component: inside the slab for all t, or never
    far = np.where(par, np.where(inside, np.inf, -np.inf), np.maximum(t1, t2))
    tn = np.maximum(near.max(axis=1), 0.0)
    tf = np.minimum(far.min(axis=1), tmax.astype(np.float64))
    return tn <= tf


def _recip(d, rng):
    """trav_init's safe_rcp: |1/d| capped at 1e24, rcp.approx (<= 1 ulp off the rounded reciprocal: emulated by a random ulp step)"""
    mag = np.maximum(np.abs(d), F(1e-24)).astype(F)
    r = (F(1) / mag).astype(F)
    step = rng.integers(-1, 2, r.shape)
    r = (r.view(np.int32) + step.astype(np.int32)).view(F)
    return np.where(d < 0, -r, r).astype(F)


def _cases(rng, n, far, ratio_log2, aniso):
    """n random (node, child, ray) cases; rays are aimed at a point inside or just outside the child box, so about half of them
    really hit.  far: distance of the ray origin in node extents; ratio_log2: the smallest direction component is up to that many
    powers of two below the largest; aniso: spread of the three node exponents."""
    e = rng.integers(-14, 3, (n, 1)) + rng.integers(0, aniso + 1, (n, 3))
    ebyte = (e + 127).astype(np.int64)
    scale = np.ldexp(1.0, e)
    n0 = (rng.uniform(-50, 50, (n, 3))).astype(F)
    a = rng.integers(0, 256, (n, 3))
    b = rng.integers(0, 256, (n, 3))
    qlo, qhi = np.minimum(a, b), np.maximum(a, b)
    # target: inside the child (or up to 3 grid units outside on one axis)
    tgt_q = qlo + rng.uniform(0, 1, (n, 3)) * (qhi - qlo)
    k = rng.integers(0, 3, n)
    off = rng.uniform(-3, 3, n) * (rng.random(n) < 0.5)
    r = np.arange(n)
    tgt_q[r, k] = np.where(off > 0, qhi[r, k] + off, np.where(off < 0, qlo[r, k] + off, tgt_q[r, k]))
    tgt = n0.astype(np.float64) + tgt_q * scale
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    small = rng.integers(0, 3, n)
    shrink = np.exp2(-rng.uniform(0, ratio_log2, n))
    d[np.arange(n), small] = np.abs(d).max(axis=1) * shrink * rng.choice([-1.0, 1.0], n)
    zero = rng.random(n) < 0.02   # exactly axis-parallel in one component
    d[zero, small[zero]] = 0.0
    dist = np.exp(rng.uniform(np.log(0.01), np.log(far), n)) * 255 * scale.max(axis=1)
    o = (tgt - d * dist[:, None]).astype(F)
    d = d.astype(F)
    inside = rng.random(n) < 0.15   # origins inside the node
    o[inside] = (n0[inside].astype(np.float64) + rng.uniform(0, 255, (inside.sum(), 3)) * scale[inside]).astype(F)
    tmax = np.where(rng.random(n) < 0.5, np.inf, dist * rng.uniform(0.5, 1.5, n)).astype(F)
    return n0, ebyte, qlo, qhi, o, d, tmax


def _run(seed, n, far, ratio_log2, aniso):
    rng = np.random.default_rng(seed)
    n0, ebyte, qlo, qhi, o, d, tmax = _cases(rng, n, far, ratio_log2, aniso)
    got = h2_box_test(n0, ebyte, qlo, qhi, o, _recip(d, rng), tmax)
    want = exact_box_test(n0, ebyte, qlo, qhi, o, d, tmax)
    lost = want & ~got
    extra = got & ~want
    return want.mean(), lost.sum(), extra.sum() / max((~want).sum(), 1)


def test_half_box_test_never_loses_a_box():
    for seed, (far, ratio, aniso) in enumerate([(3.0, 4, 0), (30.0, 10, 2), (1e4, 11, 4), (1e6, 11, 8)]):
        hit_frac, lost, fp = _run(seed, 400_000, far, ratio, aniso)
        print(f"far {far:g} node extents, direction ratio up to 2^{ratio}, exponent spread {aniso}: hits {hit_frac:.2f}, lost {lost}, "
              f"false positives among the misses {fp:.3f}")
        assert 0.2 < hit_frac < 0.9
        assert lost == 0


def test_half_box_test_is_conservative_beyond_the_hand_over_ratio():
    """Directions with components down to 2^-30 of the largest and exactly zero components: the device sends these rays to the fp32
    test (PBR_H2_RATIO), but nothing may depend on that for correctness."""
    hit_frac, lost, fp = _run(11, 400_000, 100.0, 30, 3)
    print(f"wide-range directions: hits {hit_frac:.2f}, lost {lost}, false positives among the misses {fp:.3f}")
    assert lost == 0


def test_half_box_test_still_culls():
    """The bound must not be so wide that the test stops rejecting: for ordinary rays the false positives among near misses (the
    target point is at most 3 grid units outside the box) stay rare."""
    _, lost, fp = _run(5, 400_000, 30.0, 8, 1)
    assert lost == 0 and fp < 0.12

"""CPU checks behind Arith<__half>::div_fast (b2r_kernels.cuh, div_fast_half_operands): the fp16-mode sharpen
divides two half values through rcp.approx + one correction step in float and rounds to half.

(1) the lemma: a quotient of two 11-bit significands never lies on a rounding boundary of the half grid and keeps
    a relative distance > 2^-23 from every one of them (exhaustive over all 2^20 significand pairs);
(2) a float32 model of the device sequence with the reciprocal perturbed by up to 1 ulp (MUFU.RCP's error class)
    rounds to the same half as the exact quotient for every significand pair."""
import numpy as np


def _pairs():
    a = np.arange(1024, 2048, dtype=np.int64)
    return np.meshgrid(a, a, indexing="ij")


def test_quotients_keep_their_distance_from_half_rounding_boundaries():
    A, B = _pairs()
    # Q = A/B in (1/2, 2): boundaries are the odd multiples of 2^-s, s = 11 for Q >= 1, 12 for Q < 1
    s = np.where(A >= B, 11, 12)
    X = A << s
    r = X % (2 * B)                                   # (Q * 2^s) mod 2, in units of 1/B
    dist_num = np.abs(r - B)                          # |Q*2^s - nearest odd integer| * B
    assert dist_num.min() >= 1                        # never exactly on a boundary
    rel = dist_num.astype(np.float64) / (A.astype(np.float64) * np.exp2(s))   # relative to Q
    assert rel.min() > 2.0 ** -23, rel.min()


def _device_model(a, b, delta):
    """float32 model: r = fl(1/b * (1 + delta)); q0 = fl(a r); rem = a - b q0 (exact); q1 = fl(q0 + r rem)"""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    r = ((1.0 / b64) * (1.0 + delta)).astype(np.float32)
    q0 = (a64 * r.astype(np.float64)).astype(np.float32)           # 11 x 24 bits: exact in double, one rounding
    rem = a64 - b64 * q0.astype(np.float64)                        # 11 x 24 bits and the cancellation: exact
    assert np.all(rem == rem.astype(np.float32).astype(np.float64))  # representable, as the proof says
    q1 = (q0.astype(np.float64) + r.astype(np.float64) * rem).astype(np.float32)
    return q1


def test_one_correction_step_rounds_to_the_correct_half():
    A, B = _pairs()
    rng = np.random.default_rng(0)
    for ea in (0, -3, 5):                              # a few exponent offsets (the argument is scale-free)
        a = (A * 2.0 ** (ea - 10)).astype(np.float32)
        b = (B * 2.0 ** -10).astype(np.float32)
        exact = (a.astype(np.float64) / b.astype(np.float64))
        want = exact.astype(np.float16)                # one rounding of the exact quotient (double is exact enough:
        #                                                the gap to every boundary is > 2^-23 relative)
        for delta in (0.0, 2.0 ** -23, -2.0 ** -23, None):
            d = rng.uniform(-2.0 ** -23, 2.0 ** -23, A.shape) if delta is None else delta
            got = _device_model(a, b, d).astype(np.float16)
            assert np.array_equal(got.view(np.uint16), want.view(np.uint16))


def test_subnormal_half_results():
    """tiny numerators: the quotient lands in the subnormal range of half, where the grid is absolute (2^-24)"""
    A, B = _pairs()
    a = (A[::8, ::8] * 2.0 ** -10 * 2.0 ** -20).astype(np.float32)     # ~1e-6 .. 2e-6: half subnormals are < 6.1e-5
    b = (B[::8, ::8] * 2.0 ** -10).astype(np.float32)
    a = a.astype(np.float16).astype(np.float32)                         # operands must be half values
    want = (a.astype(np.float64) / b.astype(np.float64)).astype(np.float16)
    for delta in (2.0 ** -23, -2.0 ** -23):
        got = _device_model(a, b, delta).astype(np.float16)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16))

"""CPU restatement of the integer-slicing scheme of the INT8 mode (mogpe_b200/csrc/gemm_i8.cuh): the slicing rule, the exact
integer accumulation per anti-diagonal of slice pairs, and the error bound the kernel header states.  Pure numpy (int64
arithmetic stands in for the tensor cores' exact int32 accumulation); the GPU kernel is checked against the same FP64
products in tools/i8_gemm_test.cu and tests/test_gpu_int8.py."""
import numpy as np
import pytest

NS = 8


def pow2_scale(maxabs):
    """2^e with maxabs / 2^e < 1/2 (oz::pow2_scale)."""
    maxabs = np.asarray(maxabs, dtype=np.float64)
    m, e = np.frexp(maxabs)
    return np.where(maxabs > 0, np.ldexp(1.0, e + 1), 1.0)


def slice8(x, scale):
    """-> int slices q[NS, ...] with x ~= scale * sum_p q_p 2^(-7 (p + 1)), |q_p| <= 64 (oz::slice_digits): two rounds of
    round-to-nearest to 28 bits, each split into four balanced base-128 digits."""
    r = np.asarray(x, dtype=np.float64) / scale
    H = NS // 2
    P = float(1 << (7 * H))
    x1 = r * P
    hi = np.rint(x1)
    lo = np.rint((x1 - hi) * P)
    q = [None] * NS
    for base, v in ((0, hi.astype(np.int64)), (H, lo.astype(np.int64))):
        for k in range(H - 1, 0, -1):
            d = ((v + 64) & 127) - 64  # low 7 bits, sign-extended
            q[base + k] = d
            v = (v - d) >> 7
        q[base] = v
    return np.stack(q)


def sliced_product(A, B):
    """D = A @ B.T through row-scaled byte slices and exact integer accumulation of the pairs with i + j < NS."""
    sa = pow2_scale(np.max(np.abs(A), axis=1))[:, None]
    sb = pow2_scale(np.max(np.abs(B)))
    qa, qb = slice8(A, sa), slice8(B, sb)
    assert np.max(np.abs(qa)) <= 64 and np.max(np.abs(qb)) <= 64
    D = np.zeros((A.shape[0], B.shape[0]))
    for d in range(NS):
        T = sum(qa[i] @ qb[d - i].T for i in range(d + 1))  # exact integers
        assert np.max(np.abs(T)) < 2**31
        D += T.astype(np.float64) * 2.0 ** (-7 * (d + 2))
    return D * sa * sb


def test_slices_reconstruct_to_56_bits():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1000) * np.exp(5 * rng.standard_normal(1000))
    s = pow2_scale(np.max(np.abs(x)))
    q = slice8(x, s)
    rec = s * sum(q[p] * 2.0 ** (-7 * (p + 1)) for p in range(NS))
    assert np.max(np.abs(q)) <= 64
    assert np.max(np.abs(rec - x)) <= s * 2.0 ** (-7 * NS - 1) * 1.0000001


def test_sliced_product_is_fp64_accurate_even_with_cancellation():
    """An ill-conditioned triangular operand (rows of wildly different size, sums that cancel by 1e6): the error stays at the
    truncation bound ~ (NS + 1) K 2^(-7 NS) rowmax(A) max(B), far below what any FP32-class accumulation gives."""
    rng = np.random.default_rng(1)
    M, N = 96, 40
    L = np.tril(rng.standard_normal((M, M))) + 1e-3 * np.eye(M)
    A = np.linalg.inv(L @ L.T + 1e-6 * np.eye(M))  # entries up to ~1e6, alternating signs
    A = np.tril(A)
    B = rng.random((N, M))
    want = A @ B.T
    got = sliced_product(A, B)
    bound = (NS + 1) * M * 2.0 ** (-7 * NS) * 4 * np.max(np.abs(A), axis=1)[:, None] * np.max(np.abs(B))
    assert np.all(np.abs(got - want) <= bound + 1e-15 * np.abs(want).max())
    # relative to the row scale it is ~1e-13; a float32 product of the same operands is ~1e9 times worse
    f32 = (A.astype(np.float32) @ B.T.astype(np.float32)).astype(np.float64)
    rowmax = np.max(np.abs(A), axis=1)[:, None]
    assert np.max(np.abs(got - want) / rowmax) < 1e-12 < np.max(np.abs(f32 - want) / rowmax) * 1e-3


def test_word_slicing_restatement_matches_the_digit_rule():
    """oz::slice_words (gemm_i8.cuh): bias every base-128 digit by 64, spread the bit fields of the two 7 NS / 2-bit integers
    into byte lanes, remove the bias with a carry-free per-byte add.  The bytes must equal the balanced digits of slice8 for
    eight and for six slices, including the end points |r| = 1/2."""
    global NS
    keep = NS
    rng = np.random.default_rng(0)
    try:
        for ns in (8, 6):
            NS = ns
            H = ns // 2
            r = np.concatenate([rng.uniform(-0.5, 0.5, 100000), [0.5, -0.5, 0.0, 2.0**-30, -2.0**-30, 0.4999999999999999]])
            q = slice8(r, 1.0)
            P = float(1 << (7 * H))
            x1 = r * P
            hi = np.rint(x1)
            lo = np.rint((x1 - hi) * P)
            bias = sum(64 << (7 * k) for k in range(H))
            out = np.zeros((ns, r.size), dtype=np.int64)
            for half, v in ((0, hi.astype(np.int64)), (1, lo.astype(np.int64))):
                u = v + bias
                assert np.all(u >= 0)
                sh = 0 if H == 4 else 7
                w = ((u << (3 + sh)) & 0xFF000000) | ((u << (2 + sh)) & 0x7F0000) | ((u << (1 + sh)) & 0x7F00) | ((u & 0x7F) if H == 4 else 0x40)
                w = ((w & 0x7F7F7F7F) + 0x40404040) ^ (~w & 0x80808080)  # per-byte u - 64 without carries
                for b in range(4 - H, 4):
                    byte = (w >> (8 * b)) & 0xFF
                    out[half * H + 3 - b] = np.where(byte > 127, byte - 256, byte)
            assert np.array_equal(out, q)
            rec = sum(out[p] * 2.0 ** (-7 * (p + 1)) for p in range(ns))
            assert np.max(np.abs(rec - r)) <= 2.0 ** (-7 * ns - 1)
    finally:
        NS = keep


@pytest.mark.parametrize("H", [3, 4])
def test_unslice4_dp4a_restatement(H):
    """kernels_i8.cuh: unslice4 rebuilds a half (H signed base-128 digits) as  (u0 128 + u1) 128^(H-2) + (u2 128 + u3 | u2) - bias
    with u = digit ^ 0x80 as an unsigned byte (two DP4A with weights (128, 1), the bias in the accumulator): exact, and every
    intermediate stays below 2^32."""
    rng = np.random.default_rng(H)
    d = rng.integers(-64, 65, size=(200000, H))
    d[0], d[1] = -64, 64
    u = (d.astype(np.int8).view(np.uint8) ^ 0x80).astype(np.int64)
    bias = 128 * sum(128 ** k for k in range(H))
    top = u[:, 0] * 128 + u[:, 1]
    low = u[:, 2] * 128 + u[:, 3] if H == 4 else u[:, 2]
    val = top * (16384 if H == 4 else 128) + low
    assert val.max() < 2 ** 32
    ref = sum(d[:, k] * 128 ** (H - 1 - k) for k in range(H))
    assert np.array_equal(val - bias, ref)


@pytest.mark.parametrize("NS", [6, 8])
def test_integer_slicing_restatement(NS):
    """gemm_i8.cuh: slice_words_int -- the digits of r = v 2^eo from v's bit pattern with integer operations only: R =
    round(r 2^(7 NS)) as a rounded right shift of the mantissa (M << 3), the balanced low half by sign extension, the high half
    by subtraction.  R must equal the correctly rounded value (ties away from zero) and the halves must be balanced."""
    import struct

    B, TB = 7 * (NS // 2), 7 * NS
    rng = np.random.default_rng(NS)
    vals = np.concatenate([rng.uniform(-0.5, 0.5, 20000) * 2.0 ** rng.integers(-60, 1, 20000),
                           [0.0, 0.5, -0.5, 2.0 ** -57, -(2.0 ** -58), 0.5 - 2.0 ** -54, 1.5 * 2.0 ** -57]])
    for r in vals:
        eo = int(rng.integers(-40, 41))
        v = float(r) * 2.0 ** (-eo)  # the kernel sees v and the exponent of the output scale
        bits = struct.unpack("<q", struct.pack("<d", v))[0]
        hiw = (bits >> 32) & 0xFFFFFFFF
        ef = (hiw >> 20) & 0x7FF
        M = ((((hiw & 0xFFFFF) | 0x100000) << 32) | (bits & 0xFFFFFFFF)) if ef else 0
        t = min(max(3 - (ef - 1075 + eo + TB), 0), 63)
        Ru = (((M << 3) + ((1 << t) >> 1)) & (2 ** 64 - 1)) >> t
        R = -Ru if bits < 0 else Ru
        lo = ((R & (2 ** B - 1)) ^ (1 << (B - 1))) - (1 << (B - 1))  # sign-extended low B bits
        hi = (R - lo) >> B
        # exact reference: round half away from zero of r 2^TB in integer arithmetic
        from fractions import Fraction

        x = Fraction(float(r)) * 2 ** TB
        want = int(abs(x) + Fraction(1, 2)) * (1 if x >= 0 else -1)
        assert R == want, (r, eo, R, want)
        assert hi * 2 ** B + lo == R and -(2 ** (B - 1)) <= lo < 2 ** (B - 1) and abs(hi) <= 2 ** (B - 1)

"""The temporal-attention softmax divides every probability of a row by the same sum with ONE reciprocal
(csrc/temporal_attn.cu, softmax_rows):  y = RN(1/s), q0 = RN(x*y), r = x - q0*s (exact: fma), q1 = RN(q0 + r*y).
ATen's warp softmax computes exp(x - max) / s with an IEEE division (reference path: models/attention.py:480 ->
torch.softmax). This test restates both in exact rational arithmetic with correct round-to-nearest-even to binary32 and
checks q1 == RN(x / s) bit for bit over random and adversarial operands of the ranges that occur (1 <= s <= 32,
0 < x <= 1), i.e. that the kernel's probabilities are bit-identical to the reference's before the fp16 rounding.
(The GPU side of the same claim: test_kernels_gpu.py::test_temporal_attention_bit_exact_on_exact_inputs.)"""
import random
import struct
from fractions import Fraction


def f32_to_frac(bits: int) -> Fraction:
    return Fraction(struct.unpack("<f", struct.pack("<I", bits))[0])


def rn32(v: Fraction) -> Fraction:
    """round-to-nearest-even of a positive rational to binary32 (normal range), returned as the exact value"""
    if v == 0:
        return Fraction(0)
    assert v > 0
    e = v.numerator.bit_length() - v.denominator.bit_length()  # 2^(e-1) <= v < 2^(e+1)
    if Fraction(2) ** e > v:
        e -= 1
    assert Fraction(2) ** e <= v < Fraction(2) ** (e + 1)
    assert -126 <= e <= 127, "operands of this test stay in the normal range"
    ulp = Fraction(2) ** (e - 23)
    k = v / ulp                      # in [2^23, 2^24)
    lo = k.numerator // k.denominator
    rem = k - lo
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (lo & 1)):
        lo += 1
    return lo * ulp


def signed_rn32(v: Fraction) -> Fraction:
    return -rn32(-v) if v < 0 else rn32(v)


def shared_reciprocal_quotient(x: Fraction, s: Fraction) -> Fraction:
    y = rn32(1 / s)
    q0 = rn32(x * y)
    r = signed_rn32(x - q0 * s)      # fma: one rounding (exact here; asserted by the caller through the result)
    return signed_rn32(q0 + r * y)


def rand_f32(rng, lo_exp, hi_exp):
    e = rng.randint(lo_exp, hi_exp)
    return Fraction((1 << 23) | rng.getrandbits(23)) * Fraction(2) ** (e - 23)


def test_shared_reciprocal_division_is_correctly_rounded():
    rng = random.Random(1234)
    cases = []
    for _ in range(6000):
        s = rand_f32(rng, 0, 4)                       # sum of <= 32 exponentials, each <= 1, one of them == 1
        if s > 32:
            s = Fraction(32)
        x = rand_f32(rng, -24, -1)                    # exp(x - max): anything that survives the fp16 rounding
        cases.append((x, s))
    # adversarial sums: mantissa all ones / all zeros / alternating, and x at the extremes of a binade
    for e in range(0, 5):
        for m in (0x7FFFFF, 0x000000, 0x000001, 0x555555, 0x2AAAAA, 0x7FFFFE, 0x400000):
            s = Fraction((1 << 23) | m) * Fraction(2) ** (e - 23)
            for xm in (0x000000, 0x7FFFFF, 0x000001, 0x400000, 0x3FFFFF):
                for xe in (-1, -7, -13, -20):
                    cases.append((Fraction((1 << 23) | xm) * Fraction(2) ** (xe - 23), s))
        cases.append((Fraction(1), Fraction(2) ** e))
    bad = [(x, s) for x, s in cases if shared_reciprocal_quotient(x, s) != rn32(x / s)]
    assert not bad, f"{len(bad)} of {len(cases)} quotients differ from the IEEE division, e.g. {bad[:3]}"

"""The bit-placement identities the dequant functors rely on, checked exhaustively on the CPU (numpy).  These are the arithmetic behind
`F4Quarter` (qmatmul_tc.cu) and the integer e4m3 expansion measured in profiles/r02_fp8_attention.md: a small float's bits dropped into an
f16 ARE the f16 of the value times a power of two, subnormals included, so one HMUL2 finishes the dequantisation exactly."""
import numpy as np

from oracle import fp_formats as F


def _f16(bits):
    return np.asarray(bits, np.uint16).view(np.float16).astype(np.float64)


def test_e2m1_nibble_placed_at_bits_9_to_11_is_the_value_times_2_pow_minus_14():
    nib = np.arange(16, dtype=np.uint16)
    placed = ((nib & 7) << 9) | ((nib & 8) << 12)                      # magnitude -> bits 9..11, sign -> bit 15
    want = F.e2m1_to_f32(nib.astype(np.uint8)).astype(np.float64)
    got = _f16(placed) * 2.0 ** 14
    assert np.array_equal(np.abs(got), np.abs(want)) and np.array_equal(np.signbit(got), (nib & 8) != 0)      # -0.0 keeps its sign bit


def test_e4m3_scale_bits_shifted_by_7_then_times_2_pow_14_is_the_scale_times_64_exactly():
    codes = np.arange(0x7f, dtype=np.uint16)                            # every non-negative finite e4m3 code, 0 and subnormals included
    s_h = _f16((codes & 0x7f) << 7)                                     # = scale * 2^-8 (exact; f16 subnormals for e4m3 subnormals)
    scale = F.e4m3_to_f32(codes.astype(np.uint8)).astype(np.float64)
    assert np.array_equal(s_h * 2.0 ** 8, scale)
    s6 = (s_h * 2.0 ** 14).astype(np.float16).astype(np.float64)        # the HMUL2 by 2^14, rounded to f16
    assert np.array_equal(s6, scale * 64.0) and s6.max() < 65504


def test_nvfp4_product_of_placed_nibble_and_scale_is_exact_in_f16():
    nib = np.arange(16, dtype=np.uint16)
    w_h = _f16(((nib & 7) << 9) | ((nib & 8) << 12))                   # weight * 2^-14
    codes = np.arange(0x7f, dtype=np.uint16)
    s6 = _f16((codes & 0x7f) << 7) * 2.0 ** 14                          # scale * 2^6
    prod = np.outer(w_h, s6)                                            # = weight * scale * 2^-8
    assert np.array_equal(prod.astype(np.float16).astype(np.float64), prod)      # <= 6 significant bits, inside f16's range: no rounding


def test_mxfp4_exponent_field_is_the_e8m0_scale_times_64():
    e = np.arange(107, 137, dtype=np.int64)                             # the range the kernel represents exactly: 2^-20 .. 2^9
    s6 = _f16(((e - 106).astype(np.uint16)) << 10)
    assert np.array_equal(s6, 2.0 ** (e - 127) * 64.0)


def test_e4m3_bits_moved_down_by_one_is_the_value_times_2_pow_minus_8():
    b = np.arange(256, dtype=np.uint16)
    b = b[(b & 0x7f) != 0x7f]                                           # skip the two NaN codes
    y = b << 8                                                          # the byte in the high half of a 16-bit lane
    f16_bits = ((y >> 1) & 0x3f80) | (y & 0x8000)
    assert np.array_equal(_f16(f16_bits) * 2.0 ** 8, F.e4m3_to_f32(b.astype(np.uint8)).astype(np.float64))


def test_ggml_nibble_and_6bit_subnormal_placements():
    """Q4_K / int4: a nibble at f16 mantissa bits 6..9 is q * 2^-18; Q6_K: a 6-bit value at bits 4..9 is q * 2^-20 (both f16 subnormals):
    one HFMA2 with (scale * 2^18 | 2^20, offset) then yields scale * q + offset with a single rounding."""
    q4 = np.arange(16, dtype=np.uint16)
    assert np.array_equal(_f16(q4 << 6) * 2.0 ** 18, q4.astype(np.float64))
    q6 = np.arange(64, dtype=np.uint16)
    assert np.array_equal(_f16(q6 << 4) * 2.0 ** 20, q6.astype(np.float64))
    # the exact fallback: (1024 + q) as f16 via the 0x6400 magic, minus 1024
    assert np.array_equal(_f16(q4 | 0x6400) - 1024.0, q4.astype(np.float64))

"""CPU model of the error-compensated TF32 product the noisy-head kernels use on the tensor cores (3xTF32:
csrc/rb_head_tc.cu compose stage, csrc/rb_head.cu mma3_block): hi = v & 0xFFFFE000, lo = v - hi, and
a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with every operand truncated to TF32 by the MMA and fp32 accumulation.

The GPU parity tests measure the result (<= 1e-6 relative against a fp32 library GEMM); this test pins the REASON on the host: over
the head's reduction length (K = 3136) the compensated product is as accurate as a plain fp32 dot product, and a single TF32
product is three orders of magnitude worse -- so the "fp32-equivalent" claim does not hang on one lucky input."""
import numpy as np
import pytest


def tf32(x):
    """What a TF32 MMA reads of an fp32 operand: sign, exponent, the upper 10 mantissa bits (low 13 bits ignored)."""
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = tf32(x)
    lo = (np.asarray(x, np.float32) - hi).astype(np.float32)      # exact: both share the exponent range
    return hi, lo


def dot_blocks(a, b, block=8):
    """sum_k a[k] b[k] the way an m16n8k8 MMA chain accumulates: exact products, one fp32 rounding per k-block of 8."""
    acc = np.float32(0.0)
    for k in range(0, a.size, block):
        acc = np.float32(acc + np.float32(np.dot(a[k:k + block].astype(np.float64), b[k:k + block].astype(np.float64))))
    return acc


def dot_3xtf32(a, b):
    ah, al = split(a)
    bh, bl = split(b)
    acc = np.float32(0.0)
    for k in range(0, a.size, 8):   # small terms first inside every k-block, as the kernels issue them
        s = slice(k, k + 8)
        for x, y in ((tf32(al[s]), bh[s]), (ah[s], tf32(bl[s])), (ah[s], bh[s])):
            acc = np.float32(acc + np.float32(np.dot(x.astype(np.float64), y.astype(np.float64))))
    return acc


@pytest.mark.parametrize("K,seed", [(3136, 0), (3136, 1), (576, 2), (512, 3)])
def test_compensated_tf32_dot_is_fp32_accurate(K, seed):
    rs = np.random.RandomState(seed)
    rel3, rel1, rel32 = [], [], []
    for _ in range(24):
        # a row of noisy weights (mu + sigma * eps products) against post-ReLU conv features
        w = (rs.uniform(-1, 1, K) / np.sqrt(K) + 0.5 / np.sqrt(K) * rs.standard_normal(K) * rs.standard_normal()).astype(np.float32)
        x = np.maximum(rs.standard_normal(K), 0).astype(np.float32)
        truth = float(np.dot(w.astype(np.float64), x.astype(np.float64)))
        scale = float(np.dot(np.abs(w).astype(np.float64), x.astype(np.float64)))     # condition-free error measure
        rel3.append(abs(float(dot_3xtf32(w, x)) - truth) / scale)
        rel1.append(abs(float(dot_blocks(tf32(w), tf32(x))) - truth) / scale)
        rel32.append(abs(float(dot_blocks(w, x)) - truth) / scale)
    assert max(rel3) <= 2e-7                          # the lo*lo term (2^-22 relative) is the only thing dropped
    assert max(rel3) <= 4 * max(max(rel32), 3e-8)     # as good as fp32 FMA accumulation
    assert np.median(rel1) >= 100 * np.median(rel3)   # one TF32 product alone is not


def test_split_is_exact_and_lo_fits_tf32_twice():
    rs = np.random.RandomState(5)
    v = (rs.standard_normal(4096) * np.exp(rs.uniform(-20, 20, 4096))).astype(np.float32)
    hi, lo = split(v)
    assert np.array_equal((hi.astype(np.float64) + lo.astype(np.float64)).astype(np.float32), v)
    assert np.all(np.abs(lo) <= np.abs(v) * 2.0 ** -10)
    # what the MMA drops of lo is below 2^-21 of v: the source of the 2e-7 bound above
    assert np.all(np.abs(lo - tf32(lo)) <= np.abs(v) * 2.0 ** -20)

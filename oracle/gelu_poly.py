"""TEST INFRASTRUCTURE (never imported by the product package): float32 restatement of the epilogue GELU of the CUDA kernels.

nn.GELU() (exact, erf form; reference lib/model/DSTformer.py:73, applied at :81) is evaluated by the sm_100a epilogues
(motionbert_b200/csrc/gemm_tc.cuh::gelu_erf2) as

    gelu(x) = x * (0.5 + copysign(0.5 - 2^P(t), x)),   t = min(|x|, 5.75),   2^P(t) ~= Phi(-t) = 0.5 erfc(t / sqrt 2)

with P a degree-8 polynomial (weighted-minimax fit of log2(0.5 erfc(t / sqrt 2)); one MUFU.EX2 per value).  This module
re-evaluates that formula step by step in float32 (every FMA rounded once) so that tests/test_oracle.py can pin the
coefficients compiled into the kernels against float64 erf."""
import re

import numpy as np

CLAMP = 5.75


def coefficients_from_source(path):
    """The coefficients as written in gelu_erf2 (highest degree first ... constant term), parsed from the CUDA source."""
    src = open(path).read()
    body = src[src.index("float2 gelu_erf2(float2 x)"):]
    body = body[:body.index("return __fmul2_rn(x, phi);")]
    first = re.search(r"__ffma2_rn\(make_float2\(([-0-9.e+]+)f, [-0-9.e+]+f\), t, make_float2\(([-0-9.e+]+)f", body)
    rest = re.findall(r"p = __ffma2_rn\(p, t, make_float2\(([-0-9.e+]+)f", body)
    return [float(first.group(1)), float(first.group(2))] + [float(v) for v in rest]


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def gelu_kernel_form(x, coeffs):
    """float32 emulation of gelu_erf2 (ex2 taken as correctly rounded: MUFU.EX2 is within 2^-22 relative)."""
    x = np.asarray(x, dtype=np.float32)
    t = np.minimum(np.abs(x), np.float32(CLAMP))
    c = [np.float32(v) for v in coeffs]
    p = _fma32(np.full_like(t, c[0]), t, np.full_like(t, c[1]))
    for v in c[2:]:
        p = _fma32(p, t, np.full_like(t, v))
    q = np.exp2(p.astype(np.float64)).astype(np.float32)
    d = (np.float32(0.5) - q).astype(np.float32)
    phi = (np.float32(0.5) + np.copysign(d, x)).astype(np.float32)
    return (x * phi).astype(np.float32)


def gelu_exact(x):
    from scipy.special import erf
    x = np.asarray(x, dtype=np.float64)
    return x * 0.5 * (1.0 + erf(x / np.sqrt(2.0)))

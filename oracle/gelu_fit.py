"""ORACLE / TEST INFRASTRUCTURE: derivation and check of the GELU used in the CUDA epilogues (csrc/common.cuh
gelu_erf_fast): erfc(z) = exp2(P7(z/2 - 1)) with P7 a least-squares fit of log2(erfc) on [0, 4]; fp32 Horner emulation
against the exact 0.5 x (1 + erf(x / sqrt 2)) (HF ACT2FN["gelu"])."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

ZMAX = 4.0
COEFFS = [-7.739973545e+00, -1.274813366e+01, -5.330767155e+00, -1.898051500e-01, 8.234396577e-02, -3.462206945e-02,
          1.296435855e-02, -2.886363771e-03]   # ascending powers of t; the constants in common.cuh


def fit(deg: int = 7):
    z = np.linspace(0, ZMAX, 40001)
    t = 2 * z / ZMAX - 1
    return C.cheb2poly(C.chebfit(t, np.log2(erfc(z)), deg))


def gelu_fast_fp32(x: np.ndarray, coeffs=COEFFS) -> np.ndarray:
    x = x.astype(np.float32)
    co = [np.float32(v) for v in coeffs]
    t = np.minimum(np.abs(x) * np.float32(0.35355339059327373) + np.float32(-1.0), np.float32(1.0)).astype(np.float32)
    acc = np.full_like(t, co[-1])
    for k in range(len(co) - 2, -1, -1):
        acc = (acc * t + co[k]).astype(np.float32)
    e = np.exp2(acc).astype(np.float32)
    hx = (np.float32(0.5) * x).astype(np.float32)
    a = np.abs(hx)
    return ((hx + a) - a * e).astype(np.float32)


def max_error() -> float:
    x = np.linspace(-9, 9, 400001)
    ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    return float(np.abs(gelu_fast_fp32(x) - ref).max())


if __name__ == "__main__":
    print("fitted:", ", ".join(f"{v:.9e}" for v in fit()))
    print("max |gelu_fast - gelu| in fp32:", max_error())

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


# ------------------------------------------------------------------------------------------------------------------
# Round 2: the tanh form the CUDA epilogues use now (csrc/common.cuh gelu_fast2)
#     gelu(x) ~= h + h * tanh(x * (B1 + B3 x^2 + B5 x^4)),  h = x / 2,  x^2 clamped at 64
# ------------------------------------------------------------------------------------------------------------------
TANH_COEFFS = [7.975078789e-01, 3.700565057e-02, -3.515174775e-04]


def gelu_tanh_fp32(x: np.ndarray, coeffs=TANH_COEFFS, tanh_rel_err: float = 0.0, seed: int = 1) -> np.ndarray:
    """fp32 emulation; tanh_rel_err models MUFU.TANH (PTX: max relative error 2^-11) as uniform relative noise."""
    x = x.astype(np.float32)
    b1, b3, b5 = (np.float32(v) for v in coeffs)
    x2 = np.minimum(x * x, np.float32(64.0)).astype(np.float32)
    u = ((x2 * b5 + b3) * x2 + b1).astype(np.float32) * x
    t = np.tanh(u.astype(np.float64))
    if tanh_rel_err:
        t = t * (1.0 + (np.random.default_rng(seed).random(t.shape) * 2 - 1) * tanh_rel_err)
    h = (np.float32(0.5) * x).astype(np.float32)
    return (h + h * t.astype(np.float32)).astype(np.float32)


def fit_tanh(n_coef: int = 3):
    """Minimax (iteratively re-weighted least squares) fit of the polynomial inside the tanh on [0, 6]."""
    from scipy.optimize import least_squares

    xs = np.linspace(0, 6, 60001)
    ex = 0.5 * xs * (1 + erf(xs / np.sqrt(2)))

    def model(b):
        x2 = xs * xs
        u, pw = b[0], x2
        for c in b[1:]:
            u = u + c * pw
            pw = pw * x2
        return 0.5 * xs * (1 + np.tanh(xs * u))
    b = np.array([np.sqrt(2 / np.pi), 0.044715 * np.sqrt(2 / np.pi)] + [0.0] * (n_coef - 2))
    w = np.ones_like(xs)
    for _ in range(60):
        b = least_squares(lambda bb: (model(bb) - ex) * w, b, xtol=1e-15, ftol=1e-15).x
        e = np.abs(model(b) - ex)
        w = w * (1 + 4 * e / e.max())
        w /= w.mean()
    return b


def max_error_tanh(tanh_rel_err: float = 0.0) -> float:
    x = np.linspace(-12, 12, 600001)
    ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    return float(np.abs(gelu_tanh_fp32(x, tanh_rel_err=tanh_rel_err) - ref).max())


if __name__ == "__main__":
    print("tanh form, fitted:", ", ".join(f"{v:.9e}" for v in fit_tanh()))
    print("max |gelu_tanh - gelu| in fp32, exact tanh:", max_error_tanh())
    print("max |gelu_tanh - gelu| in fp32, tanh with 2^-11 relative noise (|x| <= 12):", max_error_tanh(2.0 ** -11))

"""CPU: the float32 error bound the CUDA code relies on (encode.cuh k_fine_argmin32 guard, plan.cuh k_lut_f32 slack),
    |d32 - d| <= 8 * 2^-24 * (sqrt(d32 * S2) + ds * d32) + 1e-13 * S2,   S2 >= |p|^2 + |c|^2,
checked by emulating the device arithmetic (inputs rounded to float32, FSUB, then an FFMA chain) in NumPy over many
magnitudes, including nearly equal vectors (cancellation) and tiny / huge norms."""
import numpy as np


def _d32(p64, c64):
    p = p64.astype(np.float32)
    c = c64.astype(np.float32)
    df = (p - c).astype(np.float32)                       # FSUB, float32 rounding
    acc = np.zeros(p.shape[0], np.float32)
    for t in range(p.shape[1]):                           # FFMA chain: round32(df*df + acc), product exact in float64
        acc = (df[:, t].astype(np.float64) * df[:, t].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    return acc, p


def test_float32_distance_error_bound():
    rng = np.random.RandomState(0)
    u8 = np.float32(8.0 * 5.9604645e-08)
    worst = 0.0
    for ds in (2, 4, 8, 16):
        for scale in (1e-3, 0.1, 1.0, 30.0):
            for gap in (0.0, 1e-7, 1e-4, 1e-2, 1.0):
                n = 20000
                c = rng.randn(n, ds) * scale
                p = c * (1.0 + gap * rng.randn(n, ds)) + gap * scale * rng.randn(n, ds)
                d = ((p - c) ** 2).sum(1)                                  # float64 reference
                d32, p32 = _d32(p, c)
                s2 = ((p32.astype(np.float64) ** 2).sum(1) + (c ** 2).sum(1)) * 1.0001 + 1e-30
                bound = u8 * (np.sqrt(d32.astype(np.float64) * s2) + ds * d32.astype(np.float64)) + 1e-13 * s2
                err = np.abs(d32.astype(np.float64) - d)
                assert (err <= bound).all(), (ds, scale, gap, float((err / np.maximum(bound, 1e-300)).max()))
                worst = max(worst, float((err / np.maximum(bound, 1e-300)).max()))
    assert worst < 0.8          # the constant 8 leaves head-room


def test_quantised_sum_is_a_lower_bound():
    """plan.cuh k_lut_quant: code = clamp(floor((e - b) * inv), 0, QMAX) in float32  =>  sum(e) >= B + Delta * (S - 0.1)."""
    rng = np.random.RandomState(1)
    M, QMAX = 16, 65535 // 16
    for trial in range(200):
        e = (rng.rand(M, 256) ** 2 * rng.uniform(1e-3, 2.0)).astype(np.float32)
        b = e.min(1)
        rng_ = float(e.max() - b.min())
        delta = rng_ / QMAX * (1.0 + 9.5367431640625e-07)
        inv = np.float32(1.0 / delta)
        code = np.clip(np.floor(((e - b[:, None]).astype(np.float32) * inv).astype(np.float32)), 0, QMAX).astype(np.int64)
        pick = rng.randint(0, 256, size=(1000, M))
        S = code[np.arange(M)[None, :], pick].sum(1)
        d = e.astype(np.float64)[np.arange(M)[None, :], pick].sum(1)
        lb = b.astype(np.float64).sum() + delta * (S - 0.1)
        assert (d >= lb * (1 - 1e-12) - 1e-300).all()
        assert S.max() <= 65535

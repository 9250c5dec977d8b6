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


def test_large_v_preselection_bound():
    """largev.cuh k_presel_emit: the float32 ADC distance of a retrieved code (float32 copies of projection and codebook,
    FSUB + one FFMA chain over all D dimensions) against the float64 distance:
        |d32 - d| <= E = (D + 4) * 2^-24 * 1.01 * 2 * (|p|^2 + |c|^2)
    (the kernel uses upper bounds of |p|^2 and |c|^2, which only widens E).  Emulated over magnitudes, near-equal vectors
    (cancellation) and mixed scales per dimension; also the selection argument: with d32_k the k-th smallest d32, every
    member of the exact first k has d32 <= d32_k + 2E."""
    rng = np.random.RandomState(7)
    u = 5.9604645e-08
    worst = 0.0
    for D in (32, 128, 256):
        for scale in (1e-4, 0.05, 1.0, 300.0):
            for gap in (0.0, 1e-6, 1e-3, 0.3, 3.0):
                n = 4000
                c = rng.randn(n, D) * scale * np.exp(rng.randn(1, D))              # uneven dimensions
                p = c * (1.0 + gap * rng.randn(n, D)) + gap * scale * rng.randn(n, D)
                p32 = p.astype(np.float32).astype(np.float64)                        # the kernel reads float32 copies of both
                d = ((p - c) ** 2).sum(1)
                d32, _ = _d32(p, c)
                E = (D + 4) * u * 1.01 * 2.0 * ((p ** 2).sum(1) + (c ** 2).sum(1))
                err = np.abs(d32.astype(np.float64) - d)
                assert (err <= E + 1e-300).all(), (D, scale, gap, float((err / np.maximum(E, 1e-300)).max()))
                worst = max(worst, float((err / np.maximum(E, 1e-300)).max()))
                # selection: one query against n candidates
                k = 50
                q = rng.randn(D) * scale
                dq = ((q[None, :] - c) ** 2).sum(1)
                dq32, _ = _d32(np.repeat(q[None, :], n, 0), c)
                Eq = float(((D + 4) * u * 1.01 * 2.0 * ((q ** 2).sum() + (c ** 2).sum(1))).max())
                kth32 = np.sort(dq32.astype(np.float64))[k - 1]
                first_k = np.argsort(dq, kind="stable")[:k]
                assert (dq32[first_k].astype(np.float64) <= kth32 + 2.0 * Eq).all(), (D, scale, gap)
    assert worst < 0.5          # the bound is an upper bound with head-room, not an estimate


def test_score_form_guard_bound():
    """encode.cuh k_fine_argmin32 / k_coarse_big: the score  s = |c|^2 / 2 - x.c  evaluated in float32 (x and c rounded to
    float32, half norm by an FFMA chain over the float32 centroid, then one FFMA per dimension starting from it) against
    its float64 value:  |s32 - s| <= E = (n + 4) * 2^-24 * (|x| + |c|)^2  (the coarse kernel uses n + 16).  A centroid is
    accepted only when the runner-up is more than 3 E away, so E has to be an upper bound over magnitudes and sizes."""
    rng = np.random.RandomState(11)
    u = 5.9604645e-08
    worst = 0.0
    for nd in (2, 8, 16, 64, 128):
        for scale in (1e-3, 0.2, 1.0, 50.0):
            for near in (0.0, 1e-5, 1e-2, 1.0):
                n = 5000
                c = rng.randn(n, nd) * scale
                x = c * (1.0 + near * rng.randn(n, nd)) + near * scale * rng.randn(n, nd)
                s = 0.5 * (c ** 2).sum(1) - (x * c).sum(1)
                x32, c32 = x.astype(np.float32), c.astype(np.float32)
                hn = np.zeros(n, np.float32)
                for t in range(nd):
                    hn = (c32[:, t].astype(np.float64) * c32[:, t].astype(np.float64) + hn.astype(np.float64)).astype(np.float32)
                v = (np.float32(0.5) * hn).astype(np.float32)
                for t in range(nd):
                    v = ((-x32[:, t]).astype(np.float64) * c32[:, t].astype(np.float64) + v.astype(np.float64)).astype(np.float32)
                E = (nd + 4) * u * (np.sqrt((x ** 2).sum(1)) + np.sqrt((c ** 2).sum(1))) ** 2
                err = np.abs(v.astype(np.float64) - s)
                assert (err <= E + 1e-300).all(), (nd, scale, near, float((err / np.maximum(E, 1e-300)).max()))
                worst = max(worst, float((err / np.maximum(E, 1e-300)).max()))
    assert worst < 0.6


def _tf32_rna(x32):
    """float32 -> nearest TF32 (10 explicit mantissa bits), ties away from zero: cvt.rna.tf32.f32 / ftc_tf32_host."""
    u = np.asarray(x32, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _to_f32(x64, toward_zero):
    y = x64.astype(np.float32)
    if toward_zero:
        over = np.abs(y.astype(np.float64)) > np.abs(x64)
        y = np.where(over, np.nextafter(y, np.float32(0.0)), y).astype(np.float32)
    return y


def test_three_tf32_pieces_score_budget():
    """fine_tc.cuh: the score |c|^2/2 - p.c as the product of ( p_hi | p_hi | p_lo | 1 1 1 ) with ( -c_hi | -c_lo | -c_hi |
    h_hi h_lo h_lo2 ): TF32 pieces rounded to nearest, products exact, float32 accumulation -- emulated (a) with one
    round-to-nearest per product and (b) with ONE TRUNCATION PER MMA INSTRUCTION (8 exact products added to the accumulator,
    then rounded toward zero: the hardware measurement on adversarial rows, profiles/dev/ftc_adversarial_probe.py, shows the
    positive bias of a truncating accumulator at 1.8 * 2^-24 worst case).  Either way the error stays below
    8 * 2^-24 (|p| + |c|)^2, half of the E = 16 * 2^-24 (|p| + max|c|)^2 the guard uses.  p == c rows (all products of
    the winning score have one sign) are the worst case for (b)."""
    rng = np.random.RandomState(3)
    u = 5.9604645e-08
    for ds in (8, 16):
        for scale in (1e-3, 0.3, 1.0, 40.0):
            for near in (0.0, 1e-4, 0.1, 1.0):
                n = 4000
                c = rng.randn(n, ds) * scale
                p = c * (1.0 + near * rng.randn(n, ds)) + near * scale * rng.randn(n, ds)
                s = 0.5 * (c ** 2).sum(1) - (p * c).sum(1)
                p32, c32 = p.astype(np.float32), c.astype(np.float32)
                p_hi = _tf32_rna(p32); p_lo = _tf32_rna((p32 - p_hi).astype(np.float32))
                c_hi = _tf32_rna(c32); c_lo = _tf32_rna((c32 - c_hi).astype(np.float32))
                h = 0.5 * (c ** 2).sum(1)
                pieces, res = [], h.copy()
                for _ in range(3):
                    piece = _tf32_rna(res.astype(np.float32))
                    pieces.append(piece)
                    res = res - piece.astype(np.float64)
                bound = 8.0 * u * (np.sqrt((p ** 2).sum(1)) + np.sqrt((c ** 2).sum(1))) ** 2
                terms = [(p_hi[:, t], -c_hi[:, t]) for t in range(ds)] + [(p_hi[:, t], -c_lo[:, t]) for t in range(ds)] + \
                        [(p_lo[:, t], -c_hi[:, t]) for t in range(ds)] + [(np.ones(n, np.float32), pc) for pc in pieces]
                for per_instruction_truncation in (False, True):
                    acc = np.zeros(n, np.float32)
                    group = 8 if per_instruction_truncation else 1
                    for g0 in range(0, len(terms), group):
                        tot = acc.astype(np.float64)
                        for a, b in terms[g0:g0 + group]:
                            tot = tot + a.astype(np.float64) * b.astype(np.float64)
                        acc = _to_f32(tot, per_instruction_truncation)
                    err = np.abs(acc.astype(np.float64) - s)
                    assert (err <= bound + 1e-300).all(), (ds, scale, near, per_instruction_truncation,
                                                           float((err / np.maximum(bound, 1e-300)).max()))

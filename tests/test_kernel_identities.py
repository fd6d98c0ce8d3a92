"""CPU tests (NumPy / exact rationals): the arithmetic identities the CUDA kernels lean on, restated outside the kernels so
that each claim in DESIGN.md section 4 has a check that runs without a GPU.  The GPU tests compare the kernels themselves
with the oracle; these pin the bounds the launchers and the fast paths assume.

  * fixed-point chunk totals (k_prep_corr, k_carr_partial: navlab-dpe-sdr_b200/csrc/dpe_prepare.cu, dpe_vel.cu)
  * block-moment expansion of the carrier spectrum (k_carr_partial) and the launcher's bound (launch_score_vel)
  * centre-relative range and the recovered rounding of tx (code_index_fast_core, dpe_geom.cuh)
"""
import math
from fractions import Fraction

import numpy as np

K_C = 299792458.0
FIX_SCALE = 524288.0                    # kFixScale, dpe_internal.cuh: 2^19


def test_fixed_point_chunk_totals_are_order_independent_exact_and_cannot_overflow():
    rng = np.random.default_rng(20180704)
    nchunk, nval = 4096, 160                                  # chunks x (NLp lags x A/B x re/im)
    mag = 10.0 ** rng.uniform(-2.0, 7.6, size=(nchunk, nval))  # FP32 partials from 0.01 to 4e7
    part = (rng.standard_normal((nchunk, nval)) * mag).astype(np.float32)
    q = np.rint(part.astype(np.float64) * FIX_SCALE).astype(np.int64)     # __double2ll_rn((double)v * kFixScale)
    tot = q.sum(axis=0)
    for _ in range(3):                                        # any arrival order of the CTAs: the same integers
        assert np.array_equal(q[rng.permutation(nchunk)].sum(axis=0), tot)
    # a partial of magnitude >= 16 is represented exactly (its ulp is >= 2^-19) ...
    big = np.abs(part) >= 16.0
    assert np.array_equal(q[big].astype(np.float64) / FIX_SCALE, part[big].astype(np.float64))
    # ... and a smaller one to half a unit of 2^-19, so the total is within nchunk * 2^-20 of the exact sum
    exact = np.array([sum(Fraction(float(v)) for v in part[:, k]) for k in range(0, nval, 40)])
    got = np.array([Fraction(int(t), int(FIX_SCALE)) for t in tot[::40]])
    assert max(abs(g - e) for g, e in zip(got, exact)) <= Fraction(nchunk, 2 ** 20)
    # back to FP64 the way the last CTA does it ((double)total * 2^-19) is exact below 2^53
    assert np.array_equal(tot.astype(np.float64) * (1.0 / FIX_SCALE) * FIX_SCALE, tot.astype(np.float64))
    # overflow: a chunk of 1024 samples of |x| <= 32768 sqrt(2) -- doubled for the DC-removed baseband of the carrier
    # branch, |x - mean| <= 2 |x|max -- stays below 2^27, and S <= 2^26 means at most 2^16 chunks
    assert 1024 * 2 * 32768 * math.sqrt(2.0) < 2 ** 27
    assert (2 ** 27) * int(FIX_SCALE) * (2 ** 16) <= 2 ** 62
    # FP64 sums of the same partials DO depend on the order (why the totals are integers)
    a = part[:, 0].astype(np.float64)
    orders = {float(np.sum(a[rng.permutation(nchunk)])) for _ in range(20)}
    assert len(orders) > 1


def _moment_bin(bb, n0, m, n_fft):
    """k_carr_partial, one 32-sample block and one bin, in FP64: the four complex moments and the 4-term expansion."""
    b = np.arange(32)
    t = (b - 15.5) / 16.0
    M = [np.sum(t ** k * bb) for k in range(4)]
    phi = 32.0 * math.pi * m / n_fft
    centre = np.exp(-2j * math.pi * ((2 * n0 + 31) * m % (2 * n_fft)) / (2 * n_fft))     # exact integer phase of n0 + 15.5
    return centre * (M[0] - 1j * phi * M[1] - 0.5 * phi * phi * M[2] + 1j * phi ** 3 / 6.0 * M[3])


def test_block_moment_expansion_of_the_carrier_spectrum_meets_the_launchers_bound():
    rng = np.random.default_rng(7)
    for S, Wd in ((50000, 64), (200000, 64), (50000, 300)):
        n_fft = 8 * (1 << int(math.ceil(math.log2(S))))         # carrSTot, batchcorrscores.cu:761
        x = 2.0 * math.pi * (Wd + 1) * 15.5 / n_fft             # launch_score_vel's criterion
        bound = x ** 4 / 24.0
        takes_moments = bound < 2.0e-8 and n_fft <= (1 << 23)
        worst = 0.0
        for _ in range(200):
            bb = rng.standard_normal(32) + 1j * rng.standard_normal(32)
            n0 = 32 * int(rng.integers(0, S // 32))
            m = int(rng.integers(-Wd - 1, Wd + 2))
            exact = np.sum(bb * np.exp(-2j * math.pi * ((n0 + np.arange(32)) * m % n_fft) / n_fft))
            worst = max(worst, abs(_moment_bin(bb, n0, m, n_fft) - exact) / np.sum(np.abs(bb)))
        assert worst <= bound + 1e-15                            # Taylor remainder phi^4 t^4 / 24, |t| <= 15.5 / 16
        if (S, Wd) == (50000, 300):
            assert not takes_moments                             # too wide a Doppler window: the direct kernel
        else:
            assert takes_moments and bound < 1e-9                # demo / c4 shapes: below the FP32 rounding of the sums


def _fast_range(rho, u, D):
    """code_index_fast_core's range, in FP64 (the kernel fuses two of the products; same truncation)."""
    a = float(np.dot(u, D))
    d2 = float(np.dot(D, D))
    return (rho - a) + ((d2 - a * a) * (0.5 / rho)) * (1.0 + a / rho)


def test_centre_relative_range_stays_inside_the_margin_the_fast_path_assumes():
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(4000):
        d0 = rng.standard_normal(3)
        d0 *= rng.uniform(2.0e7, 2.7e7) / np.linalg.norm(d0)    # satellite minus grid centre
        D = rng.standard_normal(3)
        D *= rng.uniform(0.0, 1000.0) / np.linalg.norm(D)       # candidate minus grid centre, up to 1 km
        rho = float(np.linalg.norm(d0))
        exact = np.sqrt(np.sum((d0.astype(np.longdouble) - D.astype(np.longdouble)) ** 2))
        worst = max(worst, abs(float(np.longdouble(_fast_range(rho, d0 / rho, D)) - exact)))
    # kTxMargin = 4e-16 s = 1.2e-7 m was chosen as 5 x the error bound of this range (dpe_geom.cuh)
    assert worst < 4.0e-16 * K_C / 5.0


def test_rounding_error_of_tx_is_recovered_exactly_and_the_margin_decides_safely():
    rng = np.random.default_rng(13)
    margin = 4.0e-16                                            # kTxMargin
    decided = 0
    for _ in range(20000):
        rx = float(rng.uniform(1.0e5, 6.0e5))                   # GPS seconds of week
        q = float(rng.uniform(0.06, 0.09))                      # pseudorange / c
        tx = rx - q
        err = (rx - tx) - q                                     # Fast2Sum (|q| < |rx|): exact
        assert Fraction(rx) - Fraction(q) - Fraction(tx) == Fraction(err)
        half_ulp = math.ldexp(1.0, math.frexp(tx)[1] - 1 - 53)  # 2^(exponent - 53), as the kernel forms it from the bits
        assert half_ulp == math.ulp(tx) / 2.0
        if abs(err) < half_ulp - margin:                        # the fast path keeps its tx ...
            decided += 1
            for dq in (-0.99 * margin, 0.99 * margin):          # ... and a pseudorange off by less than the margin rounds alike
                assert rx - (q + dq) == tx
    assert decided > 19900                                       # the exact chain is the rare case (~1e-5 of the pairs)

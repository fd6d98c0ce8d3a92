"""NumPy float64 restatement of the reference's DPE hot path (TEST INFRASTRUCTURE).

Every function cites the CUDARecv source it follows (paths relative to
``/root/reference/cudarecv``).  Arithmetic is IEEE float64 evaluated in the
same operation order as the reference source (no FMA contraction), so integer
outputs (chip indices, nav-bit edges, correlogram bins) are reproducible
bit-for-bit by any implementation that evaluates the same expressions.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may
import this module (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

# utils/inc/consthelper.h:5-27
CONST_C = 299792458.0
CONST_PI = 3.1415926535898
CONST_2PI = 6.2831853071796
CONST_F_L1 = 1.57542e9
CONST_F_CA = 1.023e6
CONST_L_CA = 1023
CONST_T_CA = 0.001
CONST_PRN_MAX = 37
CONST_MU = 3.986005e14
CONST_F = -4.442807633e-10
CONST_OEDot = 7.2921151467e-5
CONST_WGS84_A = 6378137.0
CONST_WGS84_B = 6356752.314245
CONST_WGS84_E = 0.08181919084262149
CONST_WGS84_EP = 0.08209443794969568

GRID_UNIFORM = 0       # utils/inc/gridhelper.h:30-35
GRID_EXPONENTIAL = 1
GRID_ARTHURBASIS = 2


def posmod(a, b):
    """auxil/inc/auxil.h:11  POSMOD(a,b) = (((a)%(b))+(b))%(b) on C ints."""
    return np.mod(a, b)  # numpy mod is already the positive modulo for b > 0


def round_up_pow2(x: int) -> int:
    """auxil/src/auxil.cpp:98-109."""
    x = int(x) - 1
    for s in (1, 2, 4, 8, 16):
        x |= x >> s
    return x + 1


# --------------------------------------------------------------------------
# BatchCorrScores (modules/src/batchcorrscores.cu)
# --------------------------------------------------------------------------

_G2_TAP1 = [2, 3, 4, 5, 1, 2, 1, 2, 3, 2, 3, 5, 6, 7, 8, 9, 1, 2,
            3, 4, 5, 6, 1, 4, 5, 6, 7, 8, 1, 2, 3, 4, 5, 4, 1, 2, 4]
_G2_TAP2 = [6, 7, 8, 9, 9, 10, 8, 9, 10, 3, 4, 6, 7, 8, 9, 10, 4,
            5, 6, 7, 8, 9, 3, 6, 7, 8, 9, 10, 6, 7, 8, 9, 10, 10, 7, 8, 10]


def gen_ca_code(prn: int) -> np.ndarray:
    """C/A Gold code of ``prn`` (1..37) as int8 +/-1.

    batchcorrscores.cu:117-177 (BCS_GenCACode): two 10-stage registers
    initialised to -1, G1 feedback reg1[2]*reg1[9], G2 feedback
    reg2[1]*reg2[2]*reg2[5]*reg2[7]*reg2[8]*reg2[9], G2 output from the
    phase-select taps tap1/tap2, chip = -g1*g2.
    (The reference only fills PRN 1..36: ``prn < numChan`` with numChan=37,
    :125-127; PRN 37 is generated here for completeness and never used.)
    """
    assert 1 <= prn <= CONST_PRN_MAX
    reg1 = [-1] * 10
    reg2 = [-1] * 10
    t1 = _G2_TAP1[prn - 1] - 1
    t2 = _G2_TAP2[prn - 1] - 1
    code = np.empty(1023, dtype=np.int8)
    for i in range(1023):
        g1 = reg1[9]
        g2 = reg2[t1] * reg2[t2]
        fb1 = reg1[2] * reg1[9]
        fb2 = reg2[1] * reg2[2] * reg2[5] * reg2[7] * reg2[8] * reg2[9]
        reg1 = [fb1] + reg1[:9]
        reg2 = [fb2] + reg2[:9]
        code[i] = -g1 * g2
    return code


_CA_TABLE = None


def ca_table() -> np.ndarray:
    """[37][1023] int8 table (chipsCACode_d, batchcorrscores.cu:748-749)."""
    global _CA_TABLE
    if _CA_TABLE is None:
        _CA_TABLE = np.stack([gen_ca_code(p) for p in range(1, CONST_PRN_MAX + 1)])
    return _CA_TABLE


def time_idcs(S: int, fs: float) -> np.ndarray:
    """batchcorrscores.cu:185-196: t[i] = round(i/fs*1e9)/1e9 (C round())."""
    t = np.arange(S, dtype=np.float64) / fs
    return np.floor(t * 1.0e9 + 0.5) / 1.0e9


def nav_bit_boundary(cp_elapsed, cp_reference, code_phase, code_freq, fs):
    """batchcorrscores.cu:237-258 (BCS_NavBitBoundary).  Returns int32[C]."""
    cp_elapsed = np.asarray(cp_elapsed, dtype=np.int64)
    cp_reference = np.asarray(cp_reference, dtype=np.int64)
    cp_since = posmod(cp_elapsed - cp_reference, 20)
    cp_to_next = 20 - cp_since
    v = (CONST_L_CA * cp_to_next - np.asarray(code_phase, np.float64)) * \
        (fs / np.asarray(code_freq, np.float64))
    return (np.floor(v) + 1).astype(np.int32)


def doppler_wipeoff(carr_freq: float, carr_phase: float, t: np.ndarray) -> np.ndarray:
    """batchcorrscores.cu:277-305: conj(exp(j*2*CONST_PI*(fi*t + ri)))."""
    ph = 2 * CONST_PI * (carr_freq * t + carr_phase)
    return np.cos(ph) - 1j * np.sin(ph)


def chip_index(t: np.ndarray, code_freq: float, code_phase: float) -> np.ndarray:
    """batchcorrscores.cu:347-348: POSMOD((int)floor(t*fc + rc), 1023)."""
    return posmod(np.floor(t * code_freq + code_phase).astype(np.int64),
                  CONST_L_CA).astype(np.int32)


def code_replica(prn: int, t, code_freq, code_phase, idx_next: int, S: int):
    """batchcorrscores.cu:323-372 (BCS_ComputeCodeReplica).

    Returns (chip_idx int32[S], no_flip f64[S], flipped f64[S]).  The flipped
    replica is the no-flip replica negated from ``idx_next`` on when
    0 < idx_next < S, else all zeros (:352-367).
    """
    ci = chip_index(t, code_freq, code_phase)
    no_flip = ca_table()[prn - 1][ci].astype(np.float64)
    if 0 < idx_next < S:
        flipped = no_flip.copy()
        flipped[idx_next:] = -flipped[idx_next:]
    else:
        flipped = np.zeros(S)
    return ci, no_flip, flipped


def iq_to_complex(iq_int16: np.ndarray) -> np.ndarray:
    """batchcorrscores.cu:209-221 (BCS_Load): interleaved int16 I,Q -> complex."""
    iq = np.asarray(iq_int16, dtype=np.int16).reshape(-1, 2).astype(np.float64)
    return iq[:, 0] + 1j * iq[:, 1]


def batch_corr_scores(iq_int16, prn, rc, ri, fc, fi, cp_elapsed, cp_ref, fs,
                      want_carrier=False, n_fft=None):
    """One BatchCorrScores::Update (batchcorrscores.cu:975-1208) on the CPU.

    Inputs are referenced to the START of the block (:692, :1077-1086).
    Returns a dict with
      code_scores  complex128 [C][S]  fft-shifted circular correlogram of the
                   chosen (flip / no-flip) replica: code_scores[c][j] =
                   c_c[(j - S/2) mod S], c_c[k] = sum_n xw[(n+k) mod S] r[n]
                   (:1099-1153)
      no_flip      bool [C]           BCS_ChooseCodeCorr flag (:499-543)
      idx_next     int32 [C]          nav-bit edge sample (:237-258)
      chip_idx     int32 [C][S]       C/A chip index per sample (:347-348)
      raw_mean     complex            DC mean (:1065, :1210-1216)
      carr_scores  complex128 [C][Nc] fft-shifted zero-padded carrier spectrum
                   (:1158-1180), only if want_carrier
    """
    x = iq_to_complex(iq_int16)
    S = x.shape[0]
    C = len(prn)
    t = time_idcs(S, fs)
    idx_next = nav_bit_boundary(cp_elapsed, cp_ref, rc, fc, fs)
    code_scores = np.empty((C, S), dtype=np.complex128)
    chip_idx = np.empty((C, S), dtype=np.int32)
    no_flip_flag = np.empty(C, dtype=bool)
    raw_mean = x.sum() / np.float32(S)
    if n_fft is None:
        n_fft = round_up_pow2(S) * 8           # :761
    carr = np.empty((C, n_fft), dtype=np.complex128) if want_carrier else None
    half = S // 2
    for c in range(C):
        w = doppler_wipeoff(fi[c], ri[c], t)
        ci, r_nf, r_fl = code_replica(int(prn[c]), t, fc[c], rc[c], int(idx_next[c]), S)
        chip_idx[c] = ci
        xw = x * w                                                # :1113
        xw_fft = np.fft.fft(xw)                                   # :1120
        c_nf = np.fft.ifft(xw_fft * np.conj(np.fft.fft(r_nf)))    # :1099-1144
        c_fl = np.fft.ifft(xw_fft * np.conj(np.fft.fft(r_fl)))
        edge = 0 < idx_next[c] < S
        nf = (not edge) or (abs(c_nf[0]) > abs(c_fl[0]))          # :512
        no_flip_flag[c] = nf
        chosen = c_nf if nf else c_fl
        code_scores[c] = np.concatenate((chosen[half:], chosen[:half]))  # :554-584
        if want_carrier:
            zm = xw - raw_mean * w                                # :470-485
            bb = zm * (r_nf if nf else r_fl)                      # :422-452
            sp = np.fft.fft(bb, n_fft)                            # :1179
            carr[c] = np.concatenate((sp[n_fft // 2:], sp[:n_fft // 2]))  # :1180
    return dict(code_scores=code_scores, no_flip=no_flip_flag, idx_next=idx_next,
                chip_idx=chip_idx, raw_mean=raw_mean, carr_scores=carr, n_fft=n_fft)


# --------------------------------------------------------------------------
# BatchCorrManifold (modules/src/batchcorrmanifold.cu)
# --------------------------------------------------------------------------

def init_pos_grid(dims, spacing, grid_type=GRID_UNIFORM):
    """batchcorrmanifold.cu:148-255 (BCM_InitPosGrid) + host setup :2328-2333.

    dims: 4 ints (x,y,z,t); spacing: 4 floats.  Flat index = ((ix*Ny+iy)*Nz+iz)*Nt+it
    (t fastest, :164-170); half index = (dim-1)/2 (:2331).
    Returns (grid f64 [G][4], time_grid f64 [Nt]).
    """
    dims = [int(d) for d in dims]
    sp = [float(s) for s in spacing]
    half = [(d - 1) // 2 for d in dims]

    def axis(n, h, s):
        idx = np.arange(n, dtype=np.int64)
        if grid_type == GRID_UNIFORM:
            return s * (idx - h).astype(np.float64)
        if grid_type == GRID_ARTHURBASIS:                       # :190-240
            outer = (idx < h // 2) | ((n - idx) < h // 2)
            lo = 3 * s * (idx - h) + s * ((h // 2) + 1) * 2
            hi = 3 * s * (idx - h) - s * ((h // 2) + 1) * 2
            inner = s * (idx - h).astype(np.float64)
            return np.where(outer, np.where(idx < h, lo, hi), inner).astype(np.float64)
        raise ValueError("unsupported manifold type")           # :247-249

    ax = [axis(dims[k], half[k], sp[k]) for k in range(4)]
    gx, gy, gz, gt = np.meshgrid(*ax, indexing="ij")
    grid = np.stack([gx.ravel(), gy.ravel(), gz.ravel(), gt.ravel()], axis=1)
    return np.ascontiguousarray(grid), ax[3].copy()


def load_pos_grid_csv(path):
    """batchcorrmanifold.cu:2422-2448: 'x,y,z,delta_t' per line (atof)."""
    rows = []
    with open(path, "r") as f:
        for line in f:
            p = line.strip().split(",")
            if len(p) < 4:
                continue
            rows.append([float(v) for v in p[:4]])
    return np.asarray(rows, dtype=np.float64)


def _norm3(a, b, c):
    return np.sqrt(a * a + b * b + c * c)


def candidate_ecef(grid, center, enu2ecef):
    """batchcorrmanifold.cu:1760-1763."""
    R = np.asarray(enu2ecef, np.float64)
    gx, gy, gz, gt = grid[:, 0], grid[:, 1], grid[:, 2], grid[:, 3]
    px = R[0] * gx + R[1] * gy + R[2] * gz + center[0]
    py = R[3] * gx + R[4] * gy + R[5] * gz + center[1]
    pz = R[6] * gx + R[7] * gy + R[8] * gz + center[2]
    pt = gt + center[3]
    return px, py, pz, pt


def pos_bins(grid, center, enu2ecef, sat_states, time_dim, code_freq, code_phase_end,
             cp_ref_tow, cp_elapsed_end, cp_ref, rx_time, fs, S, per_time_sat=False):
    """Geometry -> correlogram bin for every (candidate, channel).

    batchcorrmanifold.cu:1754-1800 (BCM_PosMeasML).  ``sat_states`` is the
    [C*T][8] array of CHM_GridPrep; the ML kernel uses the MIDDLE time-grid
    state for every candidate (:1773-1775); ``per_time_sat`` selects the
    per-``it`` state of the (dormant) reduction kernel (:865-873).

    Returns idx_base f64 [G][C] (bc_rc0_idx_base), idxo (idx + S*chan),
    f_idx / c_idx int64 [G][C] (row-offset bins, :1797-1799) and valid bool.
    """
    grid = np.asarray(grid, np.float64)
    G = grid.shape[0]
    C = len(code_freq)
    sat = np.asarray(sat_states, np.float64).reshape(C, time_dim, 8)
    px, py, pz, pt = candidate_ecef(grid, center, enu2ecef)
    idx_base = np.empty((G, C))
    it = np.arange(G) % time_dim
    for c in range(C):
        if per_time_sat:
            s = sat[c, it]                     # [G][8]
            sx, sy, sz, sdt = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
        else:
            s = sat[c, time_dim // 2]
            sx, sy, sz, sdt = s[0], s[1], s[2], s[3]
        lx, ly, lz = sx - px, sy - py, sz - pz
        rng = _norm3(lx, ly, lz)                                   # :1782
        pr = rng - CONST_C * sdt + pt                              # :1783
        tx = rx_time - pr / CONST_C                                # :1784
        frac = tx - cp_ref_tow[c] - ((int(cp_elapsed_end[c]) - int(cp_ref[c])) * CONST_T_CA)  # :1785
        bc_rc = frac * CONST_F_CA                                  # :1786
        bc_rc0 = bc_rc - code_phase_end[c]                         # :1790
        idx_base[:, c] = (fs / code_freq[c]) * (-bc_rc0) + S / 2.0  # :1791
    valid = (idx_base < S) & (idx_base > 0)                        # :1795
    idxo = idx_base + (S * np.arange(C))[None, :]                  # :1797
    f_idx = np.floor(idxo).astype(np.int64)                        # :1798
    c_idx = np.floor(idxo + 1).astype(np.int64)                    # :1799
    return idx_base, idxo, f_idx, c_idx, valid


def pos_scores_from_bins(code_scores, idxo, f_idx, c_idx, valid, lpower=1):
    """batchcorrmanifold.cu:1806-1816: lerp of the two bins, sum |.|^L over PRNs.

    Out-of-window (candidate, PRN) pairs read stale indices in the reference
    (:1795-1809, undefined behaviour); the oracle defines their contribution
    as 0 and parity sets keep every pair in-window.
    """
    flat = np.asarray(code_scores).reshape(-1)
    fi = np.where(valid, f_idx, 0)
    ci = np.where(valid, c_idx, 0)
    ci = np.minimum(ci, flat.shape[0] - 1)
    v = flat[ci] * (idxo - f_idx) + flat[fi] * (c_idx - idxo)      # :1810-1812
    mag = np.abs(v) ** lpower                                      # :1816
    mag = np.where(valid, mag, 0.0)
    # reference adds channel by channel (:1771-1818)
    score = np.zeros(idxo.shape[0])
    for c in range(idxo.shape[1]):
        score = score + mag[:, c]
    return score, v


def pos_meas_ml(code_scores, grid, center, enu2ecef, sat_states, time_dim, code_freq,
                code_phase_end, cp_ref_tow, cp_elapsed_end, cp_ref, rx_time, fs, S, lpower=1):
    """BCM_PosMeasML + thrust::max_element + BCM_MakePosMeas.

    batchcorrmanifold.cu:1710-1828, :2589 (first maximum), :1977-2016.
    Returns dict(scores, argmax, z[4], bins...).
    """
    idx_base, idxo, f_idx, c_idx, valid = pos_bins(
        grid, center, enu2ecef, sat_states, time_dim, code_freq, code_phase_end,
        cp_ref_tow, cp_elapsed_end, cp_ref, rx_time, fs, S, per_time_sat=False)
    scores, v = pos_scores_from_bins(code_scores, idxo, f_idx, c_idx, valid, lpower)
    i = int(np.argmax(scores))
    g = np.asarray(grid, np.float64)[i:i + 1]
    px, py, pz, pt = candidate_ecef(g, center, enu2ecef)
    z = np.array([px[0], py[0], pz[0], pt[0]])
    return dict(scores=scores, argmax=i, z=z, idx_base=idx_base, idxo=idxo,
                f_idx=f_idx, c_idx=c_idx, valid=valid, v=v)


def pos_meas_weighted(code_scores, grid, center, enu2ecef, sat_states, time_dim, code_freq,
                      code_phase_end, cp_ref_tow, cp_elapsed_end, cp_ref, rx_time, fs, S,
                      lpower=1, per_time_sat=True):
    """Score-weighted mean estimate (dormant in the reference, required by the
    north star): z = sum_i s_i*(p_i, dt_i) / sum_i s_i.

    batchcorrmanifold.cu:816-1056 (BCM_PosMeasReduction: per-``it`` satellite
    state :873, accumulation :911-915) + :1365-1510 (BCM_ReduceAndPosMeas,
    zVal = sum/score :1497-1500).  The bin expression of the reduction kernel
    (:876-882) is algebraically the ML kernel's; the oracle evaluates the ML
    form (the one the reference actually runs) for both.
    """
    idx_base, idxo, f_idx, c_idx, valid = pos_bins(
        grid, center, enu2ecef, sat_states, time_dim, code_freq, code_phase_end,
        cp_ref_tow, cp_elapsed_end, cp_ref, rx_time, fs, S, per_time_sat=per_time_sat)
    scores, _ = pos_scores_from_bins(code_scores, idxo, f_idx, c_idx, valid, lpower)
    px, py, pz, pt = candidate_ecef(np.asarray(grid, np.float64), center, enu2ecef)
    tot = scores.sum()
    z = np.array([(scores * px).sum(), (scores * py).sum(),
                  (scores * pz).sum(), (scores * pt).sum()]) / tot
    return dict(scores=scores, z=z, sum_score=tot)


def pos_meas_reduction(code_scores, grid, center, enu2ecef, sat_states, time_dim, code_freq, tx_time,
                       rx_time, fs, S, lpower=1, n_blocks=8, n_threads=64):
    """The reference's DORMANT weighted-mean estimator exactly as written: BCM_PosMeasReduction
    (batchcorrmanifold.cu:816-1056) + BCM_ReduceAndPosMeas (:1365-1510), launch shape <<<8, 64>>> then
    <<<1, 8>>> (:2548-2563, commented out there; oracle/ref_driver.cu -DREF_WEIGHTED launches them).

    Differences from the ML kernel that this restatement keeps: the satellite state is the one of the
    candidate's own time-grid index (:865-873), and the code index is back-calculated from ``txTime``
    (:876-882: satPosTransmitTime = rxTime - (txTime + dt/c) + sat.dt; bc_rc0 = F_CA (that - range/c)),
    algebraically the ML expression (:1779-1791) in another rounding order.

    Returns dict(z[4], parts[n_blocks][5] = per-block {a, b, c, d, score} (grid-stride partition of the
    candidates over blocks * threads), sum_score, scores[G]).
    """
    grid = np.asarray(grid, np.float64)
    G = grid.shape[0]
    C = len(code_freq)
    sat = np.asarray(sat_states, np.float64).reshape(C, time_dim, 8)
    px, py, pz, pt = candidate_ecef(grid, center, enu2ecef)
    it = np.arange(G) % time_dim                                          # :865
    flat = np.asarray(code_scores).reshape(-1)
    scores = np.zeros(G)
    for c in range(C):
        s = sat[c, it]                                                    # :873
        tof = rx_time - (tx_time[c] + (pt / CONST_C)) + s[:, 3]           # :876
        rngt = _norm3(s[:, 0] - px, s[:, 1] - py, s[:, 2] - pz) / CONST_C  # :880
        bc_rc0 = CONST_F_CA * (tof - rngt)                                # :881
        idx_base = (fs / code_freq[c]) * (-bc_rc0) + S / 2.0              # :882
        idxo = idx_base + S * c                                           # :888
        f_idx = np.floor(idxo).astype(np.int64)
        c_idx = np.floor(idxo + 1).astype(np.int64)
        v = flat[np.minimum(c_idx, flat.shape[0] - 1)] * (idxo - f_idx) + flat[f_idx] * (c_idx - idxo)   # :901-902
        scores = scores + np.abs(v) ** lpower                             # :906
    stride = n_blocks * n_threads
    parts = np.zeros((n_blocks, 5))
    w = np.stack([scores * px, scores * py, scores * pz, scores * pt, scores], axis=1)    # :911-915
    tid = np.arange(G) % stride
    for b in range(n_blocks):
        parts[b] = w[(tid // n_threads) == b].sum(axis=0)
    tot = parts.sum(axis=0)
    return dict(z=tot[:4] / tot[4], parts=parts, sum_score=tot[4], scores=scores)   # :1497-1500


# ---- velocity manifold (next row f-1) --------------------------------------

def init_vel_grid(dims, spacing):
    """batchcorrmanifold.cu:265-316 (BCM_InitVelGrid; Uniform == ArthurBasis)."""
    g, _ = init_pos_grid(dims, spacing, GRID_UNIFORM)
    return g


def vel_bins(grid, center, enu2ecef, sat_states, time_dim, carr_freq, doppler_sign, fs, n_fft):
    """batchcorrmanifold.cu:1896-1945 (BCM_VelMeasML)."""
    grid = np.asarray(grid, np.float64)
    C = len(carr_freq)
    sat = np.asarray(sat_states, np.float64).reshape(C, time_dim, 8)
    R = np.asarray(enu2ecef, np.float64)
    gx, gy, gz, gt = grid[:, 0], grid[:, 1], grid[:, 2], grid[:, 3]
    vx = R[0] * gx + R[1] * gy + R[2] * gz + center[4]
    vy = R[3] * gx + R[4] * gy + R[5] * gz + center[5]
    vz = R[6] * gx + R[7] * gy + R[8] * gz + center[6]
    vd = gt + center[7]
    ex = vx - CONST_OEDot * center[1]
    ey = vy + CONST_OEDot * center[0]
    ez = vz
    idx_base = np.empty((grid.shape[0], C))
    for c in range(C):
        s = sat[c, time_dim // 2]
        lx, ly, lz = s[0] - center[0], s[1] - center[1], s[2] - center[2]
        rng = _norm3(lx, ly, lz)
        rate = ((lx / rng) * (ex - s[4])) + ((ly / rng) * (ey - s[5])) + ((lz / rng) * (ez - s[6]))
        bc_fi = CONST_F_L1 * ((rate - vd) / CONST_C + s[7]) / doppler_sign
        bc_fi0 = bc_fi - carr_freq[c]
        idx_base[:, c] = (n_fft / fs) * bc_fi0 + n_fft / 2.0
    valid = (idx_base < n_fft) & (idx_base > 0)
    idxo = idx_base + (n_fft * np.arange(C))[None, :]
    f_idx = np.floor(idxo).astype(np.int64)
    c_idx = np.floor(idxo + 1).astype(np.int64)
    return idx_base, idxo, f_idx, c_idx, valid


def vel_meas_ml(carr_scores, grid, center, enu2ecef, sat_states, time_dim, carr_freq,
                doppler_sign, fs, n_fft, lpower=1):
    """BCM_VelMeasML + max_element + BCM_MakeVelMeas (:1861-1963, :2590, :2030-2068)."""
    idx_base, idxo, f_idx, c_idx, valid = vel_bins(
        grid, center, enu2ecef, sat_states, time_dim, carr_freq, doppler_sign, fs, n_fft)
    scores, v = pos_scores_from_bins(carr_scores, idxo, f_idx, c_idx, valid, lpower)
    i = int(np.argmax(scores))
    g = np.asarray(grid, np.float64)[i]
    R = np.asarray(enu2ecef, np.float64)
    z = np.array([R[0] * g[0] + R[1] * g[1] + R[2] * g[2] + center[4],
                  R[3] * g[0] + R[4] * g[1] + R[5] * g[2] + center[5],
                  R[6] * g[0] + R[7] * g[1] + R[8] * g[2] + center[6],
                  g[3] + center[7]])
    return dict(scores=scores, argmax=i, z=z, idx_base=idx_base, idxo=idxo,
                f_idx=f_idx, c_idx=c_idx, valid=valid)


def vel_meas_reduction(carr_scores, grid, center, enu2ecef, sat_states, time_dim, carr_freq,
                       doppler_sign, fs, n_fft, lpower=1, n_blocks=8, n_threads=64):
    """The reference's DORMANT score-weighted velocity estimate exactly as written: BCM_VelMeasReduction
    (batchcorrmanifold.cu:1090-1347) + BCM_ReduceAndVelMeas (:1525-1667), launch shape <<<8, 64>>> then <<<1, 8>>>
    (:2555-2566, commented out there; oracle/ref_driver.cu -DREF_WEIGHTED launches them).  The per-candidate score is
    BCM_VelMeasML's (middle time-grid satellite state, :1149; line of sight to the grid centre, :1161-1168); the
    estimate is  z[4:8] = sum_i score_i (v_i, drift_i) / sum_i score_i  over the ECEF velocity candidates (:1193-1197,
    :1658-1661).  Returns dict(z[4], parts[n_blocks][5], sum_score, scores[Gv])."""
    idx_base, idxo, f_idx, c_idx, valid = vel_bins(
        grid, center, enu2ecef, sat_states, time_dim, carr_freq, doppler_sign, fs, n_fft)
    scores, _ = pos_scores_from_bins(carr_scores, idxo, f_idx, c_idx, valid, lpower)
    g = np.asarray(grid, np.float64)
    R = np.asarray(enu2ecef, np.float64)
    vx = R[0] * g[:, 0] + R[1] * g[:, 1] + R[2] * g[:, 2] + center[4]      # :1138-1141
    vy = R[3] * g[:, 0] + R[4] * g[:, 1] + R[5] * g[:, 2] + center[5]
    vz = R[6] * g[:, 0] + R[7] * g[:, 1] + R[8] * g[:, 2] + center[6]
    vd = g[:, 3] + center[7]
    w = np.stack([scores * vx, scores * vy, scores * vz, scores * vd, scores], axis=1)
    stride = n_blocks * n_threads
    tid = np.arange(g.shape[0]) % stride
    parts = np.zeros((n_blocks, 5))
    for b in range(n_blocks):
        parts[b] = w[(tid // n_threads) == b].sum(axis=0)
    tot = parts.sum(axis=0)
    return dict(z=tot[:4] / tot[4], parts=parts, sum_score=tot[4], scores=scores)


# --------------------------------------------------------------------------
# Brute-force identity (SURVEY section 8 a'): direct time-domain correlation
# against the blended replica; used to validate the north-star kernel's
# formulation on small cases.
# --------------------------------------------------------------------------

def blended_correlation(xw, replica, k, alpha):
    """v = sum_n xw[n] * ((1-a) r[(n-k) mod S] + a r[(n-k-1) mod S]).

    Equals (1-a) c[k] + a c[k+1] with c the circular correlogram of
    batch_corr_scores (before the fft shift).
    """
    r0 = np.roll(replica, k)
    r1 = np.roll(replica, k + 1)
    return np.sum(xw * ((1.0 - alpha) * r0 + alpha * r1))

"""Host-side sharding of the candidate grid over ranks and the combination of
per-rank partial estimates (the host mirror of ``k_finalize`` in
``csrc/dpe_score.cu``).

Partial layout (``DPE_PARTIAL_LEN`` doubles, ``include/dpe_b200.h``):
  [0..3] sum_i s_i*(x,y,z,c*dt)_i   [4] sum_i s_i   [5] max score
  [6] global arg-max index          [7] out-of-window pairs
  [8..11] ECEF x,y,z and clock (m) of the rank's arg-max candidate
"""
from __future__ import annotations

import numpy as np

PARTIAL_LEN = 16
EST_ARGMAX, EST_WEIGHTED = 0, 1


def shard_range(G: int, world: int, rank: int):
    """Contiguous index range [lo, hi) of rank ``rank``: ceil(G/world) candidates per
    rank, flat order x slowest / t fastest (batchcorrmanifold.cu:164-170), so rank
    order == ascending global index."""
    per = (G + world - 1) // world
    lo = min(rank * per, G)
    return lo, min(lo + per, G)


def make_partial(scores, ecef_dt, global_offset=0, out_of_window=0):
    """Partial of one shard from its scores [n] and candidate states [n][4]."""
    p = np.zeros(PARTIAL_LEN)
    scores = np.asarray(scores, dtype=np.float64)
    if scores.size == 0:
        p[5], p[6] = -1.0, 9.0e18
        return p
    ecef_dt = np.asarray(ecef_dt, dtype=np.float64)
    p[0:4] = (scores[:, None] * ecef_dt).sum(axis=0)
    p[4] = scores.sum()
    i = int(np.argmax(scores))                      # first maximum
    p[5], p[6], p[7] = scores[i], global_offset + i, out_of_window
    p[8:12] = ecef_dt[i]
    return p


def combine_partials(parts, est_mode=EST_ARGMAX):
    """Combine gathered partials [nranks][PARTIAL_LEN] -> dict(z[4], max_score, sum_score,
    argmax, out_of_window).  Ties on the maximum go to the lowest global index
    (thrust::max_element / np.argmax semantics, batchcorrmanifold.cu:2589)."""
    parts = np.asarray(parts, dtype=np.float64).reshape(-1, PARTIAL_LEN)
    tot = parts[:, 0:5].sum(axis=0)
    best, mx, mi = -1, -1.0, 9.0e18
    for r in range(parts.shape[0]):
        if parts[r, 5] > mx or (parts[r, 5] == mx and parts[r, 6] < mi):
            best, mx, mi = r, parts[r, 5], parts[r, 6]
    if est_mode == EST_WEIGHTED:
        z = tot[0:4] / tot[4]
    else:
        z = parts[best, 8:12].copy() if best >= 0 else np.zeros(4)
    return dict(z=z, max_score=mx, sum_score=tot[4], argmax=int(mi), out_of_window=int(parts[:, 7].sum()))

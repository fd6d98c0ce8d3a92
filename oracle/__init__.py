"""CPU oracle for the DPE batch-correlation-manifold hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker / the
reported CPU baseline.  The product path (``navlab-dpe-sdr_b200``) never imports
this package and fails loudly when its CUDA library is missing.

The oracle is a plain NumPy float64 restatement of the reference's algorithm
(CUDARecv ``modules/src/batchcorrscores.cu``, ``batchcorrmanifold.cu``,
``cuchanmgr.cu``; mirrored on the CPU by PyGNSS ``pythonreceiver``), each
function citing the reference file:line it follows.

Pinning status: the reference ships NO tests, golden vectors or fixtures for
this path.  The oracle is therefore pinned against
(a) independent known answers (IS-GPS-200 first-10-chip octal table for the C/A
codes, the ``ECEF_to_LLA`` docstring values of ``pygnss/.../utils.py:23-26``),
(b) the mutual consistency of the two checked-in demo files
(``handoff_params_usrp6.csv`` + ``nist1860.18n``: back-calculated code phase
within 0.006 chip of the handed-off code phase for all 8 PRNs), and (c) outputs
of the reference CUDARecv modules themselves, rebuilt unmodified for sm_100a
(``oracle/Makefile`` -> ``oracle/_ref``) and run on a B200 through
``oracle/ref_driver.cu``; those outputs are committed under ``tests/golden/``
together with the script that produced them (``oracle/make_golden_ref.py``).
"""

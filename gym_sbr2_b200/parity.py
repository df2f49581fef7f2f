"""The parity definition used by tests, smoke() and bench.py (DESIGN.md section "Parity").

BASELINE.json: phase-end states and rewards agree with the reference's scipy-odeint path within rtol 1e-5 in
FP64; discrete outputs match exactly.  The reference integrates with LSODA at rtol = atol = 1.49e-8, so a
component that sits near zero (So ~ 4e-7 g/m3 at the end of an unaerated cycle) carries the reference's own
ABSOLUTE error of ~5e-10; a pure relative test would be testing LSODA's noise.  Hence the mixed test
    |a - b| <= RTOL * |b| + ATOL_FRAC * scale_i ,  scale = the reference's own normalisation vector
x_1_state (gym_SBR_oneshot.py:153), RTOL = 1e-5, ATOL_FRAC = 1e-9.
"""
import numpy as np

RTOL = 1e-5
ATOL_FRAC = 1e-9
# x_1_state[1:] of the reference: V, Si, Ss, Xi, Xs, Xbh, Xba, Xp, So, Sno, Snh, Snd, Xnd, Salk
STATE_SCALE = np.array([1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], dtype=np.float64)
KLA_SCALE = 240.0


def state_close(a, b, rtol=RTOL, atol_frac=ATOL_FRAC):
    """a, b: [..., 14] (component last) or [14].  Returns (ok, worst) where worst = max over entries of
    |a-b| / (rtol*|b| + atol_frac*scale) (<= 1 passes)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bound = rtol * np.abs(b) + atol_frac * STATE_SCALE
    ratio = np.abs(a - b) / bound
    worst = float(np.nanmax(ratio)) if ratio.size else 0.0
    ok = bool(np.all(np.isfinite(a)) and worst <= 1.0)
    return ok, worst


def scalar_close(a, b, scale, rtol=RTOL, atol_frac=ATOL_FRAC):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ratio = np.abs(a - b) / (rtol * np.abs(b) + atol_frac * scale)
    worst = float(np.nanmax(ratio)) if ratio.size else 0.0
    return bool(np.all(np.isfinite(a)) and worst <= 1.0), worst

"""Influent generator: the step before the hot path (buffer_tank3.influent.buffer_tank, buffer_tank3.py:13-1197).

Eight scenarios (`switch` 0..7: dry / carbon-rich / high-load / N-rich x morning / evening) of 48-point mean
profiles for the 13 concentrations and the flow; each draw perturbs every profile with ONE shared
rnd ~ N(0,1)^48 (std = 0.1*mean for Ss, Xi, Xs, Xbh, Snh, Snd, Xnd and q, 0 for the rest, buffer_tank3.py:50-66)
and returns the flow-weighted mean concentrations, `influent_mixed = [0.66, sum(c*q)/sum(q) ...]`
(buffer_tank3.py:87-107).  Scenario 0 consumes one randn(48) from the RNG, scenarios 1..7 consume two and use
the second (buffer_tank3.py:206,224).  The mean profiles are data, extracted from the reference by
oracle/extract_influent_tables.py into data/influent_tables.npz.  `SBR-v0` / `SBR-v1` draw from buffer_tank2 instead
(96-point profiles with explicit standard deviations; see `mix_numpy_bt2` below).

`mix_numpy` is bit-exact with the reference (sequential sums, same operation order) and serves the single-env
Gym wrappers, where "identical seeds" means np.random.seed(s) before reset().  The vector envs draw on the device
(sbr_influent_sample / sbr_influent_mix, same arithmetic in the same order).
"""
import os

import numpy as np

_TABLES = None
N_POINTS = 48


def tables():
    global _TABLES
    if _TABLES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "influent_tables.npz")
        z = np.load(path)
        _TABLES = dict(mean=z["mean"], std_frac=z["std_frac"], draws=z["draws"])
    return _TABLES


def draws_per_reset(switch):
    """How many randn(48) vectors one buffer_tank(switch) call consumes (the last one is used)."""
    return int(tables()["draws"][int(switch)])


def mix_numpy(switch, rnd):
    """influent_mixed (14-vector, [0] = 0.66) for one env from one rnd[48]; bit-exact with the reference."""
    t = tables()
    mean, frac = t["mean"][int(switch)], t["std_frac"][int(switch)]
    rnd = np.asarray(rnd, dtype=np.float64)
    prof = np.empty((14, N_POINTS))
    for j in range(14):
        prof[j] = mean[j] + (frac[j] * mean[j]) * rnd if frac[j] != 0 else mean[j] + 0 * rnd
    q = prof[0]
    qsum = np.cumsum(q)[-1]                      # sequential, like the builtin sum() the reference uses
    out = np.empty(14)
    out[0] = 0.66
    for j in range(1, 14):
        out[j] = np.cumsum(prof[j] * q)[-1] / qsum
    return out


def sample_numpy(switch, rng=None):
    """One buffer_tank(switch) call: consumes draws_per_reset(switch) x randn(48) from `rng` (default: the
    global numpy RNG, as the reference does) and returns influent_mixed."""
    rng = np.random if rng is None else rng
    rnd = None
    for _ in range(draws_per_reset(switch)):
        rnd = rng.randn(N_POINTS)
    return mix_numpy(switch, rnd)


def sample_numpy_random_scenario(rng=None):
    """SbrEnv4.reset's draw (gym_SBR_env4.py:104): buffer_tank(np.random.choice(8, 1)).  Consumes the RNG exactly like
    the reference: one choice(8, 1), then the scenario's randn(48) draws.  Returns (switch, influent_mixed)."""
    rng = np.random if rng is None else rng
    switch = int(rng.choice(8, 1)[0])
    return switch, sample_numpy(switch, rng)


# ---------------------------------------------------------------------------------------------------------
# buffer_tank2.influent.buffer_tank(0, 12) -- the influent of `SBR-v0` / `SBR-v1` (gym_SBR_env0.py:74,208).  Its `switch`
# is drawn and then forced to 1 (buffer_tank2.py:16-18); that branch perturbs 96-point profiles with their own
# standard-deviation profiles and ONE shared rnd ~ N(0,1)^96 -- the flow with the OPPOSITE sign, q = q_m + q_s (-rnd)
# (:246-261) -- and returns the flow-weighted means over the first 48 points (hours 0..12 of 24, :264-303).
# ---------------------------------------------------------------------------------------------------------
BT2_POINTS = 96


def tables_bt2():
    """(mean [14,48], std [14,48]) of the points buffer_tank2 uses, row 0 = flow with the sign of its perturbation folded
    into std (so that every row is mean + std * rnd, the form sbr_influent_mix / sbr_influent_sample evaluate)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "influent_tables.npz")
    z = np.load(path)
    mean, std = z["bt2_mean"][:, :N_POINTS].copy(), z["bt2_std"][:, :N_POINTS].copy()
    std[0] = -std[0]
    return mean, std


def mix_numpy_bt2(rnd):
    """influent_mixed (14-vector, [0] = 0.66) from one rnd[96] (only its first 48 entries matter); bit-exact with the
    reference (sequential sums, same operation order)."""
    mean, std = tables_bt2()
    rnd = np.asarray(rnd, dtype=np.float64)[:N_POINTS]
    q = mean[0] + (-std[0]) * (-rnd)                # q_m + q_s * (-rnd), buffer_tank2.py:261
    qsum = np.cumsum(q)[-1]
    out = np.empty(14)
    out[0] = 0.66
    for j in range(1, 14):
        out[j] = np.cumsum((mean[j] + std[j] * rnd) * q)[-1] / qsum
    return out


def sample_numpy_bt2(rng=None):
    """One buffer_tank2 call: consumes np.random.choice(2, 1) and randn(96) from `rng` (default: the global numpy RNG, as
    the reference does) and returns influent_mixed."""
    rng = np.random if rng is None else rng
    rng.choice(2, 1)
    return mix_numpy_bt2(rng.randn(BT2_POINTS))

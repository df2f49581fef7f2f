"""Host-side logic that needs no GPU: schedule tables, influent generator, C-ABI exports, argument errors."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gym_sbr2_b200 import _abi, influent, schedule
from oracle import sbr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schedule_matches_reference_counts(golden_v2):
    s = schedule.cycle_schedule()
    assert list(s.n_int) == [24, 48, 223, 186, 11, 0, 0, 36]
    assert list(s.n_sub) == [9, 9, 9, 9, 10, 0, 0, 9]
    assert (s.n_int[2], s.n_int[4], s.n_int[7]) == (golden_v2["n3"][0], golden_v2["n5"][0], golden_v2["n8"][0])
    # same grid as the oracle (which is pinned to the reference)
    for k, (t0, t1) in enumerate(O.cycle_schedule()):
        if k in (5, 6):
            continue
        grid, pts = O.phase_grid(t0, t1)
        assert len(grid) - 1 == s.n_int[k] and set(pts) == {s.n_sub[k] + 1}
        assert np.isclose(s.interval[k], grid[1] - grid[0], rtol=1e-12)
    assert np.isclose(s.settle_time, 0.5 * 0.083, rtol=1e-12)


def test_influent_generator_bit_exact_with_reference(golden_v2):
    g = golden_v2
    for i in range(0, len(g["seed"]), 3):
        np.random.seed(int(g["seed"][i]))
        assert np.array_equal(influent.sample_numpy(0), g["influent"][i])


def test_influent_draw_counts():
    assert [influent.draws_per_reset(k) for k in range(8)] == [1, 2, 2, 2, 2, 2, 2, 2]


def _header_functions():
    text = open(os.path.join(ROOT, "include", "sbr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sbr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = _header_functions()
    assert names == _abi.exported_symbols()
    lib = C.CDLL(_abi.lib_path())
    for n in names:
        assert hasattr(lib, n), n
    assert _abi.load().sbr_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_header(built):
    text = open(os.path.join(ROOT, "include", "sbr_b200.h")).read()
    body = re.search(r"typedef struct SbrParams \{(.*?)\} SbrParams;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"double\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f for f, _ in _abi.SbrParams._fields_]
    assert C.sizeof(_abi.SbrSchedule) == 8 * 4 + 8 * 4 + 8 * 8 + 8
    assert C.sizeof(_abi.SbrTol) == 24
    assert C.sizeof(_abi.SbrIlcLayout) == 13 * 4
    body = re.search(r"typedef struct SbrIlcLayout \{(.*?)\} SbrIlcLayout;", text, flags=re.S).group(1)
    assert [m for m in re.findall(r"int32_t\s+(\w+)", body)] == [f for f, _ in _abi.SbrIlcLayout._fields_]
    body = re.search(r"typedef struct SbrOsSchedule \{(.*?)\} SbrOsSchedule;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:double|int32_t)\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f for f, _ in _abi.SbrOsSchedule._fields_]
    assert C.sizeof(_abi.SbrOsSchedule) == 10 * 8 + 4 * 4
    rows = re.search(r"enum \{\s*SBR_OS_X = 0,(.*?)SBR_OS_ROWS", text, flags=re.S).group(1)
    assert (_abi.OS_T, _abi.OS_KLA_RING, _abi.OS_RETURN, _abi.OS_QW, _abi.OS_ROWS) == (14, 22, 32, 34, 35)
    assert rows.count("SBR_OS_") == 13


def test_default_params_match_reference_constants(built):
    p = _abi.default_params()
    assert p.so_sat == O.SO_SAT
    for k, v in O.KPAR.items():
        assert getattr(p, k) == v
    for k, v in O.SPAR.items():
        assert getattr(p, k) == v
    assert (p.pid_Kc, p.pid_tauI, p.pid_tauD, p.pid_dt) == (5.0, 0.00035, 0.005, 0.02 / 24)
    assert p.Qin == O.QIN and p.Qeff == 0.66 and p.biomass_setpoint == 2700 and p.ec_conc == 4800000.0
    from oracle.twin import binding as twin
    q = twin.default_params()
    for f, _ in _abi.SbrParams._fields_:
        assert getattr(p, f) == getattr(q, f), f


def test_argument_errors_are_reported_without_gpu(built):
    lib = _abi.load()
    p = _abi.default_params()
    s = schedule.cycle_schedule()
    rc = lib.sbr_cycle_v2(0, 0, None, None, None, C.byref(p), C.byref(s), None, None, None, None, None, None, 0,
                          None, None, None)
    assert rc == -1 and b"n must be positive" in lib.sbr_last_error()
    rc = lib.sbr_cycle_v2(4, 2, None, None, None, C.byref(p), C.byref(s), None, None, None, None, None, None, 0,
                          None, None, None)
    assert rc == -1 and b"ld" in lib.sbr_last_error()
    rc = lib.sbr_cycle_v2(4, 4, None, None, None, C.byref(p), C.byref(s), None, None, None, None, None, None, 0,
                          None, None, None)
    assert rc == -1 and b"NULL" in lib.sbr_last_error()
    with pytest.raises(_abi.SbrLibraryError):
        _abi.check(rc, "sbr_cycle_v2")
    # the batch-to-batch and trajectory entry points validate before they launch, too
    from gym_sbr2_b200 import ilc
    lay = ilc.layout(s)
    rc = lib.sbr_cycle_ilc(4, 4, None, None, None, C.byref(p), C.byref(s), C.byref(lay), 0.021, None, None, None, None,
                           None, None, None, None, 1, None, None)
    assert rc == -1 and b"NULL" in lib.sbr_last_error()
    rc = lib.sbr_ilc_update(4, 2, C.byref(lay), None, None, None, None, None, None, None, 1e-4, 1.0, 0.25, 0.1, None)
    assert rc == -1 and b"ld" in lib.sbr_last_error()
    assert lib.sbr_cycle_v2_traj_records(C.byref(s)) == 529
    rc = lib.sbr_cycle_v2_traj(4, 4, None, None, None, C.byref(p), C.byref(s), None, None, None, None, None, None, None,
                               None, 0, None, None)
    assert rc == -1 and b"NULL" in lib.sbr_last_error()


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle, and the vector env refuses non-CUDA devices."""
    import gym_sbr2_b200
    pkg = os.path.dirname(gym_sbr2_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert not re.search(r"^\s*(from|import)\s+scipy", src, flags=re.M), fn
    from gym_sbr2_b200.vec_env import SbrV2VecEnv
    with pytest.raises(_abi.SbrLibraryError):
        SbrV2VecEnv(4, device="cpu")


def test_cpu_baseline_workers_survive_the_reference_failure_regimes(monkeypatch):
    """The reference's own failure regimes raise (round(inf) in the draw, gym_SBR_oneshot.py:2338); a CPU baseline
    worker must count such a step and go on -- a crash there once cost the whole bench line."""
    from oracle import cpu_baseline, sbr_oracle as O
    calls = {"n": 0}
    real_step = O.SbrOsOracle.step

    def flaky_step(self, action):
        calls["n"] += 1
        if calls["n"] % 3 == 0:
            raise OverflowError("cannot convert float infinity to integer")
        return real_step(self, action)

    monkeypatch.setattr(O.SbrOsOracle, "step", flaky_step)
    dt, acc = cpu_baseline._worker_os((0, 7))
    assert calls["n"] == 7 and dt > 0 and np.isfinite(acc)

    def broken_cycle(action, infl):
        raise ValueError("oracle left the physical regime")

    monkeypatch.setattr(O, "sbr_v2_step", broken_cycle)
    dt, acc = cpu_baseline._worker((0, 2))
    assert dt >= 0 and acc == 0.0

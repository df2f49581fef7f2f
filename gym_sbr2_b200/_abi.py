"""ctypes binding of the C ABI declared in include/sbr_b200.h (the CUDA library libsbr_b200.so).

There is no CPU fallback: if the library is missing this module raises at first use, and every compute entry
point fails when no CUDA device is present.  Struct layouts mirror include/sbr_b200.h field for field.
"""
import ctypes as C
import os

NX = 14
NPHASE = 8
MODE_RK4, MODE_DP45 = 0, 1
TAIL_REACT, TAIL_FILL, TAIL_EC = 0, 1, 2
ST_NONFINITE, ST_WASTE, ST_STEPLIMIT, ST_LAYERS, ST_DONE = 1, 2, 4, 8, 16
AUX_ROWS = 12
PERMUTE_MAX = 8
AUX_NAMES = ("OCI", "Qw", "EQI", "eff_Q", "eff_Ntot", "eff_COD", "eff_Snh", "eff_BOD5", "eff_Sno",
             "kla3_mean", "kla5_mean", "kla8_mean")
ABI_VERSION = 8
# rows of the persistent per-env state of the interval-per-step path (enum SBR_OS_* in include/sbr_b200.h)
OS_X, OS_T, OS_SO_PREV, OS_SNO_LAST, OS_SNO_PREV, OS_IE_DO, OS_IE_EC, OS_EC_LAST, OS_H, OS_KLA_RING = \
    0, 14, 15, 16, 17, 18, 19, 20, 21, 22
OS_RETURN, OS_STEPS, OS_QW, OS_ROWS = 32, 33, 34, 35
OS_NOBS, OS_NSTATE = 9, 15
# rows of one trajectory record (enum SBR_TRAJ_*)
TRAJ_T, TRAJ_X, TRAJ_KLA, TRAJ_EC, TRAJ_U_DO, TRAJ_U_EC, TRAJ_REWARD, TRAJ_EQI, TRAJ_OCI, TRAJ_AE, TRAJ_ECO, TRAJ_ROWS = \
    0, 1, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24
# rows of one record of sbr_cycle_v2_traj (enum SBR_TRAJ2_*)
TRAJ2_T, TRAJ2_X, TRAJ2_KLA, TRAJ2_ROWS = 0, 1, 15, 16
# rows of the persistent state of the SBR-v4 env (enum SBR_V4_*)
V4_T, V4_U, V4_SO_PREV, V4_IE, V4_KLA_LAST, V4_KLA_SUM, V4_H, V4_RETURN, V4_STEPS, V4_QW, V4_ROWS = \
    14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24

_PARAM_FIELDS = [
    "muh", "Ks", "Koh", "Kno", "bh", "etag", "etah", "kh", "Kx", "mua", "Knh", "ba", "Koa", "ka",
    "Ya", "Yh", "fp", "ixb", "ixp", "so_sat",
    "pid_Kc", "pid_tauI", "pid_tauD", "pid_dt", "kla_min", "kla_max",
    "WV", "Qin", "Qeff", "biomass_setpoint", "settler_area", "settler_vmax", "kla0", "action_scale",
    "os_Kc_DO", "os_tauI_DO", "os_tauD_DO", "os_Kc_EC", "os_tauI_EC", "os_tauD_EC",
    "os_pid_dt", "ec_min", "ec_max", "ec_conc", "do_sp_max", "no_sp_max", "IV",
]


class SbrParams(C.Structure):
    _fields_ = [(name, C.c_double) for name in _PARAM_FIELDS]


class SbrSchedule(C.Structure):
    _fields_ = [("n_int", C.c_int32 * NPHASE), ("n_sub", C.c_int32 * NPHASE),
                ("interval", C.c_double * NPHASE), ("settle_time", C.c_double)]


class SbrOsSchedule(C.Structure):
    _fields_ = [("tm3_0", C.c_double), ("tm3_1", C.c_double), ("tm4_1", C.c_double), ("tm5_1", C.c_double),
                ("dt", C.c_double), ("t_delta", C.c_double), ("t_fill", C.c_double),
                ("settle_len", C.c_double), ("draw_len", C.c_double), ("t_cycle", C.c_double),
                ("fill_pts", C.c_int32), ("rk4_sub_interval", C.c_int32), ("rk4_sub_fill", C.c_int32),
                ("rk4_sub_idle", C.c_int32)]


class SbrTol(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("max_steps", C.c_int32), ("flags", C.c_int32)]


class SbrCntConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32),
                ("Kc_DO", C.c_double), ("tauI_DO", C.c_double), ("tauD_DO", C.c_double),
                ("Kc_EC", C.c_double), ("tauI_EC", C.c_double), ("tauD_EC", C.c_double),
                ("ec_conc", C.c_double), ("ec_fill_max", C.c_double), ("u_ec_init", C.c_double),
                ("u_ec_max", C.c_double), ("tm2_0", C.c_double), ("tm2_1", C.c_double), ("tm4_0", C.c_double)]


class SbrPolicyMlp(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("w2", C.c_void_p), ("lo", C.c_void_p), ("span", C.c_void_p),
                ("n_in", C.c_int32), ("hidden", C.c_int32), ("n_out", C.c_int32), ("reserved", C.c_int32)]


# kinds of the sbr_cnt_* entry points (enum SBR_CNT_* of include/sbr_b200.h) and their persistent-state rows
CNT_V0, CNT_V1, CNT_V2, CNT_MA1, CNT_OS2 = range(5)
(CNT_T, CNT_U_DO, CNT_U_EC, CNT_SO_PREV, CNT_CV_LAST, CNT_CV_PREV, CNT_IE_DO, CNT_IE_EC, CNT_KLA_LAST, CNT_EC_LAST,
 CNT_H, CNT_RETURN, CNT_STEPS, CNT_QW, CNT_ROWS) = range(14, 29)
CNT_NOBS_MAX = 33


class SbrIlcLayout(C.Structure):
    _fields_ = [("off", C.c_int32 * 6), ("n_samples", C.c_int32), ("tp", C.c_int32 * 6)]


# rows of sbr_cycle_ilc's `out` (enum SBR_ILC_*)
ILC_QEFF, ILC_QW, ILC_REWARD, ILC_OCI, ILC_KLA3_MEAN, ILC_KLA5_MEAN, ILC_KLA8_MEAN, ILC_OUT_ROWS = range(8)


class SbrLibraryError(RuntimeError):
    pass


_LIB = None
LIB_NAME = "libsbr_b200.so"


def lib_path():
    """In-tree library; SBR_B200_LIB overrides it (A/B builds of the same ABI during kernel tuning)."""
    return os.environ.get("SBR_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


_P = C.c_void_p
_PROTOS = {
    "sbr_abi_version": (C.c_int, []),
    "sbr_last_error": (C.c_char_p, []),
    "sbr_device_count": (C.c_int, []),
    "sbr_params_default": (None, [C.POINTER(SbrParams)]),
    "sbr_cycle_v2": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrSchedule),
                               _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P, _P]),
    "sbr_cycle_v2_traj_records": (C.c_int, [C.POINTER(SbrSchedule)]),
    "sbr_cycle_v2_traj": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrSchedule),
                                    C.POINTER(C.c_double), _P, _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_integrate_interval": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(SbrParams), C.c_int,
                                         C.c_double, C.c_int, C.c_int, C.POINTER(SbrTol), _P, _P]),
    "sbr_rhs": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(SbrParams), C.c_int, _P, _P]),
    "sbr_settle_draw": (C.c_int, [C.c_int64, C.c_int64, _P, C.c_double, C.POINTER(SbrParams), _P, _P, _P, _P]),
    "sbr_os_reset": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrOsSchedule),
                               _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_os_step": (C.c_int, [C.c_int64, C.c_int64, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrOsSchedule),
                              _P, _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_os_step_k": (C.c_int, [C.c_int64, C.c_int64, C.c_int, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrOsSchedule),
                                _P, _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_os_step_traj": (C.c_int, [C.c_int64, C.c_int64, C.c_int, _P, _P, C.POINTER(SbrParams),
                                   C.POINTER(SbrOsSchedule), _P, _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P,
                                   C.c_int, _P]),
    "sbr_v4_reset": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(SbrParams), _P, _P, _P, _P, _P]),
    "sbr_v4_step": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrOsSchedule),
                              _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P, _P]),
    "sbr_cnt_obs_rows": (C.c_int, [C.c_int]),
    "sbr_cnt_reset": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(SbrCntConfig), _P, _P, _P, C.POINTER(SbrParams),
                                C.POINTER(SbrOsSchedule), _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_cnt_step": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(SbrCntConfig), _P, _P, C.POINTER(SbrParams),
                               C.POINTER(SbrOsSchedule), _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_influent_mix": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, _P, _P]),
    "sbr_influent_sample": (C.c_int, [C.c_int64, C.c_int64, C.c_uint64, C.c_int64, _P, C.c_int64, C.c_int, _P, _P,
                                      _P, _P, _P, _P]),
    "sbr_philox_normals": (C.c_int, [C.c_int64, C.c_int64, C.c_uint64, C.c_int64, C.c_int64, _P, _P]),
    "sbr_permute_rows": (C.c_int, [C.c_int64, _P, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.c_int, _P]),
    "sbr_os_rollout_k": (C.c_int, [C.c_int64, C.c_int64, C.c_int, _P, _P, C.POINTER(SbrPolicyMlp), C.POINTER(SbrParams),
                                   C.POINTER(SbrOsSchedule), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int,
                                   C.POINTER(SbrTol), _P]),
    "sbr_v4_rollout_k": (C.c_int, [C.c_int64, C.c_int64, C.c_int, _P, _P, _P, C.POINTER(SbrPolicyMlp), C.POINTER(SbrParams),
                                   C.POINTER(SbrOsSchedule), _P, _P, _P, _P, _P, _P, _P, C.c_int, C.POINTER(SbrTol), _P]),
    "sbr_cnt_rollout_k": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.POINTER(SbrCntConfig), _P, _P, C.POINTER(SbrPolicyMlp),
                                    C.POINTER(SbrParams), C.POINTER(SbrOsSchedule), _P, _P, _P, _P, _P, _P, _P, C.c_int,
                                    C.POINTER(SbrTol), _P]),
    "sbr_policy_mlp": (C.c_int, [C.c_int64, C.c_int64, _P, C.c_int, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "sbr_cycle_ilc": (C.c_int, [C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(SbrParams), C.POINTER(SbrSchedule),
                                C.POINTER(SbrIlcLayout), C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int,
                                C.POINTER(SbrTol), _P]),
    "sbr_ilc_update": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(SbrIlcLayout), _P, _P, _P, _P, _P, _P, _P, C.c_double,
                                 C.c_double, C.c_double, C.c_double, _P]),
    "sbr_reward_stats_init": (C.c_int, [_P, _P]),
    "sbr_reward_stats": (C.c_int, [C.c_int64, _P, _P, _P, _P]),
    "sbr_fp64_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_double), _P]),
}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite against include/sbr_b200.h)."""
    return sorted(_PROTOS)


def load():
    """Load libsbr_b200.so (built in-tree by __graft_entry__.build()).  Raises if it is missing or stale."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise SbrLibraryError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise SbrLibraryError("%s does not export %s (stale build?)" % (path, name))
        fn.restype = res
        fn.argtypes = args
    if lib.sbr_abi_version() != ABI_VERSION:
        raise SbrLibraryError("ABI version mismatch: library %d, binding %d" % (lib.sbr_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().sbr_last_error()
        raise SbrLibraryError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def default_params():
    p = SbrParams()
    load().sbr_params_default(C.byref(p))
    return p


FLAG_RAW_KLA = 1


def make_tol(rtol=1e-8, atol=1e-10, max_steps=200, flags=0):
    return SbrTol(float(rtol), float(atol), int(max_steps), int(flags))

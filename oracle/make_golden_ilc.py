"""Golden vectors of the reference's batch-to-batch (ILC) feed-forward KLa path (`SBR-v0`), authoring container only.

DISCLOSURE.  `SbrEnv.step` of gym_SBR_env0.py cannot run as shipped, for two independent reasons:
  1. sub_phases_batchPID_fbPID.py passes a FLOAT `num` to np.linspace in nine places (:144,183,209,393,394,434,461,727,793),
     TypeError on numpy >= 1.18.  The file dates from the numpy < 1.18 era, where linspace truncated with int().  This
     harness runs the UNMODIFIED source with that behaviour restored FOR THAT MODULE ONLY (the module-level name `np`
     is replaced by a proxy whose `linspace` casts `num` with int()) -- the same shim as oracle/make_golden_v4.py.
  2. gym_SBR_env0.py:203 calls module_reward.sbr_reward (ten parameters) with seven arguments: TypeError, and no
     function with that signature exists anywhere in the reference.  Nothing is repaired here: the fixtures stop
     where `step()` would compute the reward.  What is recorded is everything `step()` does before that line --
     `_take_action` (module_batch_PID.batch_PID on the module's memories) and `_next_observation`
     (SBR_model_batchPID_fbPID.run) -- through the env's own methods ("env" chain), and the same two reference
     functions called directly with the So / set-point memories of the PREVIOUS cycle fed back ("learning" chain:
     `step()` assigns the new memories to locals, gym_SBR_env0.py:200, so the env itself learns from cycle 0 for ever).
No reference source is copied or edited.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden_v4 import _NumpyPre118  # noqa: E402

NAMES = ("1", "2", "3", "4", "5", "8")


def load_env0(seed):
    np.random.seed(seed)              # buffer_tank2 draws the module's influent at import (gym_SBR_env0.py:74)
    ref_shim.load_reference()
    import gym_SBR.envs.gym_SBR_env0 as m
    import gym_SBR.envs.sub_phases_batchPID_fbPID as sp
    sp.np = _NumpyPre118()
    return m


# per fixture: the seed of the module-level influent draw and the three actions of the two chains.  Fixture 1 touches the
# edges of the action box [0, 5]; a zero set-point is only given to the env chain: in the learning chain it would make the
# NEXT set-point memory 0 / 0 = NaN (gym_SBR_env0.py:251-253), a state the module itself never reaches
FIXTURES = {
    0: (np.array([[2.0, 2.5, 1.5], [1.0, 3.0, 2.0], [3.5, 0.5, 4.0]]), None),
    1: (np.array([[0.0, 5.0, 0.3], [4.8, 0.05, 5.0], [0.6, 2.2, 0.0]]),
        np.array([[0.05, 5.0, 0.3], [4.8, 0.05, 5.0], [0.6, 2.2, 0.05]])),
}


def main(out_dir, seed=0):
    import scipy
    m = load_env0(seed)
    from gym_SBR.envs.module_batch_PID import batch_PID
    from gym_SBR.envs import SBR_model_batchPID_fbPID as SBR
    g = dict(versions=np.array([np.__version__, scipy.__version__, sys.version.split()[0]]))
    g["influent"] = np.array(m.influent_mixed, dtype=float)
    g["x0"] = np.array(m.x0, dtype=float)
    g["x_last0"] = np.array(m.x_last, dtype=float)
    g["par_batchPID"] = np.array(m.par_batchPID, dtype=float)
    g["do_control_par"] = np.array(m.DO_control_par, dtype=float)
    t_mem = [m.t_memory1, m.t_memory2, m.t_memory3, m.t_memory4, m.t_memory5, m.t_memory8]
    so0 = [m.So_memory1, m.So_memory2, m.So_memory3, m.So_memory4, m.So_memory5, m.So_memory8]
    sp0 = [m.sp_memory1, m.sp_memory2, m.sp_memory3, m.sp_memory4, m.sp_memory5, m.sp_memory8]
    kla0 = [m.kla_memory_1_1, m.kla_memory_2_1, m.kla_memory_3_1, m.kla_memory_4_1, m.kla_memory_5_1, m.kla_memory_8_1]
    for n, t, s, p, k in zip(NAMES, t_mem, so0, sp0, kla0):
        g["t_memory" + n] = np.array(t, dtype=float)
        g["So0_" + n] = np.array(s, dtype=float)
        g["sp0_" + n] = np.array(p, dtype=float)
        g["kla0_" + n] = np.array(k, dtype=float)
    actions, actions_learn = FIXTURES[seed]
    actions_learn = actions if actions_learn is None else actions_learn
    g["actions"], g["actions_learn"] = actions, actions_learn

    def run_cycle(x_last, sp_set, u_rows):
        with ref_shim.quiet():
            return SBR.run(m.WV, m.IV, m.t_ratio, list(m.influent_mixed), list(m.DO_control_par), list(x_last), list(sp_set),
                           *[u[-1, :] for u in u_rows], *kla0)

    # ---- "learning" chain: the two reference functions, the new memories fed back (computed FIRST: the env chain below
    # mutates the module's globals) ----
    e_mem = [np.zeros((1, len(t))) for t in t_mem]
    u_rows = [np.zeros((1, len(t))) for t in t_mem]
    so_prev, sp_prev = [np.array(s, dtype=float) for s in so0], [np.array(p, dtype=float) for p in sp0]
    x_last = list(m.x_last)
    for c, a in enumerate(actions_learn):
        sp_in = list(sp_prev)
        sp_in[2] = sp_prev[2] / sp_prev[2][0] * a[0]
        sp_in[4] = sp_prev[4] / sp_prev[4][0] * a[1]
        sp_in[5] = sp_prev[5] / sp_prev[5][0] * a[2]
        out = batch_PID(m.par_batchPID, *t_mem, m.t_delta, *so_prev, *sp_in, *e_mem, *u_rows)
        u_rows, e_mem = list(out[:6]), list(out[6:])
        r = run_cycle(x_last, [0, 0, a[0], 0, a[1], 0, 0, a[2]], u_rows)
        x_last = r[2]
        sp_new = [r[4], r[7], r[10], r[13], r[16], r[19]]
        so_new = [r[5], r[8], r[11], r[14], r[17], r[20]]
        kla_new = list(r[21:27])
        for j, n in enumerate(NAMES):
            g["learn_c%d_E%s" % (c, n)] = np.array(e_mem[j][-1])
            g["learn_c%d_u%s" % (c, n)] = np.array(u_rows[j][-1])
            g["learn_c%d_So%s" % (c, n)] = np.array(so_new[j], dtype=float)
            g["learn_c%d_Kla%s" % (c, n)] = np.array(kla_new[j], dtype=float)
        g["learn_c%d_x_last" % c] = np.array(x_last, dtype=float)
        g["learn_c%d_Qeff_Qw" % c] = np.array([r[27], r[28]], dtype=float)
        so_prev, sp_prev = [np.array(s, dtype=float) for s in so_new], [np.array(p, dtype=float) for p in sp_new]
        print("learning cycle %d: x_last[8..10] %s Qw %.6g max|u3| %.4g" % (c, np.array(x_last)[8:11], r[28],
                                                                          np.abs(u_rows[2][-1]).max()), flush=True)

    # ---- "env" chain: the env's own methods up to the reward call ----
    env = m.SbrEnv()
    with ref_shim.quiet():
        g["reset_obs"] = np.array(env.reset(), dtype=float)
    for c, a in enumerate(actions):
        a = np.clip(a, env.action_space.low, env.action_space.high)
        m.influent_mixed[0] = 31.4285
        with ref_shim.quiet():
            env._take_action(a)
            r = env._next_observation(m.WV, m.IV, m.t_ratio, m.influent_mixed, m.DO_control_par, m.x_last, m.DO_setpoints,
                                      m.u_batch_1, m.u_batch_2, m.u_batch_3, m.u_batch_4, m.u_batch_5, m.u_batch_8,
                                      m.kla_memory_1_1, m.kla_memory_2_1, m.kla_memory_3_1, m.kla_memory_4_1,
                                      m.kla_memory_5_1, m.kla_memory_8_1)
        m.x_last = r[2]                                  # `global x_last` in step() (gym_SBR_env0.py:191,200)
        u_now = [m.u_batch_1, m.u_batch_2, m.u_batch_3, m.u_batch_4, m.u_batch_5, m.u_batch_8]
        e_now = [m.memory_e_batch_1, m.memory_e_batch_2, m.memory_e_batch_3, m.memory_e_batch_4, m.memory_e_batch_5,
                 m.memory_e_batch_8]
        for j, n in enumerate(NAMES):
            g["env_c%d_E%s" % (c, n)] = np.array(e_now[j][-1])
            g["env_c%d_u%s" % (c, n)] = np.array(u_now[j][-1])
        g["env_c%d_So3" % c] = np.array(r[11], dtype=float)
        g["env_c%d_x_last" % c] = np.array(r[2], dtype=float)
        g["env_c%d_Qeff_Qw" % c] = np.array([r[27], r[28]], dtype=float)
        print("env cycle %d: x_last[8..10] %s Qw %.6g max|u3| %.4g" % (c, np.array(r[2])[8:11], r[28],
                                                                     np.abs(u_now[2][-1]).max()), flush=True)
    # ---- `SBR-v1` (gym_SBR_env1.py): the same plant without the batch-to-batch controller.  Its step() dies on the same
    # seven-argument reward call (:151); before it, _take_action writes the set-points and _next_observation runs
    # SBR_model_FBc_implemented.run (= SBR_model_PID_on.run) from the carried-over state.  Unmodified, no shim needed. ----
    import gym_SBR.envs.gym_SBR_env1 as m1
    env1 = m1.SbrEnv1()
    with ref_shim.quiet():
        g["v1_reset_obs"] = np.array(env1.reset(), dtype=float)
    g["v1_influent"] = np.array(m1.influent_mixed, dtype=float)
    x1 = list(m1.x)
    for c, a in enumerate(actions):
        a = np.clip(a, env1.action_space.low, env1.action_space.high)
        m1.influent_mixed[0] = 31.4285
        with ref_shim.quiet():
            env1._take_action(a)
            r = env1._next_observation(m1.WV, m1.IV, m1.t_ratio, m1.influent_mixed, m1.DO_control_par, x1, m1.DO_setpoints)
        x1 = r[2]                                        # `x = x_last` (gym_SBR_env1.py:158)
        g["v1_c%d_x_last" % c] = np.array(r[2], dtype=float)
        g["v1_c%d_Qeff_Qw" % c] = np.array([r[27], r[28]], dtype=float)
        g["v1_c%d_kla3" % c] = np.array(r[29], dtype=float)
        print("v1 cycle %d: x_last[8..10] %s Qw %.6g" % (c, np.array(r[2])[8:11], r[28]), flush=True)
    # ---- the influent source of both envs: three consecutive buffer_tank2.influent.buffer_tank(0, 12) calls on a seeded
    # global numpy RNG (unmodified) ----
    from gym_SBR.envs import buffer_tank2
    np.random.seed(123)
    g["bt2_seed123_draws"] = np.array([buffer_tank2.influent.buffer_tank(0, 12)[1] for _ in range(3)], dtype=float)
    np.savez_compressed(os.path.join(out_dir, "ilc_seed%d.npz" % seed), **g)
    print("lengths", [len(t) for t in t_mem])


if __name__ == "__main__":
    # one fixture per process: the reference runs its cycle 0 at import
    main(os.path.join(os.path.dirname(HERE), "tests", "golden"), int(sys.argv[1]) if len(sys.argv) > 1 else 0)

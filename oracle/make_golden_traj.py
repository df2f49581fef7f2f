"""Golden trajectory of the UNMODIFIED reference `SBROS-v1` env: SbrOS.trajectory() (gym_SBR_oneshot.py:1275-1288)
after a whole seed-0 episode with constant set-points, reduced to the samples at the END of every env.step (the
reference lists 9-10 output points per PID interval; the product records interval ends).  Authoring container only:
    python oracle/make_golden_traj.py        -> tests/golden/sbros_v1_traj_seed0_const.npz"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    ref_shim.load_reference()
    import gym_SBR.envs.gym_SBR_oneshot as m
    np.random.seed(0)
    env = m.SbrOS()
    action = [2.0, 5.0]
    idx, rewards = [], []
    with ref_shim.quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env.reset()
        influent = np.array(m.influent_mixed, dtype=float)
        n0 = len(m.t_t)
        while True:
            obs, state, reward, done, info = env.step(action)
            idx.append(len(m.t_t) - 1)
            rewards.append(float(reward))
            if done:
                break
        tup = env.trajectory()
    names = ("t_t", "x_t", "u_DO_t", "u_EC_t", "state_t", "So_t", "Ss_t", "EC", "Sno_t", "dcv_EC", "ie_EC", "e_EC",
             "reward_t", "reward_EQI_t", "reward_OCI_t", "reward_AE_t", "reward_EC_t", "Snh_t")
    tr = dict(zip(names, tup))
    for k, v in tr.items():
        print(k, np.shape(v))
    t_t, x_t = np.array(tr["t_t"], dtype=float), np.array(tr["x_t"], dtype=float)
    idx = np.array(idx)
    out = dict(influent=influent, action=np.array(action), n_fill_points=n0, step_end_index=idx,
               t_step_end=t_t[idx], x_step_end=x_t[idx], ec_step_end=np.array(tr["EC"], dtype=float)[idx],
               u_do_step_end=np.array(tr["u_DO_t"], dtype=float)[idx], u_ec_step_end=np.array(tr["u_EC_t"], dtype=float)[idx],
               ie_EC=np.array(tr["ie_EC"], dtype=float), state_t_len=len(tr["state_t"]), t_fill_end=t_t[n0 - 1], x_fill_end=x_t[n0 - 1],
               reward_t=np.array(tr["reward_t"], dtype=float), reward_EQI_t=np.array(tr["reward_EQI_t"], dtype=float),
               reward_OCI_t=np.array(tr["reward_OCI_t"], dtype=float), reward_AE_t=np.array(tr["reward_AE_t"], dtype=float),
               reward_EC_t=np.array(tr["reward_EC_t"], dtype=float), step_rewards=np.array(rewards),
               tuple_names=np.array(names), n_points=len(t_t),
               versions=np.array(["numpy " + np.__version__, "scipy " + __import__("scipy").__version__]))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "sbros_v1_traj_seed0_const.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "steps", len(idx), "points", len(t_t))


if __name__ == "__main__":
    main()

"""GPU parity tests of the batch-to-batch (ILC) feed-forward KLa path of `SBR-v0` -- all calls through the C ABI
(sbr_cycle_ilc, sbr_ilc_update) and `SbrIlcVecEnv`.  Fixtures: outputs of the reference's own functions
(tests/golden/ilc_seed0.npz; shim and scope disclosed in oracle/make_golden_ilc.py)."""
import os

import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, ilc, parity, schedule
from oracle.twin import binding as twin

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ilc_seed%d.npz")
NAMES = ("1", "2", "3", "4", "5", "8")
SO_RTOL, SO_ATOL, KLA_ATOL = 1e-5, 1e-6, 2e-4          # see tests/test_twin_parity_ilc.py


@pytest.fixture(scope="module", params=[0, 1], ids=["seed0", "seed1_box_edges"])
def g(request):
    """Fixture 0: mid-range set-points; fixture 1: another influent draw and set-points at the edges of the action box."""
    return np.load(GOLDEN % request.param, allow_pickle=True)


def cat(g, prefix):
    return np.concatenate([g[prefix + n] for n in NAMES])


def dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(device)


def col(a, n, device):
    return dev(np.tile(np.asarray(a, dtype=float)[:, None], (1, n)), device)


def test_cycle0_and_feed_forward_cycles_match_reference(built, cuda_device, g):
    p = ilc.apply_constants(_abi.default_params())
    sched = schedule.cycle_schedule()
    w, D, lay = ilc.weights(sched)
    n = 3                                   # the same env three times: every lane must give the same bits
    r0 = ilc.cycle_ilc(col(g["x0"], n, cuda_device), col(g["influent"], n, cuda_device), col([2.0, 2.0, 2.0], n, cuda_device),
                       p, sched, lay)
    x0 = r0.x_last.cpu().numpy()
    assert np.array_equal(x0[:, 0], x0[:, 2])
    ok, worst = parity.state_close(x0[:, 0], g["x_last0"])
    assert ok, worst
    assert np.allclose(r0.so_mem.cpu().numpy()[:, 0], cat(g, "So0_"), rtol=SO_RTOL, atol=SO_ATOL)
    assert np.allclose(r0.kla_mem.cpu().numpy()[:, 0], cat(g, "kla0_"), rtol=1e-5, atol=KLA_ATOL)
    assert int(r0.status.abs().sum()) == 0
    for c in range(3):
        a = g["actions_learn"][c]
        x_in = g["x_last0"] if c == 0 else g["learn_c%d_x_last" % (c - 1)]
        r = ilc.cycle_ilc(col(x_in, n, cuda_device), col(g["influent"], n, cuda_device), col(a, n, cuda_device), p, sched, lay,
                          kla_base=col(cat(g, "kla0_"), n, cuda_device), u=col(cat(g, "learn_c%d_u" % c), n, cuda_device))
        ok, worst = parity.state_close(r.x_last.cpu().numpy()[:, 1], g["learn_c%d_x_last" % c])
        assert ok, (c, worst)
        assert np.allclose(r.so_mem.cpu().numpy()[:, 1], cat(g, "learn_c%d_So" % c), rtol=SO_RTOL, atol=SO_ATOL)
        assert np.array_equal(r.kla_mem.cpu().numpy()[:, 1], cat(g, "learn_c%d_Kla" % c))
        assert np.allclose(r.out.cpu().numpy()[:2, 1], g["learn_c%d_Qeff_Qw" % c], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("chain", ["env", "learn"])
def test_update_kernel_matches_reference(built, cuda_device, g, chain):
    sched = schedule.cycle_schedule()
    w, D, lay = ilc.weights(sched)
    S, n = int(lay.n_samples), 2
    wd, Dd = dev(w, cuda_device), dev(D, cuda_device)
    e_sum = torch.zeros((S, n), dtype=torch.float64, device=cuda_device)
    e_last, u = torch.zeros_like(e_sum), torch.zeros_like(e_sum)
    so = col(cat(g, "So0_"), n, cuda_device)
    for c, a in enumerate(g["actions" if chain == "env" else "actions_learn"]):
        sp6 = col([0, 0, a[0], 0, a[1], a[2]], n, cuda_device)
        ilc.ilc_update(lay, wd, Dd, sp6, so, e_sum, e_last, u)
        assert np.allclose(e_last.cpu().numpy()[:, 1], cat(g, "%s_c%d_E" % (chain, c)), rtol=1e-9, atol=1e-12), c
        assert np.allclose(u.cpu().numpy()[:, 0], cat(g, "%s_c%d_u" % (chain, c)), rtol=1e-9, atol=1e-11), c
        if chain == "learn":
            so = col(cat(g, "learn_c%d_So" % c), n, cuda_device)


@pytest.mark.parametrize("learn", ["frozen", "feedback"])
def test_vec_env_closed_loop_matches_the_reference_chain(built, cuda_device, g, learn):
    """SbrIlcVecEnv.reset() + three step()s with the module's influent held: end states, u_batch and observation of
    the reference's env chain (memories frozen at cycle 0) and of the learning chain."""
    chain = "env" if learn == "frozen" else "learn"
    n = 4
    env = ilc.SbrIlcVecEnv(n, device=cuda_device, seed=0, learn=learn, record_feed_forward=True)
    infl = np.tile(g["influent"][None, :], (n, 1))
    obs = env.reset(influent=infl)
    # reset observation (gym_SBR_env0.py:150-176): (x_last + influent) / scale, first entry 1
    assert np.allclose(obs.cpu().numpy()[0], g["reset_obs"], rtol=1e-5, atol=1e-7)
    for c, a in enumerate(g["actions" if chain == "env" else "actions_learn"]):
        act = np.tile(a[None, :], (n, 1))
        obs, reward, done, info = env.step(act, influent=infl)
        ok, worst = parity.state_close(info["x_last"].cpu().numpy()[:, 3], g["%s_c%d_x_last" % (chain, c)], rtol=3e-5)
        assert ok, (c, worst)
        assert np.allclose(info["u_batch"].cpu().numpy()[:, 0], cat(g, "%s_c%d_u" % (chain, c)), rtol=1e-4, atol=2e-4), c
        assert np.allclose([float(info["Qeff"][0]), float(info["Qw"][0])], g["%s_c%d_Qeff_Qw" % (chain, c)], rtol=1e-5)
        if chain == "learn":       # the clamped feed-forward profile, written on request only
            assert np.allclose(info["kla_ff"].cpu().numpy()[:, 2], cat(g, "learn_c%d_Kla" % c), rtol=1e-4, atol=2e-4)
        assert bool(done.all()) and bool(torch.isfinite(reward).all()) and info["reward_pinned"] is False
        assert obs.shape == (n, 14) and float(obs[0, 0]) == 1.0
        assert int(info["status"].abs().sum()) == 0


def test_batch_against_cpu_twin_and_action_clipping(built, cuda_device):
    """512 envs with per-env influent and random actions (some outside [0, 5]): every env of the GPU batch against the g++
    build of the same source, the update kernel's recursion included."""
    n = 512
    rng = np.random.RandomState(3)
    env = ilc.SbrIlcVecEnv(n, device=cuda_device, seed=11, learn="feedback")
    env.reset()
    infl0 = env.influent.cpu().numpy().copy()
    base, so0, x0 = env.kla_base.cpu().numpy(), env.so_learn.cpu().numpy(), env.x.cpu().numpy()
    act = rng.uniform(-0.5, 5.5, (n, 3))
    act[:, :] = np.where(np.abs(act) < 0.05, 0.3, act)
    obs, reward, done, info = env.step(act)
    p = ilc.apply_constants(twin.default_params())
    sched = schedule.cycle_schedule()
    w, D, lay = ilc.weights(sched)
    t_fill = schedule.T_CYCLE * schedule.T_RATIO[0]
    # cycle 0 on the twin from the same influent
    xi = np.tile(np.array(ilc.X0_ILC)[:, None], (1, n))
    r0 = twin.cycle_ilc(xi, infl0, np.full((3, n), 2.0), p, sched, lay, t_fill)
    assert np.allclose(r0["kla_mem"], base, rtol=1e-7, atol=1e-7) and np.allclose(r0["so_mem"], so0, rtol=1e-7, atol=1e-10)
    assert np.allclose(r0["x_last"], x0, rtol=1e-9, atol=1e-12)
    a = np.clip(act, 0.0, 5.0).T
    sp6 = np.zeros((6, n)); sp6[2], sp6[4], sp6[5] = a[0], a[1], a[2]
    e_sum, e_last = np.zeros_like(so0), np.zeros_like(so0)
    u = twin.ilc_update(lay, w, D, sp6, so0, e_sum, e_last, schedule.T_DELTA, ilc.KC_B, ilc.TAUI_B, ilc.TAUD_B)
    assert np.allclose(info["u_batch"].cpu().numpy(), u, rtol=1e-9, atol=1e-10)
    r = twin.cycle_ilc(x0, infl0, a, p, sched, lay, t_fill, kla_base=base, u=u)
    ok, worst = parity.state_close(info["x_last"].cpu().numpy().T, r["x_last"].T, rtol=1e-7)
    assert ok, worst
    assert np.allclose(info["so_mem"].cpu().numpy(), r["so_mem"], rtol=1e-6, atol=1e-9)


def test_argument_errors(built, cuda_device):
    p = ilc.apply_constants(_abi.default_params())
    sched = schedule.cycle_schedule()
    w, D, lay = ilc.weights(sched)
    n = 2
    x = col(ilc.X0_ILC, n, cuda_device)
    infl = torch.zeros((14, n), dtype=torch.float64, device=cuda_device)
    sp = torch.zeros((3, n), dtype=torch.float64, device=cuda_device)
    with pytest.raises(_abi.SbrLibraryError):       # kla_base without u
        ilc.cycle_ilc(x, infl, sp, p, sched, lay, kla_base=torch.zeros((lay.n_samples, n), dtype=torch.float64,
                                                                      device=cuda_device))
    short = schedule.cycle_schedule(substeps=12)    # more samples than the layout holds
    with pytest.raises(_abi.SbrLibraryError):
        ilc.cycle_ilc(x, infl, sp, p, short, lay)


def test_checkpoint_resume_is_bitwise(built, cuda_device):
    """state_dict() after two steps, two more steps; a fresh env loaded from the checkpoint repeats them bit for bit
    (controller memories, plant state and the influent stream are all part of the checkpoint)."""
    n = 64
    g = torch.Generator(device="cpu").manual_seed(2)
    acts = [(torch.rand((n, 3), generator=g, dtype=torch.float64) * 4 + 0.5).to(cuda_device) for _ in range(4)]
    env = ilc.SbrIlcVecEnv(n, device=cuda_device, seed=5, learn="feedback")
    env.reset()
    for a in acts[:2]:
        env.step(a)
    sd = env.state_dict()
    ref = [env.step(a) for a in acts[2:]]
    ref = [(o.clone(), r.clone(), i["x_last"].clone(), i["u_batch"].clone()) for o, r, d, i in ref][-1]
    env2 = ilc.SbrIlcVecEnv(n, device=cuda_device, seed=99, learn="frozen")
    env2.load_state_dict(sd)
    for a in acts[2:]:
        o, r, d, i = env2.step(a)
    assert torch.equal(o, ref[0]) and torch.equal(r, ref[1]) and torch.equal(i["x_last"], ref[2])
    assert torch.equal(i["u_batch"], ref[3])


def test_sbr_v1_vec_env_matches_the_reference_chain(built, cuda_device, g):
    """SbrV1VecEnv: reset observation of the module's initial state, then three chained feedback-PID cycles
    (SBR_model_FBc_implemented.run through gym_SBR_env1's own methods) with the module's influent held."""
    n = 3
    env = ilc.SbrV1VecEnv(n, device=cuda_device, seed=0)
    infl = np.tile(g["v1_influent"][None, :], (n, 1))
    obs = env.reset(influent=infl)
    assert np.allclose(obs.cpu().numpy()[1], g["v1_reset_obs"], rtol=1e-12)
    for c, a in enumerate(g["actions"]):
        obs, reward, done, info = env.step(np.tile(a[None, :], (n, 1)), influent=infl)
        ok, worst = parity.state_close(info["x_last"].cpu().numpy()[:, 2], g["v1_c%d_x_last" % c], rtol=2e-5)
        assert ok, (c, worst)
        assert np.allclose([float(info["Qeff"][0]), float(info["Qw"][0])], g["v1_c%d_Qeff_Qw" % c], rtol=1e-5)
        assert abs(float(info["kla3_mean"][1]) - g["v1_c%d_kla3" % c][1:].mean()) < 1e-3
        assert bool(done.all()) and int(info["status"].abs().sum()) == 0 and bool(torch.isfinite(reward).all())


def test_shards_reproduce_the_single_batch_run_bit_for_bit(built, cuda_device):
    """the path shards by env index with no collective; per-env influent draws are keyed by the GLOBAL env
    index, so two shards (env_offset) give exactly what one batch gives, through reset + two steps with drawn influent."""
    n, cut = 192, 80
    gen = torch.Generator(device="cpu").manual_seed(4)
    acts = [(torch.rand((n, 3), generator=gen, dtype=torch.float64) * 4 + 0.5).to(cuda_device) for _ in range(2)]
    full = ilc.SbrIlcVecEnv(n, device=cuda_device, seed=21, learn="feedback")
    parts = [ilc.SbrIlcVecEnv(cut, device=cuda_device, seed=21, learn="feedback", env_offset=0),
             ilc.SbrIlcVecEnv(n - cut, device=cuda_device, seed=21, learn="feedback", env_offset=cut)]
    o_full = full.reset()
    o_parts = torch.cat([parts[0].reset(), parts[1].reset()], dim=0)
    assert torch.equal(o_full, o_parts)
    for a in acts:
        of, rf, _, inf_f = full.step(a)
        res = [parts[0].step(a[:cut]), parts[1].step(a[cut:])]
        assert torch.equal(of, torch.cat([r[0] for r in res], dim=0))
        assert torch.equal(rf, torch.cat([r[1] for r in res], dim=0))
        assert torch.equal(inf_f["u_batch"], torch.cat([r[3]["u_batch"] for r in res], dim=1))
        assert torch.equal(inf_f["x_last"], torch.cat([r[3]["x_last"] for r in res], dim=1))


def test_device_sampler_serves_buffer_tank2_bit_for_bit(built, cuda_device):
    """The counter-based sampler with the buffer_tank2 tables (the influent of SbrIlcVecEnv / SbrV1VecEnv) equals the
    reference's arithmetic (influent.mix_numpy_bt2, pinned bit-exactly to buffer_tank2 in the CPU suite) on its own normals."""
    from gym_sbr2_b200 import core, influent
    n, seed, off = 300, 17, 1000
    z = core.philox_normals(n, cuda_device, seed, env_offset=off, epoch0=2).cpu().numpy()          # [48, n]
    got = core.influent_sample(n, cuda_device, seed, env_offset=off, scenario=0, epoch0=2,
                               table_set="buffer_tank2").cpu().numpy()
    for i in range(0, n, 37):
        ref = influent.mix_numpy_bt2(np.concatenate([z[:, i], np.zeros(48)]))
        assert np.array_equal(got[:, i], ref), i
    env = ilc.SbrIlcVecEnv(8, device=cuda_device, seed=seed, env_offset=off)
    env.reset()
    first = core.influent_sample(8, cuda_device, seed, env_offset=off, scenario=0, epoch0=0, table_set="buffer_tank2")
    assert torch.equal(env.influent[1:], first[1:]) and float(env.influent[0, 0]) == ilc.FILL_FLOW
    with pytest.raises(ValueError):
        core.influent_sample(8, cuda_device, seed, scenario=3, table_set="buffer_tank2")


def test_zero_setpoints_do_not_poison_the_set_point_memory(built, cuda_device):
    """gym_SBR_env0.py:251-253 rescales the previous set-point memory, sp_prev / sp_prev[0] * action.  In the module the
    previous memory is cycle 0's for ever (set-point 2), so the expression is the action itself; with the last cycle fed back
    (learn="feedback") it would become 0 / 0 = NaN after a zero set-point -- a state the reference cannot reach.  Both modes
    take the action: a zero set-point followed by a non-zero one stays finite."""
    n = 4
    a0 = torch.tensor([[0.0, 2.0, 2.0]] * n, dtype=torch.float64, device=cuda_device)
    a1 = torch.tensor([[1.0, 0.0, 2.0]] * n, dtype=torch.float64, device=cuda_device)
    for learn in ("feedback", "frozen"):
        env = ilc.SbrIlcVecEnv(n, device=cuda_device, seed=3, learn=learn)
        env.reset()
        for a in (a0, a1, a0):
            _, reward, _, info = env.step(a)
            assert int(info["status"].abs().sum()) == 0 and bool(torch.isfinite(info["x_last"]).all())
            assert bool(torch.isfinite(info["u_batch"]).all()) and bool(torch.isfinite(reward).all())


def test_batch_too_large_for_the_device_is_refused_before_allocating(built, cuda_device):
    free, _ = torch.cuda.mem_get_info(cuda_device)
    n = int(free / (6 * 4769 * 8)) + 4096
    with pytest.raises(ValueError, match="sample memories"):
        ilc.SbrIlcVecEnv(n, device=cuda_device, seed=0)


def test_sbr_v1_ordered_launch_is_bit_identical(built, cuda_device):
    """SbrV1VecEnv hands its envs to the adaptive kernel sorted by set-point (gather in, scatter out): same bits as the
    launch in caller order, over chained cycles with drawn influent."""
    n = 500
    gen = torch.Generator(device="cpu").manual_seed(8)
    acts = [(torch.rand((n, 3), generator=gen, dtype=torch.float64) * 5).to(cuda_device) for _ in range(3)]
    a, b = ilc.SbrV1VecEnv(n, device=cuda_device, seed=4, order="action"), ilc.SbrV1VecEnv(n, device=cuda_device, seed=4, order="none")
    assert torch.equal(a.reset(), b.reset())
    for act in acts:
        oa, ra, _, ia = a.step(act)
        ob, rb, _, ib = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ia["x_last"], ib["x_last"])
        assert torch.equal(ia["counters"], ib["counters"]) and torch.equal(ia["status"], ib["status"])

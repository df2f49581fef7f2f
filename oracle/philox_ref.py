"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the counter-based influent randomness of sbr_influent_sample
(gym_sbr2_b200/csrc/sbr_kernels.cu): Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; the Random123 known-answer vectors pin it, tests/test_philox_ref.py), counter = (global env index lo, hi,
episode number, block), key = the run's 64-bit seed; two 53-bit uniforms per block -> Box-Muller -> two normals.
The reference itself draws from numpy's global Mersenne Twister (buffer_tank3.py:206,224): a per-process stream has no
batched, shard-invariant counterpart, so the product's default generator is its own and only the MIXING arithmetic is
bit-identical with the reference (rng="numpy" reproduces the reference's stream for identical-seed comparisons)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: [..., 4] uint32, key: (k0, k1) ints -> [..., 4] uint32."""
    c = [np.asarray(ctr[..., j], dtype=np.uint64) for j in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        n0 = (p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1)
        c = [n0 & MASK, p1 & MASK, n2 & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def _counters(env, epoch, block):
    env = np.asarray(env, dtype=np.uint64)
    ctr = np.empty(env.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (env & MASK).astype(np.uint32)
    ctr[..., 1] = (env >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = np.uint32(int(epoch) & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32(block)
    return ctr


def normals(seed, env, epoch, points=48):
    """z [points, n]: the standard normals env `env[i]` draws for its episode number `epoch`."""
    env = np.asarray(env, dtype=np.int64)
    key = (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    z = np.empty((points, env.shape[0]))
    for b in range(points // 2):
        r = philox4x32_10(_counters(env, epoch, b), key).astype(np.uint64)
        u1 = (((r[:, 1] << np.uint64(32)) | r[:, 0]) >> np.uint64(11)).astype(np.float64)
        u2 = (((r[:, 3] << np.uint64(32)) | r[:, 2]) >> np.uint64(11)).astype(np.float64)
        u1 = (u1 + 0.5) * (1.0 / 9007199254740992.0)
        u2 = (u2 + 0.5) * (1.0 / 9007199254740992.0)
        rad = np.sqrt(-2.0 * np.log(u1))
        z[2 * b] = rad * np.cos(2.0 * np.pi * u2)
        z[2 * b + 1] = rad * np.sin(2.0 * np.pi * u2)
    return z


def scenario(seed, env, epoch):
    """SbrEnv4's per-reset scenario (np.random.choice(8, 1), gym_SBR_env4.py:104) as the sampler draws it."""
    env = np.asarray(env, dtype=np.int64)
    key = (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return (philox4x32_10(_counters(env, epoch, 24), key)[:, 0] >> np.uint32(29)).astype(np.int32)

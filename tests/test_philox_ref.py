"""The numpy restatement of the influent sampler's counter-based generator (oracle/philox_ref.py) against the
Random123 known-answer vectors for Philox4x32-10, and its basic statistics.  CPU only."""
import numpy as np

from oracle import philox_ref as P


def test_philox4x32_10_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = P.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in got) == want


def test_normals_are_standard_and_keyed_by_env_and_epoch():
    env = np.arange(20000)
    z = P.normals(1234, env, 0)
    assert z.shape == (48, 20000) and abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 4) - 3.0) < 0.05                               # kurtosis of a normal
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.03                        # the two Box-Muller outputs
    # depends only on (seed, global env index, epoch): shards reproduce the full batch
    assert np.array_equal(P.normals(1234, env[7000:7100], 0), z[:, 7000:7100])
    assert not np.array_equal(P.normals(1234, env[:100], 1), z[:, :100])
    assert not np.array_equal(P.normals(1235, env[:100], 0), z[:, :100])
    s = P.scenario(1234, env, 0)
    assert s.min() == 0 and s.max() == 7 and np.all(np.abs(np.bincount(s) / len(s) - 0.125) < 0.01)

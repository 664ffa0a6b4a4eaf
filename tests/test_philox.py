"""CPU: the oracle's Philox4x32-10 restatement against the known-answer vectors published with Random123
(kat_vectors: philox4x32 10 rounds), and sanity of the uniform -> normal mapping the device generator uses."""
import numpy as np

from oracle import philox


def test_philox4x32_10_known_answers():
    kat = [
        ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
         (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
         (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array([ctr], dtype=np.uint64), key)[0]
        assert tuple(int(v) for v in got) == want


def test_normals_are_standard():
    z = philox.randn(200001, seed=12345, stream_id=3)
    assert z.shape == (200001,)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1.0) < 0.01
    assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3.0) < 0.06
    # different streams / seeds are different sequences; same (seed, stream) is reproducible
    assert np.array_equal(z[:100], philox.randn(100, 12345, 3))
    assert not np.allclose(z[:100], philox.randn(100, 12345, 4))
    assert not np.allclose(z[:100], philox.randn(100, 12346, 3))

"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device normal generator (gpcsd_randn, csrc/gpcsd_aux.cu).

Philox4x32-10 is the counter-based generator of Salmon, Moraes, Dror & Shaw, "Parallel random numbers: as easy as 1, 2, 3"
(SC'11); the known-answer vectors in tests/test_philox.py are the ones published with the Random123 library.  The reference
(np.random.normal, gpcsd1d.py:308) uses numpy's global Mersenne Twister, whose stream a device generator cannot replay
(SURVEY.md section 9.8): parity for sample_prior is therefore (i) arithmetic parity given the same normals and (ii) this
bit-exact restatement of the device stream.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (n, 4) uint32-valued array, key: (2,) -> (n, 4) uint32 outputs."""
    c = [np.asarray(ctr)[:, k].astype(np.uint64) & MASK for k in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def raw(ncounters, seed, stream_id=0):
    """Outputs for counters (p, p >> 32, stream_id, 0), p = 0..ncounters-1, key = (seed lo, seed hi)."""
    p = np.arange(ncounters, dtype=np.uint64)
    ctr = np.stack([p & MASK, p >> np.uint64(32), np.full(ncounters, stream_id, dtype=np.uint64), np.zeros(ncounters, dtype=np.uint64)], axis=1)
    return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def randn(n, seed, stream_id=0):
    """n standard normals: element i is normal (i & 1) of counter i >> 1 (Box-Muller on two 53-bit uniforms)."""
    r = raw((n + 1) // 2, seed, stream_id).astype(np.float64)
    u1 = (np.floor(r[:, 0] / 32.0) * 67108864.0 + np.floor(r[:, 1] / 64.0) + 0.5) / 9007199254740992.0
    u2 = (np.floor(r[:, 2] / 32.0) * 67108864.0 + np.floor(r[:, 3] / 64.0) + 0.5) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.stack([rad * np.cos(2.0 * np.pi * u2), rad * np.sin(2.0 * np.pi * u2)], axis=1).reshape(-1)
    return z[:n]

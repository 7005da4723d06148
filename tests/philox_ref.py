"""Philox4x32-10 in numpy: the checker for codegen/cuda_kernel.uniform_philox (test infrastructure only).

Pinned by the known-answer vectors of Random123 (kat_vectors, philox4x32 10 rounds) in tests/test_host_logic.py."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: four uint32 arrays (or scalars) of equal shape, key: two python ints.  Returns four uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in counter]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = key[0] & 0xFFFFFFFF, key[1] & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def uniform(numel: int, seed: int, call: int, low: float = 0.0, high: float = 1.0) -> np.ndarray:
    """What uniform_philox writes for draw number `call` under `seed`: thread t -> elements 4t .. 4t+3."""
    t = np.arange(-(-numel // 4), dtype=np.uint64)
    out = philox4x32_10((t, 0, call & 0xFFFFFFFF, call >> 32), (seed & 0xFFFFFFFF, seed >> 32))
    u = np.stack([(x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24) for x in out], axis=1).reshape(-1)[:numel]
    low, span = np.float32(low), np.float32(high) - np.float32(low)
    return (low + span * u).astype(np.float32)

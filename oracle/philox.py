"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based generator the injector
uses (Philox4x32-10, Salmon et al. 2011, as in Random123): checked against the published
known-answer vectors (tests/test_inject.py) and used to predict the injector's per-cell
decisions on the GPU (tests/test_gpu_inject.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(ctr, key, rounds=10):
    """ctr: uint32 array [..., 4], key: uint32 array [..., 2] -> uint32 array [..., 4]"""
    c = [np.asarray(ctr[..., k], np.uint32).copy() for k in range(4)]
    k0 = np.asarray(key[..., 0], np.uint32).copy()
    k1 = np.asarray(key[..., 1], np.uint32).copy()
    for _ in range(rounds):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        with np.errstate(over="ignore"):
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def first_uniform(seed, step, call, cells):
    """the first Random<real_t> of every cell's stream: key = seed, counter = (0, cell, call, step);
    the device hands out the four words of a block last to first"""
    cells = np.asarray(cells, np.uint32)
    ctr = np.zeros(cells.shape + (4,), np.uint32)
    ctr[..., 1], ctr[..., 2], ctr[..., 3] = cells, np.uint32(call), np.uint32(step)
    key = np.zeros(cells.shape + (2,), np.uint32)
    key[..., 0], key[..., 1] = np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)
    w = philox4x32(ctr, key)[..., 3]
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(5.9604645e-08)

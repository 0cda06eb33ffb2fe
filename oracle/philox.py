"""Counter-based corruption mask (rng_mode='philox'), numpy restatement of k_batch_prep's generator.

Not a reference behaviour: this is the documented deviation for throughput configurations (the reference's
MT19937 stream needs n_items draws per sampled user).  Philox4x32-10 (Salmon et al., SC'11) with
counter = (item, slot, step_lo, step_hi) and key = (seed_lo, seed_hi); keep iff u >= q, u = (x0 >> 8) * 2^-24.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox_first(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(x, np.uint32) for x in (c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0


def keep_mask(items, slot, step, seed, q):
    """keep flags for the items of batch slot `slot` at optimizer step `step`."""
    items = np.asarray(items, np.uint32)
    x = philox_first(items, np.full_like(items, slot), np.full_like(items, step & 0xFFFFFFFF),
                     np.full_like(items, step >> 32), seed & 0xFFFFFFFF, seed >> 32)
    u = (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return ~(u < np.float32(q))

"""Host replica of the counter-based dropout mask of the fused attention kernels (csrc/flash_attn.cu,
csrc/common.cuh: pcm_row_seed per row, one pcm_pair_bits hash per group of 8 keys advanced by a
32-bit LCG per key): element (z, l, s) of the (B*nh, L, S) probability tensor is KEPT iff the top
16 bits of its value are >= round(p * 65536)."""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def keep_mask(seed_base: int, seed_offset: int, Z: int, L: int, S: int, p: float) -> np.ndarray:
    thr16 = np.uint32(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))
    with np.errstate(over="ignore"):
        seed = np.uint64(seed_base & 0xFFFFFFFFFFFFFFFF) * np.uint64(0xD1342543DE82EF95) + np.uint64(seed_offset & 0xFFFFFFFFFFFFFFFF)
        rows = np.arange(Z * L, dtype=np.uint64)
        x = seed + rows * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
        rseed = ((x >> np.uint64(32)) ^ (x & np.uint64(0xFFFFFFFF))).astype(np.uint32)
        groups = (np.arange((S + 7) // 8, dtype=np.uint32))
        y = rseed[:, None] ^ (groups[None, :] * np.uint32(0x9E3779B9))
        y ^= y >> np.uint32(16); y *= np.uint32(0x7FEB352D)
        y ^= y >> np.uint32(15); y *= np.uint32(0x846CA68B)
        y ^= y >> np.uint32(16)
        # key k of a group: the group hash advanced k times by the LCG x -> x * A + C
        cols = []
        for _ in range(8):
            cols.append(y.copy())
            y = y * np.uint32(0x2C9277B5) + np.uint32(0x9E3779B9)
        x = np.stack(cols, axis=-1).reshape(Z * L, -1)[:, :S]
    return ((x >> np.uint32(16)) >= thr16).reshape(Z, L, S)


def keep_scale(p: float) -> float:
    thr16 = int(np.float32(p) * np.float32(65536.0) + np.float32(0.5))
    return 65536.0 / (65536.0 - thr16)

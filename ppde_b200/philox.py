"""Host mirror of the device random streams (Philox4x32-10, counter based).

The reference draws its randomness from three global generators
(`torch.randint` on the CPU generator for the path length, `torch.multinomial`
and `torch.rand_like` on the device generator: ppde/protein_samplers/ppde.py:67,
109,138).  Global generator state cannot be shared between a lock-step PyTorch
program and fused kernels, so this build defines *indexed* streams instead:
every random number is a pure function of

    (seed, iteration t, sub-step s, global chain id b, entry j, stream kind)

and is therefore independent of how chains are sharded over GPUs.  The CUDA
kernels (csrc/philox.cuh) and this numpy mirror produce bit-identical uint32
words; the parity tests feed the very same words to the reference through
patched `torch.randint` / `torch.multinomial` / `torch.rand_like`.

Counter layout (4 x uint32):  c0 = j // 4   (each call yields 4 entries)
                              c1 = global chain id
                              c2 = iteration t
                              c3 = sub-step s | (kind << 16)
Key (2 x uint32):             seed low / high word.
"""
import numpy as np

KIND_PROPOSAL = 0   # exponential race over the [20L] proposal entries
KIND_PATHLEN = 1    # U ~ {1 .. 2*pas-1}
KIND_ACCEPT = 2     # Metropolis-Hastings uniform

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds. All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)))
    c0 = c0.astype(np.uint64); c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64); c3 = c3.astype(np.uint64)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0
            p1 = _M1 * c2
            hi0 = p0 >> np.uint64(32); lo0 = p0 & _MASK
            hi1 = p1 >> np.uint64(32); lo1 = p1 & _MASK
            n0 = hi1 ^ c1 ^ np.uint64(k0)
            n2 = hi0 ^ c3 ^ np.uint64(k1)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def u32_to_unit(x):
    """uint32 -> float32 strictly inside (0,1): (x>>8 + 0.5) * 2^-24 (exact in fp32)."""
    x = np.asarray(x, dtype=np.uint32)
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def proposal_uniforms(seed, t, s, chain_ids, n_entries):
    """float32 [len(chain_ids), n_entries] uniforms for iteration t, sub-step s."""
    k0, k1 = _key(seed)
    chain_ids = np.asarray(chain_ids, dtype=np.uint32)
    nblk = (n_entries + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)[None, :]
    w = philox4x32_10(blk, chain_ids[:, None], np.uint32(t),
                      np.uint32(s | (KIND_PROPOSAL << 16)), k0, k1)
    u = np.stack([u32_to_unit(x) for x in w], axis=-1).reshape(len(chain_ids), nblk * 4)
    return np.ascontiguousarray(u[:, :n_entries])


def path_lengths(seed, t, chain_ids, pas_length):
    """int32 [n] path lengths U in {1 .. 2*pas-1} (reference: ppde.py:67)."""
    k0, k1 = _key(seed)
    chain_ids = np.asarray(chain_ids, dtype=np.uint32)
    w = philox4x32_10(np.uint32(0), chain_ids, np.uint32(t),
                      np.uint32(KIND_PATHLEN << 16), k0, k1)[0]
    span = 2 * int(pas_length) - 1
    return (1 + (w % np.uint32(span))).astype(np.int32)


def accept_uniforms(seed, t, chain_ids):
    """float32 [n] Metropolis-Hastings uniforms (reference: ppde.py:138)."""
    k0, k1 = _key(seed)
    chain_ids = np.asarray(chain_ids, dtype=np.uint32)
    w = philox4x32_10(np.uint32(0), chain_ids, np.uint32(t),
                      np.uint32(KIND_ACCEPT << 16), k0, k1)[0]
    return u32_to_unit(w)

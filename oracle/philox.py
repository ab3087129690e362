"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the in-kernel noise source of
tsdiff_b200/csrc/ld_step.cu: Philox4x32-10 (Salmon et al., SC'11; Random123 constants)
keyed by (seed, step, global atom id) followed by Box-Muller on 24-bit uniforms.  The
reference itself draws torch.randn_like from the global generator (sampler.py:213); the
keyed stream is this framework's production replacement so that results do not depend on
how reactions are sharded over GPUs."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter: (n, 4) uint32, key: (2,) uint32 -> (n, 4) uint32."""
    c = counter.astype(np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[:, 0].astype(np.uint64)
            p1 = M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c


def normals(num_nodes, seed, step, atom_offset=0):
    """(num_nodes, 3) float32 standard normals, as tsd_philox_normal3 produces them."""
    atom = np.arange(num_nodes, dtype=np.uint64) + np.uint64(atom_offset)
    ctr = np.stack([(atom & np.uint64(0xFFFFFFFF)).astype(np.uint32), (atom >> np.uint64(32)).astype(np.uint32),
                    np.full(num_nodes, step, dtype=np.uint32), np.zeros(num_nodes, dtype=np.uint32)], axis=1)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32))
    u = ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    two_pi = np.float32(6.283185307179586)
    ra = np.sqrt(np.float32(-2.0) * np.log(u[:, 0]))
    rb = np.sqrt(np.float32(-2.0) * np.log(u[:, 2]))
    return np.stack([ra * np.cos(two_pi * u[:, 1]), ra * np.sin(two_pi * u[:, 1]), rb * np.cos(two_pi * u[:, 3])],
                    axis=1).astype(np.float32)

"""Host copy of the counter RNG addressing (Philox4x32-10, Salmon et al. SC'11) used by the
kernels (csrc/device_math.cuh).  Only the camera jitter of SPPMRayTracer::render
(Raytracer.h:434) is drawn on the host; everything else is drawn on the device."""
from __future__ import annotations

import math

import numpy as np

P_EMIT, P_BOUNCE, P_JITTER = 0, 1, 4
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = (int(x) & M32 for x in ctr)
    k0, k1 = (int(x) & M32 for x in key)
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def _u53(a, b):
    return (float(a >> 5) * 67108864.0 + float(b >> 6)) * (1.0 / 9007199254740992.0)


def draw2(seed, pass_id, ident, node, purpose, sub, patch):
    ctr = (ident & M32, ((ident >> 32) & 0xFFFF) | ((node << 16) & M32),
           (purpose & 0xF) | ((sub & 0xFFF) << 4) | ((patch << 16) & M32), pass_id & M32)
    o = philox4x32_10(ctr, (seed & M32, (seed >> 32) & M32))
    return _u53(o[0], o[1]), _u53(o[2], o[3])


def unit_rand_vec(seed, pass_id, ident, purpose):
    """Vec3::GetUnitRandVec (Vec3.h:57-65) with addressed draws."""
    a = 0
    while True:
        p, q = draw2(seed, pass_id, ident, 0, purpose, 2 * a, 0)
        r, _ = draw2(seed, pass_id, ident, 0, purpose, 2 * a + 1, 0)
        x, y, z = p * 2.0 + -1.0, q * 2.0 + -1.0, r * 2.0 + -1.0
        r2 = (x * x + y * y) + z * z
        a += 1
        if not ((r2 > 1 or r2 < 1e-4) and a < 2000):
            break
    mo = math.sqrt(r2)
    v = np.array([x, y, z])
    return v if abs(mo) < 1e-4 else v * (1.0 / mo)

"""R's default RNG restated (TEST INFRASTRUCTURE ONLY -- see oracle/README.md).

Mersenne-Twister + Inversion, exactly as R's src/main/RNG.c and src/nmath/{snorm,rbinom}.c
define them (R >= 3.6 defaults).  Used only to regenerate the README's seed-42 example of the
reference (README.Rmd:35-39 `withr::local_seed(42)`, data at README.Rmd:44-55) so that the
oracle can be pinned to the reference's own printed output (README.md:73-120).
"""
import numpy as np
from scipy.special import ndtri

_N, _M = 624, 397
_I2_32M1 = 2.328306437080797e-10


class RRng:
    def __init__(self, seed: int):
        # RNG.c:RNG_Init -- initial scrambling, then LCG-fill of dummy[0..624]; FixupSeeds sets mti=N
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            dummy = np.empty(_N + 1, dtype=np.uint32)
            for j in range(_N + 1):
                s = np.uint32(69069) * s + np.uint32(1)
                dummy[j] = s
        self.mt = [int(v) for v in dummy[1:]]
        self.mti = _N  # dummy[0] = 624
        self.n_unif = 0

    def _genrand(self) -> int:
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N):
                y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % _N] & 0x7FFFFFFF)
                v = mt[(kk + _M) % _N] ^ (y >> 1)
                if y & 1:
                    v ^= 0x9908B0DF
                mt[kk] = v
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def unif_rand(self) -> float:
        self.n_unif += 1
        v = self._genrand() * 2.3283064365386963e-10  # [0,1)
        # RNG.c:fixup -- keep strictly inside (0,1)
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def norm_rand(self) -> float:
        # snorm.c INVERSION: 2 uniforms -> 53-bit u -> qnorm
        BIG = 134217728.0
        u = self.unif_rand()
        u = int(BIG * u) + self.unif_rand()
        return float(ndtri(u / BIG))

    def runif(self, n):
        return np.array([self.unif_rand() for _ in range(n)])

    def rnorm(self, n, mean=0.0, sd=1.0):
        mean = np.broadcast_to(np.asarray(mean, dtype=float), (n,))
        return np.array([mean[i] + sd * self.norm_rand() for i in range(n)])

    def rbinom_size1(self, n, p=0.5):
        # rbinom.c, n*p < 30 branch (inverse cdf): size=1 -> qn=q, g=2r, one uniform per draw
        q = 1.0 - min(p, 1.0 - p)
        pp = min(p, 1.0 - p)
        r = pp / q
        g = r * 2.0
        out = np.empty(n)
        for i in range(n):
            while True:
                ix, f, u = 0, q, self.unif_rand()
                done = False
                while True:
                    if u < f:
                        done = True
                        break
                    if ix > 110:
                        break
                    u -= f
                    ix += 1
                    f *= (g / ix - r)
                if done:
                    break
            out[i] = (1 - ix) if p > 0.5 else ix
        return out

"""R's default RNG (R 3.6.0: Mersenne-Twister, Inversion, Rejection sampling).

Test infrastructure (see ``oracle/__init__.py``).  Needed because the
reference's known-answer artefacts were produced under ``set.seed()``:

* ``tests/testthat/test_BiclusterExperiment.R:5-7`` builds its 6x6 input with
  ``set.seed(1261534); matrix(runif(36), 6)``;
* ``als_nmf`` draws ``W <- matrix(runif(m*k), m, k)`` and
  ``Hold <- matrix(runif(k*n), k, n)`` per replicate (``R/bicluster.R:95,104``)
  and ``NMF::.fcnnls`` silently consumes uniforms through
  ``max.col(ties.method = "random")``;
* ``bcvGivenKs`` draws its folds with ``sample()`` (``R/bcv.R:170-174``).

Pinned: ``runif`` reproduces the fixture's 6x6 input bit for bit
(``tests/test_oracle_golden.py``).  ``sample()`` is restated from R's
``R_unif_index`` (rejection sampling) -- parity unpinned (no artefact records
a permutation).
"""
from __future__ import annotations

import numpy as np

_N, _M = 624, 397
_I2_32M1 = 2.328306437080797e-10  # 1/(2^32 - 1)


class RRng:
    """Mersenne-Twister stream positioned exactly like R's ``.Random.seed``."""

    def __init__(self, seed: int | None = None):
        self.mt = [0] * _N
        self.mti = _N + 1
        self.draws = 0
        if seed is not None:
            self.set_seed(seed)

    # RNG.c: RNG_Init + FixupSeeds for MERSENNE_TWISTER
    def set_seed(self, seed: int) -> None:
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        iseed = []
        for _ in range(_N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            iseed.append(s)
        # dummy[0] = mti = N  (FixupSeeds: i_seed[0] = 624)
        self.mt = iseed[1:]
        self.mti = _N
        self.draws = 0

    def _genrand(self) -> int:
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N - _M):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + _M] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            for kk in range(_N - _M, _N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + (_M - _N)] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            y = (mt[_N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
            mt[_N - 1] = mt[_M - 1] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    @classmethod
    def from_random_seed(cls, seed_vector) -> "RRng":
        """Position the stream from an R ``.Random.seed`` integer vector (kind code, mti, 624 state words)."""
        r = cls()
        r.mti = int(seed_vector[1])
        r.mt = [int(v) & 0xFFFFFFFF for v in seed_vector[2:2 + _N]]
        return r

    def get_state(self):
        return list(self.mt), self.mti, self.draws

    def set_state(self, state) -> None:
        self.mt, self.mti, self.draws = list(state[0]), state[1], state[2]

    def unif_rand(self) -> float:
        self.draws += 1
        v = self._genrand() * 2.3283064365386963e-10
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def runif(self, n: int) -> np.ndarray:
        return np.array([self.unif_rand() for _ in range(n)], dtype=np.float64)

    # R_unif_index (R >= 3.6.0, sample.kind = "Rejection")
    def _rbits(self, bits: int) -> int:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(np.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn: int) -> int:
        if dn <= 0:
            return 0
        bits = int(np.ceil(np.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample(self, x) -> np.ndarray:
        """``sample(x, size = length(x))``: permutation without replacement."""
        x = list(x)
        n = len(x)
        out = []
        for _ in range(n):
            j = self.unif_index(n)
            out.append(x[j])
            n -= 1
            x[j] = x[n]
        return np.array(out)

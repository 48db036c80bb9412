"""std::mt19937 + libstdc++ uniform_real_distribution<double>(0,1) restated in pure Python.

Test infrastructure (see oracle/__init__.py).

Reference use: ``MonteCarloSweepUpdaterBase`` keeps ``std::mt19937 random_engine_`` seeded by an
``unsigned int`` and draws with ``std::uniform_real_distribution<double> u_double_(0, 1.0)``
(vmc_basic/configuration_update_strategies/monte_carlo_sweep_updater_base.h:28-46).

libstdc++ ``generate_canonical<double, 53>`` on a 32-bit engine takes two draws x0, x1 and returns
``(double(x0) + double(x1) * 2^32) / 2^64`` (one rounding of the 64-bit integer to 53 bits), replaced by
``nextafter(1, 0)`` when the quotient rounds up to 1.0.
"""
import math

N, M = 624, 397
MATRIX_A, UPPER, LOWER = 0x9908B0DF, 0x80000000, 0x7FFFFFFF


class MT19937:
    """ISO C++ ``std::mt19937`` (32-bit Mersenne twister, ``init_genrand`` seeding)."""

    def __init__(self, seed=5489):
        self.mt = [0] * N
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, N):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = N

    def _twist(self):
        mt = self.mt
        for k in range(N):
            y = (mt[k] & UPPER) | (mt[(k + 1) % N] & LOWER)
            mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        self.idx = 0

    def next_u32(self):
        if self.idx >= N:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def uniform01(self):
        """``std::uniform_real_distribution<double>(0,1)(engine)`` for libstdc++."""
        x0 = self.next_u32()
        x1 = self.next_u32()
        r = float(x0 + (x1 << 32)) / 18446744073709551616.0
        if r >= 1.0:
            r = math.nextafter(1.0, 0.0)
        return r

    def uniform_long_double(self, a, b):
        """``std::uniform_real_distribution<long double>(a, b)(engine)`` for libstdc++ on x86 (80-bit long double,
        ``generate_canonical<long double, 64>`` = two 32-bit draws, exact in the 64-bit significand).
        ``a``, ``b`` are numpy longdouble."""
        import numpy as np
        ld = np.longdouble
        assert np.finfo(ld).nmant == 63, "oracle needs the x87 80-bit long double of the reference's platform"
        x0 = self.next_u32()
        x1 = self.next_u32()
        u = (ld(x0) + ld(x1) * ld(4294967296.0)) / (ld(4294967296.0) * ld(4294967296.0))
        if u >= ld(1.0):
            u = np.nextafter(ld(1.0), ld(0.0))
        return u * (b - a) + a

    def state(self):
        """(624 words, index) -- the layout the C ABI's set/get RNG state uses."""
        return list(self.mt), self.idx


def shuffle_std(lst, rng):
    """``std::shuffle(first, last, g)`` is implementation-defined; the oracle and the GPU harness
    both use this explicit Fisher-Yates with ``next_u32() % (i+1)`` so seeds are portable."""
    a = list(lst)
    for i in range(len(a) - 1, 0, -1):
        j = rng.next_u32() % (i + 1)
        a[i], a[j] = a[j], a[i]
    return a

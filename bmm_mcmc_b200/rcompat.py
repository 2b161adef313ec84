"""R-compatible Mersenne-Twister front end (host side).

The reference draws its *initial states* with R's global RNG before entering C++
(R/utils.R:42 `sample(1:K, N, replace=T)`, R/utils.R:68-74,98-103 `runif`), and its bundled
datasets were generated with `set.seed(17)` + `rbinom(n, 1, p)` (R/simulate_data.R:3-52).
This module restates the documented R algorithms (set.seed scrambling, MT19937 `unif_rand`,
n=1 `rbinom` inversion, R >= 3.6 "Rejection" `sample`) so the host API can reproduce them.

`set.seed` + `unif_rand` + `rbinom(.,1,p)` are pinned bit-exactly by the bundled data
(tests/test_fixtures.py); `sample_int` follows R 3.6's R_unif_index and is unpinned.
"""
import math
import numpy as np

_I2_32M1 = 2.328306437080797e-10


class RRng:
    """`set.seed(seed)`; then `unif_rand()` etc. (R default: Mersenne-Twister, Inversion)."""

    def __init__(self, seed):
        x = np.uint32(int(seed) & 0xFFFFFFFF)
        x = int(x)
        for _ in range(50):
            x = (69069 * x + 1) & 0xFFFFFFFF
        words = []
        for _ in range(625):
            x = (69069 * x + 1) & 0xFFFFFFFF
            words.append(x)
        # i_seed[0] (dummy[0] = mti) is overwritten with 624 by FixupSeeds
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = np.array(words[1:], dtype=np.uint32)
        st["state"]["pos"] = 624
        self._bg.state = st

    def unif_rand(self, n=None):
        raw = self._bg.random_raw(1 if n is None else n).astype(np.float64)
        u = raw * 2.3283064365386963e-10
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        return float(u[0]) if n is None else u

    def runif(self, n):
        return self.unif_rand(int(n))

    def rbinom1(self, pp):
        """rbinom(1, 1, pp): one uniform unless pp is 0 or 1."""
        if pp == 0.0:
            return 0
        if pp == 1.0:
            return 1
        p = min(pp, 1.0 - pp)
        q = 1.0 - p
        u = self.unif_rand()
        ix = 0 if u < q else 1
        return 1 - ix if pp > 0.5 else ix

    def rbinom_n1(self, n, pp):
        return np.array([self.rbinom1(pp) for _ in range(n)], dtype=np.int32)

    def _rbits(self, bits):
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return float(v)

    def unif_index(self, dn):
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return int(dv)

    def sample_int(self, k, n):
        """sample(1:k, n, replace=TRUE) under R >= 3.6 sample.kind="Rejection"."""
        return np.array([self.unif_index(k) + 1 for _ in range(n)], dtype=np.int32)


_DATASETS = {
    # name: (N, theta_actual rows, cluster ratios)  -- R/simulate_data.R:3-18,22-35,39-52
    "K3_N1000_P5": (1000, [[0.7, 0.8, 0.2, 0.1, 0.1],
                           [0.3, 0.5, 0.9, 0.8, 0.6],
                           [0.1, 0.2, 0.5, 0.4, 0.9]], [0.6, 0.2, 0.2]),
    "K2_N100_P5": (100, [[0.7, 0.8, 0.2, 0.1, 0.1],
                         [0.2, 0.2, 0.9, 0.8, 0.6]], [0.7, 0.3]),
    "K2_N1000_P5": (1000, [[0.7, 0.8, 0.2, 0.1, 0.1],
                           [0.2, 0.2, 0.9, 0.8, 0.6]], [0.7, 0.3]),
}


def load_dataset(name):
    """Regenerate a bundled dataset (data/*.RData) from `set.seed(17)` as R/simulate_data.R does.

    Returns an int32 (N, P) array.  Equality with the decoded .RData files is checked in
    tests/test_fixtures.py.
    """
    N, theta, ratios = _DATASETS[name]
    rng = RRng(17)
    blocks = []
    for row, ratio in zip(theta, ratios):
        # R's round() is half-even, like Python's
        n = int(round(N * ratio))
        cols = [rng.rbinom_n1(n, p) for p in row]
        blocks.append(np.stack(cols, axis=1))
    return np.ascontiguousarray(np.concatenate(blocks, axis=0), dtype=np.int32)


DATASET_NAMES = tuple(_DATASETS)

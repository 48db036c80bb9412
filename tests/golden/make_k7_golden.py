"""K7: all-pairs correlators of the reference's 4x4 D=8 Heisenberg fixture against exact diagonalisation
(tests/test_data/ed_reference/square_heisenberg_4x4_obc_ed.json). Computes the oracle amplitude of EVERY S_z = 0 configuration
(12870 of them, boundary-MPS contraction at the reference's truncation (8, 16, 1e-15)) and stores them with the ED numbers, so
that tests/test_oracle_kat.py can form <Sz Sz> and <S+ S- + S- S+>/2 by exact summation without the reference tree.
Run in the build container (needs /root/reference): python tests/golden/make_k7_golden.py   (a few minutes on 8 cores)."""
import itertools
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import load_golden_tps      # noqa: E402
from oracle import vmc                   # noqa: E402

ED = "/root/reference/tests/test_data/ed_reference/square_heisenberg_4x4_obc_ed.json"
_T = {}


def _amp(bits):
    if "tps" not in _T:
        try:
            from threadpoolctl import threadpool_limits
            _T["lim"] = threadpool_limits(limits=1)
        except Exception:
            pass
        _T["tps"], _ = load_golden_tps("heis4x4_D8_double")
    cfg = np.array([(bits >> k) & 1 for k in range(16)], dtype=np.int64).reshape(4, 4)
    return vmc.Walker(_T["tps"], cfg, (8, 16, 1e-15)).amplitude


if __name__ == "__main__":
    states = np.array([sum(1 << k for k in ups) for ups in itertools.combinations(range(16), 8)], dtype=np.int64)
    with Pool(os.cpu_count() or 1) as pool:
        amps = np.array(pool.map(_amp, [int(s) for s in states], chunksize=64))
    ed = json.load(open(ED))
    pairs = sorted((tuple(int(x) for x in k.split(",")) for k in ed["correlations"]))
    np.savez_compressed(os.path.join(HERE, "heis4x4_D8_all_amplitudes.npz"), states=states, amplitudes=amps,
                        ed_energy=ed["ground_state_energy"], pairs=np.array(pairs),
                        ed_szsz=np.array([ed["correlations"][f"{i},{j}"]["SzSz"] for i, j in pairs]),
                        ed_spsm=np.array([ed["correlations"][f"{i},{j}"]["SpSm_plus_SmSp_over_2"] for i, j in pairs]))
    print("stored", len(states), "amplitudes; norm", float(np.sum(amps ** 2)))

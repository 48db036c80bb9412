"""Readers for the reference's on-disk formats (TrivialRepQN ``.qlten`` tensors, TPS directories,
``configuration<rank>`` files).  Test infrastructure (see oracle/__init__.py).

Format (decoded in SURVEY.md section 8c / Appendix B):
  ASCII header ``rank``; per index ``nsct``, per sector ``dgnc hash`` (TrivialRepQN), then
  ``dir dim hash``; then ``nblocks`` and ``rank`` coordinates per block; then the raw little-endian
  payload, row-major, index order (L, D, R, U); then ``\\n``.
Directory layout ``tps_ten{row}_{col}_{phys}.qlten`` + ``tps_meta.txt`` = ``rows cols phy_dim [bc]``
(two_dim_tn/tps/split_index_tps.h:23-29, split_index_tps_impl.h:317-322).
Configuration text grid: vmc_basic/configuration.h:446-464.
"""
import os
import numpy as np


def load_qlten(path, complex_=False):
    b = open(path, "rb").read()
    pos = 0

    def tok():
        nonlocal pos
        e = b.index(b"\n", pos)
        v = int(b[pos:e])
        pos = e + 1
        return v

    rank = tok()
    dims = []
    for _ in range(rank):
        nsct = tok()
        for _ in range(nsct):
            tok()
            tok()
        tok()
        dims.append(tok())
        tok()
    nblk = tok()
    for _ in range(nblk * rank):
        tok()
    dt = np.complex128 if complex_ else np.float64
    n = int(np.prod(dims))
    if nblk == 0:
        return np.zeros(dims, dtype=dt)
    arr = np.frombuffer(b[pos:pos + n * np.dtype(dt).itemsize], dtype=dt).reshape(dims)
    return np.array(arr)


def load_tps_dir(path, rows=None, cols=None, phys=None, complex_=False):
    """Returns tps[r][c] = list over physical index of arrays (L, D, R, U)."""
    meta = os.path.join(path, "tps_meta.txt")
    if os.path.exists(meta):
        toks = open(meta).read().split()
        if len(toks) >= 3:
            rows, cols, phys = int(toks[0]), int(toks[1]), int(toks[2])
    if rows is None:
        raise ValueError("tps_meta.txt is empty; pass rows/cols/phys")
    tps = [[[load_qlten(os.path.join(path, f"tps_ten{r}_{c}_{s}.qlten"), complex_)
             for s in range(phys)] for c in range(cols)] for r in range(rows)]
    return tps


def load_configuration(path):
    rows = [list(map(int, ln.split())) for ln in open(path).read().strip().splitlines() if ln.strip()]
    return np.array(rows, dtype=np.int64)

"""On-disk formats either side of the path, so states and warmed-up configurations written
by the reference drop in and vice versa: dense TrivialRepQN tensors (the bosonic models) and multi-block fZ2 tensors
(fermionic states: read_qlten_fz2 / load_fermion_tps return the dense blocks plus the parity of every index value, which
is all the engine's fermion mode needs).

  * ``.qlten`` tensor stream      TensorToolkit QLTensor stream I/O as used by SplitIndexTPS::Dump / Load
                                  (two_dim_tn/tps/split_index_tps_impl.h:300-400); decoded in SURVEY.md section 8c:
                                  ASCII header ``rank``; per index ``nsct``, per sector ``dgnc hash``, then
                                  ``dir dim hash``; ``nblocks`` + block coordinates; raw little-endian payload,
                                  row-major, legs (L, D, R, U); trailing newline.
  * TPS directory                 ``tps_ten{row}_{col}_{phys}.qlten`` + ``tps_meta.txt`` = ``rows cols phy_dim``
                                  (two_dim_tn/tps/split_index_tps.h:23-29)
  * ``configuration<rank>``       text grid, one row per line (vmc_basic/configuration.h:446-464)

Hash fields: TensorToolkit stores a hash per sector and per index, derived from its internal hashing of
(qn, degeneracy, direction); TensorToolkit's source is not available here, so the function is not reproduced. Writers
copy the hash fields from a ``template`` file of the same index layout (e.g. the state the reference dumped before
the optimisation step) -- that is the only way files written here are loadable by the reference. WITHOUT a template
the hash fields are written as 0: such files round-trip through this module and the oracle reader (which ignore the
hashes) but may be rejected by TensorToolkit's loader. Leg directions follow the reference's convention for projected
site tensors: L and U legs IN (-1), D and R legs OUT (+1) (cf. tests/test_data fixtures).
"""
import os
import struct

import numpy as np

from .api import SplitIndexTPS, FermionSplitIndexTPS, Configuration

def _tokens(buf):
    pos = 0
    while True:
        e = buf.index(b"\n", pos)
        yield int(buf[pos:e]), e + 1
        pos = e + 1


def read_qlten(path, dtype=None):
    """Returns the dense array (legs in file order). ``dtype`` None detects float64 / complex128 from the payload
    length. Raises ValueError for multi-sector (symmetric) tensors and for truncated or mistyped payloads."""
    buf = open(path, "rb").read()
    it = _tokens(buf)
    rank, pos = next(it)
    dims = []
    for _ in range(rank):
        nsct, pos = next(it)
        if nsct != 1:
            raise ValueError(f"{path}: index with {nsct} sectors; only TrivialRepQN tensors are supported")
        next(it); next(it)                      # dgnc, sector hash
        next(it)                                # direction
        d, pos = next(it)
        dims.append(d)
        _, pos = next(it)                       # index hash
    nblk, pos = next(it)
    for _ in range(nblk * rank):
        _, pos = next(it)
    n = int(np.prod(dims)) if dims else 1
    if nblk == 0:
        return np.zeros(dims, dtype=dtype or np.float64)
    payload = len(buf) - pos
    if payload and buf[-1:] == b"\n" and payload % 8 == 1:
        payload -= 1                            # trailing newline of the stream
    if dtype is None:
        if payload == n * 8:
            dtype = np.float64
        elif payload == n * 16:
            dtype = np.complex128
        else:
            raise ValueError(f"{path}: payload of {payload} bytes is neither {n} float64 nor {n} complex128 elements")
    item = np.dtype(dtype).itemsize
    if payload != n * item:
        raise ValueError(f"{path}: payload of {payload} bytes, expected {n} x {item} (wrong element type or truncated file)")
    return np.frombuffer(buf[pos:pos + n * item], dtype=dtype).reshape(dims).copy()


def _fz2_header(path, buf):
    """Parses the header of an fZ2 stream: (legs = [(sectors [(qn, dgnc)], dir, dim)], block coordinates, payload offset)."""
    it = _tokens(buf)
    rank, pos = next(it)
    legs = []
    for _ in range(rank):
        nsct, pos = next(it)
        scts = []
        for _ in range(nsct):
            qn, _ = next(it); next(it)
            dg, _ = next(it); next(it)
            scts.append((qn, dg))
        di, _ = next(it)
        dm, _ = next(it)
        _, pos = next(it)
        if sum(dg for _, dg in scts) != dm:
            raise ValueError(f"{path}: sector degeneracies do not add up to the index dimension (not an fZ2 tensor?)")
        legs.append((scts, di, dm))
    nblk, pos = next(it)
    coords = []
    for _ in range(nblk):
        c = []
        for _ in range(rank):
            v, pos = next(it)
            c.append(v)
        coords.append(c)
    return legs, coords, pos


def read_qlten_fz2(path, dtype=None):
    """fZ2-graded (block-sparse) QLTensor stream. Per index: ``nsct``, per sector ``qnval qnhash dgnc hash``, then
    ``dir dim hash``; ``nblocks`` and the sector coordinates of every block; payload = the blocks in listed order, each
    row-major. ``dtype`` None detects float64 / complex128 from the payload length. Returns (dense array, [parity of every
    index value per leg], [direction per leg])."""
    buf = open(path, "rb").read()
    legs, coords, pos = _fz2_header(path, buf)
    rank = len(legs)
    if dtype is None:                           # the element count is known: the stream ends with up to two newlines
        nel = sum(int(np.prod([legs[k][0][c[k]][1] for k in range(rank)])) for c in coords)
        extra = len(buf) - pos - 16 * nel
        dtype = np.complex128 if nel and 0 <= extra <= 2 and buf[len(buf) - extra:] == b"\n" * extra else np.float64
    out = np.zeros([l[2] for l in legs], dtype=dtype)
    offs = [np.concatenate([[0], np.cumsum([dg for _, dg in l[0]])]) for l in legs]
    item = np.dtype(dtype).itemsize
    for c in coords:
        shp = [legs[k][0][c[k]][1] for k in range(rank)]
        n = int(np.prod(shp))
        if pos + n * item > len(buf):
            raise ValueError(f"{path}: truncated payload (wrong element type?)")
        out[tuple(slice(offs[k][c[k]], offs[k][c[k] + 1]) for k in range(rank))] = \
            np.frombuffer(buf[pos:pos + n * item], dtype=dtype).reshape(shp)
        pos += n * item
    if len(buf) - pos > 2 or buf[pos:].strip(b"\n"):
        raise ValueError(f"{path}: {len(buf) - pos} trailing bytes (wrong element type?)")
    par = [np.concatenate([np.full(dg, qn % 2, dtype=np.int32) for qn, dg in l[0]]) for l in legs]
    return out, par, [l[1] for l in legs]


def write_qlten_fz2(path, array, template):
    """Rewrites an fZ2 tensor with new elements: header (sectors, hashes, block list) copied verbatim from the ``template``
    file -- the state the reference dumped before the optimisation has the same index structure as the optimised one --
    followed by the blocks of ``array`` (dense, legs in file order) in the template's block order. Elements of ``array``
    outside the template's blocks must be zero (they would be lost)."""
    buf = open(template, "rb").read()
    legs, coords, pos = _fz2_header(template, buf)
    rank = len(legs)
    a = np.asarray(array)
    if list(a.shape) != [l[2] for l in legs]:
        raise ValueError(f"{path}: array shape {a.shape} does not match the template's index dimensions")
    dt = np.complex128 if np.iscomplexobj(a) else np.float64
    offs = [np.concatenate([[0], np.cumsum([dg for _, dg in l[0]])]) for l in legs]
    mask = np.zeros(a.shape, dtype=bool)
    parts = []
    for c in coords:
        sl = tuple(slice(offs[k][c[k]], offs[k][c[k] + 1]) for k in range(rank))
        parts.append(np.ascontiguousarray(a[sl], dtype=dt).tobytes())
        mask[sl] = True
    if np.any(a[~mask] != 0):
        raise ValueError(f"{path}: non-zero elements outside the template's blocks (parity-violating entries)")
    with open(path, "wb") as f:
        f.write(buf[:pos] + b"".join(parts) + b"\n")


def dump_fermion_tps(ftps, directory, template_dir):
    """SplitIndexTPS<T, fZ2QN>::Dump with the headers of ``template_dir`` (a TPS directory of the same index structure, e.g.
    the initial state): the files load back into the reference."""
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "tps_meta.txt"), "w") as f:
        f.write(f"{ftps.rows()} {ftps.cols()} {ftps.PhysicalDim()}")
    for r in range(ftps.rows()):
        for c in range(ftps.cols()):
            for s in range(ftps.PhysicalDim()):
                name = f"tps_ten{r}_{c}_{s}.qlten"
                write_qlten_fz2(os.path.join(directory, name), np.asarray(ftps((r, c))[s])[..., None], os.path.join(template_dir, name))


def load_fermion_tps(directory):
    """SplitIndexTPS<T, fZ2QN>::Load: tensors (L, D, R, U, parity leg of dim 1) -> FermionSplitIndexTPS."""
    rows, cols, phys = map(int, open(os.path.join(directory, "tps_meta.txt")).read().split()[:3])
    T = [[[None] * phys for _ in range(cols)] for _ in range(rows)]
    par = [[None] * cols for _ in range(rows)]
    phys_par = [None] * phys
    for r in range(rows):
        for c in range(cols):
            for s in range(phys):
                d, p, dirs = read_qlten_fz2(os.path.join(directory, f"tps_ten{r}_{c}_{s}.qlten"))
                if d.ndim != 5 or d.shape[4] != 1 or dirs != [-1, 1, 1, -1, -1]:
                    raise ValueError("expected (L in, D out, R out, U in, parity in) site tensors")
                T[r][c][s] = np.ascontiguousarray(d[..., 0])
                if par[r][c] is None:
                    par[r][c] = p[:4]
                elif any((a != b).any() for a, b in zip(par[r][c], p[:4])):
                    raise ValueError("virtual index sectors differ between the physical components of a site")
                if phys_par[s] is None:
                    phys_par[s] = int(p[4][0])
                elif phys_par[s] != int(p[4][0]):
                    raise ValueError("parity of a physical state differs between sites")
    return FermionSplitIndexTPS(T, par, phys_par)


def _header_template(path):
    """Parses the index descriptors (dir, hashes) of an existing file so writers can reproduce them."""
    buf = open(path, "rb").read()
    it = _tokens(buf)
    rank, _ = next(it)
    idx = []
    for _ in range(rank):
        next(it)
        dg, _ = next(it)
        sh, _ = next(it)
        di, _ = next(it)
        dm, _ = next(it)
        ih, _ = next(it)
        idx.append(dict(dgnc=dg, sector_hash=sh, dir=di, dim=dm, index_hash=ih))
    return idx


def write_qlten(path, array, dirs=(-1, 1, 1, -1), template=None):
    """Writes a dense real tensor as a one-block TrivialRepQN ``.qlten`` stream. ``template`` (index descriptors from
    _header_template of a file with the same dims) supplies TensorToolkit's hash fields; without it the hashes are
    written as 0, which this module and the oracle reader ignore."""
    a = np.ascontiguousarray(array, dtype=np.float64)
    out = [str(a.ndim)]
    for k, d in enumerate(a.shape):
        t = template[k] if template is not None else None
        if t is not None and t["dim"] != d:
            raise ValueError("template dims do not match")
        out += ["1", str(d), str(t["sector_hash"] if t else 0), str(t["dir"] if t else dirs[k % len(dirs)]), str(d),
                str(t["index_hash"] if t else 0)]
    out.append("1")
    out += ["0"] * a.ndim
    with open(path, "wb") as f:
        f.write(("\n".join(out) + "\n").encode())
        f.write(a.tobytes())
        f.write(b"\n")


def load_tps(directory, rows=None, cols=None, phys=None):
    """SplitIndexTPS::Load: returns a peps_b200.api.SplitIndexTPS."""
    meta = os.path.join(directory, "tps_meta.txt")
    if os.path.exists(meta):
        toks = open(meta).read().split()
        if len(toks) >= 3:
            rows, cols, phys = int(toks[0]), int(toks[1]), int(toks[2])
    if rows is None or cols is None or phys is None:
        raise ValueError("tps_meta.txt missing or empty: pass rows, cols, phys")
    return SplitIndexTPS([[[read_qlten(os.path.join(directory, f"tps_ten{r}_{c}_{s}.qlten")) for s in range(phys)]
                           for c in range(cols)] for r in range(rows)])


def dump_tps(tps: SplitIndexTPS, directory, template_dir=None):
    """SplitIndexTPS::Dump."""
    os.makedirs(directory, exist_ok=True)
    phys = tps.PhysicalDim()
    with open(os.path.join(directory, "tps_meta.txt"), "w") as f:
        f.write(f"{tps.rows()} {tps.cols()} {phys}")
    for r in range(tps.rows()):
        for c in range(tps.cols()):
            for s in range(phys):
                name = f"tps_ten{r}_{c}_{s}.qlten"
                tmpl = None
                if template_dir is not None and os.path.exists(os.path.join(template_dir, name)):
                    tmpl = _header_template(os.path.join(template_dir, name))
                write_qlten(os.path.join(directory, name), tps((r, c))[s], template=tmpl)


def load_configuration(path):
    rows = [list(map(int, ln.split())) for ln in open(path).read().strip().splitlines() if ln.strip()]
    return Configuration(np.array(rows, dtype=np.int32))


def dump_configuration(cfg: Configuration, path):
    with open(path, "w") as f:
        for row in cfg.data:
            f.write(" ".join(str(int(x)) for x in row) + "\n")

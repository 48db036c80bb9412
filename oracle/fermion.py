"""Fermionic (fZ2-graded) PEPS sampling path: CPU restatement.  Test infrastructure (see oracle/__init__.py).

Two independent restatements live here and are checked against each other and against the reference's goldens (K8):

1. ``graded_amplitude``: a brute-force Z2-graded tensor-network contraction (dense data + a parity per index value,
   Koszul sign on every transposition, supertrace sign on the contracted pair).  The graded algebra itself lives in
   QuantumLiquids/TensorToolkit (``qlten``; unpinned transitive dependency of the reference, absent from
   /root/reference), so its one free convention -- which of an (IN, OUT) pair is contracted without a sign -- is
   pinned by the reference's own known answers: exact summation over the 2x2 spinless-fermion fixtures
   (tests/test_data/spinless_fermion_tps_*) reproduces the analytic energies and the simple-update energies
   ``-4.1879072654, -1.98218053854, -4.98966397657`` (tests/test_algorithm/test_exact_summation_evaluator.cpp:137-151,
   353-425, 428-458) only with ``CONV = 1``; the t-J fixture reproduces ``-2.9431635706137875`` (:807).

2. the *dressed bosonic* formulation the device path uses.  A graded network of parity-conserving site tensors equals an
   ordinary (bosonic) network of sign-dressed site tensors: B[l,d,r,u] = T[l,d,r,u] * (-1)^q with
       horizontal machinery (UP/DOWN boundary MPS, LEFT/RIGHT BTen, BTen2):
           q_H = l d + l r + d r + l + d + u J_H,   J_H(r,c) = parity of the sites left of (r,c) in its row
       vertical machinery (LEFT/RIGHT boundary MPS, UP/DOWN BTen):
           q_V = l d + l r + l u + l + d + l J_V,   J_V(r,c) = parity of the sites above (r,c) in its column
   (l, d, r, u = parities of the index values).  psi_H is the amplitude in row-major mode order, psi_V in column-major
   order.  Every contraction the reference performs with graded tensors (bmps_impl.h:756-862 fermionic branches,
   bmps_contractor_trace.h:90-205, grow.h:150-183) is then the bosonic restatement of oracle/contractor.py on the
   dressed tensors: the boundary MPS differ from the reference's by diagonal sign gauges only, so singular values,
   truncation and every |amplitude| agree, and amplitude *ratios* along one contraction path carry the physical sign.
   The rules for the replaced tensors of a move (hop_variants) were derived from the graded contraction and are checked
   against it by tests/test_fermion_oracle.py.

Reference (relative to include/qlpeps/):
  * SquareSpinlessFermion              algorithm/vmc_update/model_solvers/square_spinless_fermion.h:51-213
  * SquaretJNNModel / MixIn            algorithm/vmc_update/model_solvers/square_tJ_model.h:85-420
  * fermionic traversal (psi per bond) algorithm/vmc_update/model_solvers/base/square_nnn_energy_solver.h:143-310
  * CalGTenForFermionicTensors         utility/helpers.h:57-67  (+ ActFermionPOps: O* is the Euclidean gradient of
                                       log psi*, docs/dev/design/math/fermion-vmc-math.md)
  * MCEnergyGradEvaluator (fermion)    algorithm/vmc_update/mc_energy_grad_evaluator.h:245-272
"""
import os
import numpy as np
from .bmps import LEFT, DOWN, RIGHT, UP, HORIZONTAL, VERTICAL
from .contractor import BMPSContractor
from .mt19937 import MT19937

CONV = 1            # pinned by K8 (see module docstring)
ML, MD, MR, MU = 1, 2, 4, 8      # leg-sign mask bits


# ---------------------------------------------------------------------------------------------------------------
# fZ2 .qlten reader (multi-block): dense array + parity of every index value + index directions
# ---------------------------------------------------------------------------------------------------------------
def load_qlten_fz2(path, complex_=False):
    """Header: rank; per index: nsct, per sector ``qnval qnhash dgnc hash``, then ``dir dim hash``; nblocks and the
    block coordinates; payload = the blocks in listed order, each row-major."""
    b = open(path, "rb").read()
    pos = 0

    def tok():
        nonlocal pos
        e = b.index(b"\n", pos)
        v = int(b[pos:e])
        pos = e + 1
        return v

    rank = tok()
    legs = []
    for _ in range(rank):
        nsct = tok()
        scts = []
        for _ in range(nsct):
            qn = tok(); tok(); dg = tok(); tok()
            scts.append((qn, dg))
        d = tok(); dim = tok(); tok()
        legs.append((scts, d, dim))
    nblk = tok()
    coords = [[tok() for _ in range(rank)] for _ in range(nblk)]
    dt = np.complex128 if complex_ else np.float64
    out = np.zeros([l[2] for l in legs], dtype=dt)
    offs = []
    for scts, d, dim in legs:
        o = [0]
        for qn, dg in scts:
            o.append(o[-1] + dg)
        if o[-1] != dim:
            raise ValueError("sector degeneracies do not add up to the index dimension")
        offs.append(o)
    isz = np.dtype(dt).itemsize
    for c in coords:
        shp = [legs[k][0][c[k]][1] for k in range(rank)]
        n = int(np.prod(shp))
        if pos + n * isz > len(b):
            raise ValueError(f"{path}: truncated payload")
        blk = np.frombuffer(b[pos:pos + n * isz], dtype=dt).reshape(shp)
        pos += n * isz
        out[tuple(slice(offs[k][c[k]], offs[k][c[k] + 1]) for k in range(rank))] = blk
    if len(b) - pos > 1:
        raise ValueError(f"{path}: {len(b) - pos} trailing bytes (wrong element type?)")
    par = [np.concatenate([np.full(dg, qn % 2, dtype=np.int64) for qn, dg in scts]) for scts, d, dim in legs]
    return out, par, [l[1] for l in legs]


class FermionTPS:
    """Split-index fermionic TPS as dense tensors + parities.
    T[r][c][s] has shape (L, D, R, U); par[r][c] = [pL, pD, pR, pU] parity vectors; phys_par[s] = parity of state s."""

    def __init__(self, T, par, phys_par):
        self.T, self.par, self.phys_par = T, par, tuple(int(x) for x in phys_par)
        self.rows, self.cols, self.phys = len(T), len(T[0]), len(T[0][0])

    @staticmethod
    def load(path, complex_=False):
        rows, cols, phys = map(int, open(os.path.join(path, "tps_meta.txt")).read().split()[:3])
        T = [[[None] * phys for _ in range(cols)] for _ in range(rows)]
        par = [[None] * cols for _ in range(rows)]
        phys_par = [None] * phys
        for r in range(rows):
            for c in range(cols):
                for s in range(phys):
                    d, p, dirs = load_qlten_fz2(os.path.join(path, f"tps_ten{r}_{c}_{s}.qlten"), complex_)
                    if dirs != [-1, 1, 1, -1, -1]:
                        raise ValueError("unexpected index directions")
                    T[r][c][s] = d[..., 0]
                    if par[r][c] is None:
                        par[r][c] = p[:4]
                    elif any((a != b_).any() for a, b_ in zip(par[r][c], p[:4])):
                        raise ValueError("virtual index sectors differ between physical components")
                    pp = int(p[4][0])
                    if phys_par[s] is None:
                        phys_par[s] = pp
                    elif phys_par[s] != pp:
                        raise ValueError("parity of a physical state differs between sites")
        return FermionTPS(T, par, phys_par)

    @staticmethod
    def random(rows, cols, D, seed, phys_par=(1, 0), complex_=False, n_odd=None):
        """Z2-symmetric random state: every bond has D//2 (or n_odd) odd index values placed last."""
        rng = np.random.default_rng(seed)
        n_odd = D // 2 if n_odd is None else n_odd
        bp = np.array([0] * (D - n_odd) + [1] * n_odd, dtype=np.int64)
        one = np.array([0], dtype=np.int64)
        T = [[None] * cols for _ in range(rows)]
        par = [[None] * cols for _ in range(rows)]
        for r in range(rows):
            for c in range(cols):
                ps = [bp if c > 0 else one, bp if r < rows - 1 else one, bp if c < cols - 1 else one, bp if r > 0 else one]
                tot = (ps[0][:, None, None, None] + ps[1][None, :, None, None] + ps[2][None, None, :, None]
                       + ps[3][None, None, None, :]) % 2
                par[r][c] = ps
                T[r][c] = []
                for pp in phys_par:
                    a = rng.uniform(-1.0, 1.0, tot.shape)
                    if complex_:
                        a = a + 1j * rng.uniform(-1.0, 1.0, tot.shape)
                    T[r][c].append(a * (tot == pp))
        return FermionTPS(T, par, phys_par)

    # -- dressing -------------------------------------------------------------------------------------------------
    def local_sign(self, r, c, mode):
        pl, pd, pr, pu = self.par[r][c]
        l = pl[:, None, None, None]; d = pd[None, :, None, None]; rr = pr[None, None, :, None]; u = pu[None, None, None, :]
        q = l * d + l * rr + d * rr + l + d if mode == HORIZONTAL else l * d + l * rr + l * u + l + d
        return 1 - 2 * (q % 2)

    def mask_sign(self, r, c, mask):
        s = 1
        shp = [(-1, 1, 1, 1), (1, -1, 1, 1), (1, 1, -1, 1), (1, 1, 1, -1)]
        for k in range(4):
            if mask >> k & 1:
                s = s * (1 - 2 * self.par[r][c][k].reshape(shp[k]))
        return s

    def sign(self, r, c, mode, mask):
        return self.local_sign(r, c, mode) * self.mask_sign(r, c, mask)

    def variant(self, r, c, s, mode, mask):
        return self.T[r][c][s] * self.sign(r, c, mode, mask)

    def parities(self, config):
        return np.array(self.phys_par, dtype=np.int64)[np.asarray(config)]

    def jw(self, config, mode):
        """J_H (parity of the sites to the left in the row) or J_V (above in the column)."""
        p = self.parities(config)
        ex = np.cumsum(p, axis=1 if mode == HORIZONTAL else 0) - p
        return ex % 2

    def canonical_masks(self, config, mode):
        return self.jw(config, mode) * (MU if mode == HORIZONTAL else ML)

    def tn(self, config, mode):
        m = self.canonical_masks(config, mode)
        return [[self.variant(r, c, int(config[r][c]), mode, int(m[r, c])) for c in range(self.cols)]
                for r in range(self.rows)]


# ---------------------------------------------------------------------------------------------------------------
# brute-force graded contraction (pins the convention; small lattices only)
# ---------------------------------------------------------------------------------------------------------------
class _GT:
    def __init__(self, data, par, dirs, lab):
        self.d, self.par, self.dirs, self.lab = data, list(par), list(dirs), list(lab)

    def transpose(self, perm):
        perm = list(perm)
        n = len(perm)
        sign = 1
        for x in range(n):
            for y in range(x + 1, n):
                if perm[x] > perm[y]:                       # the two legs change their relative order
                    pa = self.par[perm[y]].reshape([-1 if k == perm[y] else 1 for k in range(n)])
                    pb = self.par[perm[x]].reshape([-1 if k == perm[x] else 1 for k in range(n)])
                    sign = sign * (1 - 2 * (pa * pb))
        return _GT((self.d * sign).transpose(perm), [self.par[p] for p in perm], [self.dirs[p] for p in perm],
                   [self.lab[p] for p in perm])


def _gcontract(A, B, shared):
    ia = [A.lab.index(x) for x in shared]
    ib = [B.lab.index(x) for x in shared]
    ra = [k for k in range(len(A.lab)) if k not in ia]
    rb = [k for k in range(len(B.lab)) if k not in ib]
    A2 = A.transpose(ra + ia)
    B2 = B.transpose(ib[::-1] + rb)                         # nested pairing: (.., a1, a2)(b2, b1, ..)
    da = A2.d
    nA, k = len(ra), len(ia)
    for j, ax in enumerate(range(nA, nA + k)):
        if A2.dirs[ax] != -B2.dirs[k - 1 - j]:
            raise ValueError("index directions do not match")
        if (A2.dirs[ax] == 1) if CONV == 0 else (A2.dirs[ax] == -1):
            da = da * (1 - 2 * A2.par[ax].reshape([-1 if i == ax else 1 for i in range(da.ndim)]))
    res = np.tensordot(da, B2.d, axes=(list(range(nA, nA + k)), list(range(k - 1, -1, -1))))
    return _GT(res, A2.par[:nA] + B2.par[k:], A2.dirs[:nA] + B2.dirs[k:], A2.lab[:nA] + B2.lab[k:])


def graded_amplitude(ftps, config):
    """psi(S) with the parity legs in row-major site order: the coefficient of prod_{row-major} c^dag |0>."""
    config = np.asarray(config)
    acc = None
    for r in range(ftps.rows):
        for c in range(ftps.cols):
            s = int(config[r, c])
            g = _GT(ftps.T[r][c][s][..., None], ftps.par[r][c] + [np.array([ftps.phys_par[s]])], [-1, 1, 1, -1, -1],
                    [f"h{r}_{c - 1}", f"v{r}_{c}", f"h{r}_{c}", f"v{r - 1}_{c}", f"p{r}_{c}"])
            if acc is None:
                acc = g
            else:
                acc = _gcontract(acc, g, [x for x in acc.lab if x in g.lab])
    pl = [f"p{r}_{c}" for r in range(ftps.rows) for c in range(ftps.cols)]
    rest = [k for k, x in enumerate(acc.lab) if x not in pl]
    acc = acc.transpose(rest + [acc.lab.index(x) for x in pl])
    return acc.d.reshape(-1)[0]


def colmajor_sign(ftps, config):
    """psi_V = colmajor_sign * psi_H: sign of the permutation between row-major and column-major mode order."""
    p = ftps.parities(config)
    s = 0
    for r in range(ftps.rows):
        for c in range(ftps.cols):
            if p[r, c]:
                s += int(p[r + 1:, :c].sum())
    return 1 - 2 * (s % 2)


def between_sign(ftps, config, a, b):
    """Matrix-element sign of a hop between sites a and b in row-major mode order."""
    p = ftps.parities(config).reshape(-1)
    i, j = a[0] * ftps.cols + a[1], b[0] * ftps.cols + b[1]
    lo, hi = min(i, j), max(i, j)
    return 1 - 2 * (int(p[lo + 1:hi].sum()) % 2)


# ---------------------------------------------------------------------------------------------------------------
# walker, updater, models on the dressed formulation
# ---------------------------------------------------------------------------------------------------------------
def hop_variants(ftps, config, jw_h, jw_v, kind, a, b):
    """Replacement tensors and the sign for psi(S') / psi(S), S' = S with the states of a and b exchanged, evaluated
    inside the environments of S.  kind: 'h' (horizontal NN, horizontal machinery), 'v' (vertical NN, vertical
    machinery), 'dr' (a = left-up, b = right-down) and 'ur' (a = left-down, b = right-up; horizontal machinery, BTen2).
    Returns (ten_a, ten_b, sign)."""
    ca, cb = int(config[a]), int(config[b])
    delta = ftps.phys_par[ca] ^ ftps.phys_par[cb]          # does a fermion move?
    if kind == 'v':
        ma, mb = ML * int(jw_v[a]), ML * (int(jw_v[b]) ^ delta)
        return ftps.variant(*a, cb, VERTICAL, ma), ftps.variant(*b, ca, VERTICAL, mb), 1
    ma, mb = MU * int(jw_h[a]), MU * int(jw_h[b])
    sign = 1
    if delta:
        if kind == 'h':
            mb ^= MU
        elif kind == 'dr':
            ma ^= MR; mb ^= MU
            sign = 1 - 2 * int(jw_h[b])
        elif kind == 'ur':
            mb ^= MR | MD
            sign = 1 - 2 * int(jw_h[a])
    return ftps.variant(*a, cb, HORIZONTAL, ma), ftps.variant(*b, ca, HORIZONTAL, mb), sign


def bond_variants(ftps, config, jw_h, jw_v, kind, a, b, na, nb):
    """Replacement tensors of an NN bond (kind 'h' / 'v') for ANY parity-consistent pair of new states (na, nb): a hop, a
    pair creation / annihilation (both site parities change) or a parity-preserving change. The same masks as
    hop_variants; psi(S') / psi(S) along one contraction path equals between_sign * the ratio of the graded amplitudes
    (checked in tests/test_fermion_oracle.py)."""
    ca, cb = int(config[a]), int(config[b])
    da, db = ftps.phys_par[ca] ^ ftps.phys_par[na], ftps.phys_par[cb] ^ ftps.phys_par[nb]
    assert da == db, "one site parity alone cannot change"
    if kind == 'v':
        return ftps.variant(*a, na, VERTICAL, ML * int(jw_v[a])), ftps.variant(*b, nb, VERTICAL, ML * (int(jw_v[b]) ^ da))
    return ftps.variant(*a, na, HORIZONTAL, MU * int(jw_h[a])), ftps.variant(*b, nb, HORIZONTAL, MU * (int(jw_h[b]) ^ da))


def jastrow_ratio(v, density, config, a, b):
    """Jastrow-factor ratio (new / old) of exchanging the states of a and b: JastrowFieldAtSite
    (vmc_basic/jastrow_factor.h:99-111) and the exp(field difference) of square_nn_updater.h:402-419."""
    n = np.asarray(density)[np.asarray(config)]
    if n[a] == n[b]:
        return 1.0
    cols = n.shape[1]
    ia, ib = a[0] * cols + a[1], b[0] * cols + b[1]
    fa = fb = 0.0
    for j, nj in enumerate(n.reshape(-1)):
        if j != ia:
            fa += v[ia, j] * float(nj)
        if j != ib:
            fb += v[ib, j] * float(nj)
    return float(np.exp(fa - fb)) if n[a] < n[b] else float(np.exp(fb - fa))


class FermionWalker:
    """TPSWaveFunctionComponent for fZ2 tensors (wave_function_component.h:136-379): two dressed projections of the
    same configuration, tn_h for the row machinery and tn_v for the column machinery."""

    def __init__(self, ftps, config, trunc):
        self.ftps = ftps
        self.config = np.array(config, dtype=np.int64)
        self.rows, self.cols = self.config.shape
        self.trunc = trunc
        self.contractor = BMPSContractor(self.rows, self.cols)
        self._project()
        self.contractor.init(self.tn_h)
        self.amplitude = 0.0
        self.evaluate_amplitude()

    def _project(self):
        self.jw_h = self.ftps.jw(self.config, HORIZONTAL)
        self.jw_v = self.ftps.jw(self.config, VERTICAL)
        self.tn_h = self.ftps.tn(self.config, HORIZONTAL)
        self.tn_v = self.ftps.tn(self.config, VERTICAL)

    def evaluate_amplitude(self):
        c = self.contractor
        c.set_truncate_params(*self.trunc)
        c.grow_bmps_for_row(self.tn_h, 0)
        c.grow_full_bten(self.tn_h, RIGHT, 0, 2, True)
        c.init_bten(self.tn_h, LEFT, 0)
        self.amplitude = c.trace(self.tn_h, (0, 0), HORIZONTAL)
        return self.amplitude

    def update_local(self, new_amplitude, *site_configs):
        for (site, cfg) in site_configs:
            self.config[site[0], site[1]] = cfg
            self.contractor.erase_envs_after_update(site)
        # the dressings of the touched rows (tn_h) and columns (tn_v) follow the configuration; every cached
        # environment that contains one of them has just been erased
        self._project()
        self.amplitude = new_amplitude


class FermionNNExchangeUpdater:
    """MCUpdateSquareNNExchangeOBC on fZ2 tensors (square_nn_updater.h:29-81, 146-188): same decisions, same draws."""

    def __init__(self, seed, jastrow=None):
        """jastrow = (v[nsites][nsites], density[phys]): MCUpdateSquareNNExchangeJastrowDressedTJ
        (square_nn_updater.h:380-438)."""
        self.rng = MT19937(seed)
        self.jastrow = jastrow

    def two_site_update(self, a, b, bond_dir, w):
        c1, c2 = int(w.config[a]), int(w.config[b])
        if c1 == c2:
            return False
        kind = 'h' if bond_dir == HORIZONTAL else 'v'
        ta, tb, _ = hop_variants(w.ftps, w.config, w.jw_h, w.jw_v, kind, a, b)
        tn = w.tn_h if bond_dir == HORIZONTAL else w.tn_v
        psi_b = w.contractor.replace_nn_site_trace(tn, a, b, bond_dir, ta, tb)
        psi_a = w.amplitude
        if self.jastrow is not None:
            ratio = abs(psi_b * jastrow_ratio(self.jastrow[0], self.jastrow[1], w.config, a, b)) / abs(psi_a)
            if not (ratio >= 1.0 or self.rng.uniform01() < ratio * ratio):
                return False
        elif not abs(psi_b) >= abs(psi_a):
            div = abs(psi_b) / abs(psi_a)
            if not (self.rng.uniform01() < div * div):
                return False
        w.update_local(psi_b, (a, c2), (b, c1))
        return True

    def sweep(self, w):
        c = w.contractor
        rows, cols = w.rows, w.cols
        accepted = 0
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(w.tn_h, UP)
        for row in range(rows):
            c.init_bten(w.tn_h, LEFT, row)
            c.grow_full_bten(w.tn_h, RIGHT, row, 2, True)
            for col in range(cols - 1):
                accepted += self.two_site_update((row, col), (row, col + 1), HORIZONTAL, w)
                if col < cols - 2:
                    c.shift_bten_window(w.tn_h, RIGHT)
            if row < rows - 1:
                c.shift_bmps_window(w.tn_h, DOWN)
        c.delete_inner_bmps(LEFT)
        c.delete_inner_bmps(RIGHT)
        c.generate_bmps_approach(w.tn_v, LEFT)
        for col in range(cols):
            c.init_bten(w.tn_v, UP, col)
            c.grow_full_bten(w.tn_v, DOWN, col, 2, True)
            for row in range(rows - 1):
                accepted += self.two_site_update((row, col), (row + 1, col), VERTICAL, w)
                if row < rows - 2:
                    c.shift_bten_window(w.tn_v, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(w.tn_v, RIGHT)
        c.delete_inner_bmps(UP)
        return [accepted / (cols * (rows - 1) + rows * (cols - 1))]


def line_variants(ftps, jw_h, jw_v, orient, sites, states):
    """Dressed tensors of consecutive sites along a row (orient HORIZONTAL, row machinery) or a column (VERTICAL, column
    machinery) carrying NEW states: the Jordan-Wigner bit of each site follows the new states of the sites before it in
    the line (the bit of the first site is that of the current configuration); sites after the line are unaffected as
    long as the total parity of the line is conserved."""
    bit = int((jw_h if orient == HORIZONTAL else jw_v)[sites[0]])
    out = []
    for site, st in zip(sites, states):
        out.append(ftps.variant(site[0], site[1], st, orient, (MU if orient == HORIZONTAL else ML) * bit))
        bit ^= int(ftps.phys_par[st])
    return out


class FermionNNFullSpaceUpdater(FermionNNExchangeUpdater):
    """MCUpdateSquareNNFullSpaceUpdateOBC on fZ2 tensors (square_nn_updater.h:253-293; "work for both fermion and boson",
    :251): all d^2 local states of the bond, Suwa-Todo choice. Targets that change the parity of one site alone have zero
    amplitude (the parity-conserving tensors contract to zero), hence zero weight."""

    def two_site_update(self, a, b, bond_dir, w):
        from .vmc import suwa_todo_state_update, _std_norm
        d = w.ftps.phys
        par = w.ftps.phys_par
        c1, c2 = int(w.config[a]), int(w.config[b])
        init = c1 * d + c2
        tn = w.tn_h if bond_dir == HORIZONTAL else w.tn_v
        alt = [0.0] * (d * d)
        alt[init] = w.amplitude
        for q in range(d * d):
            na, nb = q // d, q % d
            if q == init or (par[c1] ^ par[na]) != (par[c2] ^ par[nb]):
                continue
            ta, tb = line_variants(w.ftps, w.jw_h, w.jw_v, bond_dir, [a, b], [na, nb])
            alt[q] = w.contractor.replace_nn_site_trace(tn, a, b, bond_dir, ta, tb)
        weights = [_std_norm(x / w.amplitude) for x in alt]
        final = suwa_todo_state_update(init, weights, self.rng)
        if final == init:
            return False
        w.update_local(alt[final], (a, final // d), (b, final % d))
        return True


class FermionTNN3SiteExchangeUpdater:
    """MCUpdateSquareTNN3SiteExchange on fZ2 tensors (square_3site_updater.h:23-160): permutations of the states of three
    consecutive sites (how holes move two sites in the t-J runs), Suwa-Todo choice; the cached amplitude is refreshed by a
    three-site trace at the start of every row / column."""

    def __init__(self, seed):
        self.rng = MT19937(seed)

    def three_site_update(self, sites, bond_dir, w):
        import itertools
        from .vmc import suwa_todo_state_update, _std_norm
        spins = [int(w.config[s]) for s in sites]
        if spins[0] == spins[1] == spins[2]:
            return False
        perms = sorted(set(itertools.permutations(sorted(spins))))
        init = perms.index(tuple(spins))
        tn = w.tn_h if bond_dir == HORIZONTAL else w.tn_v
        psis = []
        for i, pm in enumerate(perms):
            if i == init:
                psis.append(w.amplitude)
            else:
                t = line_variants(w.ftps, w.jw_h, w.jw_v, bond_dir, sites, pm)
                psis.append(w.contractor.replace_tnn_site_trace(tn, sites[0], bond_dir, t[0], t[1], t[2]))
        mx = max(abs(x) for x in psis)
        weights = [_std_norm(complex(x.real / mx, x.imag / mx) if np.iscomplexobj(x) else x / mx) for x in psis]
        final = suwa_todo_state_update(init, weights, self.rng)
        if final == init:
            return False
        pm = perms[final]
        w.update_local(psis[final], *[(s, st) for s, st in zip(sites, pm)])
        return True

    def sweep(self, w):
        c = w.contractor
        rows, cols = w.rows, w.cols
        accepted = 0
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(w.tn_h, UP)
        for row in range(rows):
            c.init_bten(w.tn_h, LEFT, row)
            c.grow_full_bten(w.tn_h, RIGHT, row, 3, True)
            w.amplitude = c.replace_tnn_site_trace(w.tn_h, (row, 0), HORIZONTAL, w.tn_h[row][0], w.tn_h[row][1], w.tn_h[row][2])
            for col in range(cols - 2):
                accepted += self.three_site_update([(row, col), (row, col + 1), (row, col + 2)], HORIZONTAL, w)
                if col < cols - 3:
                    c.shift_bten_window(w.tn_h, RIGHT)
            if row < rows - 1:
                c.shift_bmps_window(w.tn_h, DOWN)
        c.delete_inner_bmps(LEFT)
        c.delete_inner_bmps(RIGHT)
        c.generate_bmps_approach(w.tn_v, LEFT)
        for col in range(cols):
            c.init_bten(w.tn_v, UP, col)
            c.grow_full_bten(w.tn_v, DOWN, col, 3, True)
            w.amplitude = c.replace_tnn_site_trace(w.tn_v, (0, col), VERTICAL, w.tn_v[0][col], w.tn_v[1][col], w.tn_v[2][col])
            for row in range(rows - 2):
                accepted += self.three_site_update([(row, col), (row + 1, col), (row + 2, col)], VERTICAL, w)
                if row < rows - 3:
                    c.shift_bten_window(w.tn_v, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(w.tn_v, RIGHT)
        c.delete_inner_bmps(UP)
        return [accepted / (cols * (rows - 2) + rows * (cols - 2))]


class FermionModel:
    """SquareNNNModelEnergySolver traversal for fermionic tensors (square_nnn_energy_solver.h:104-310): psi is
    recomputed per bond by Trace (NN) / ReplaceNNNSiteTrace with the original tensors (NNN, once per plaquette)."""
    has_nnn = False
    jastrow = None          # (v, density): Jastrow-dressed solver (square_tJ_model.h:352-410)

    def _jr(self, w, a, b):
        return 1.0 if self.jastrow is None else jastrow_ratio(self.jastrow[0], self.jastrow[1], w.config, a, b)

    def diag_nn(self, c1, c2):
        return 0.0

    def offdiag_nn(self, c1, c2):
        """coefficient of conj(psi_ex / psi) for exchanging the two states (0 = no term)."""
        return 0.0

    def offdiag_nnn(self, c1, c2):
        return 0.0

    def onsite_energy(self, config):
        return 0.0

    pin = None              # (a, b, H): extra two-site table term on one NN bond (singlet-pair pinning, square_tJ_model.h:256-289)

    def table_value(self, H, a, b, orient, w, psi=None):
        """sum_p' H[p, p'] conj(psi(p') / psi) on the NN bond (a, b): a two-site operator given by its matrix in the basis
        p = c1 * d + c2 (EvaluateBondSC-style hooks, square_tJ_model.h:546-602)."""
        d = w.ftps.phys
        p = int(w.config[a]) * d + int(w.config[b])
        tn = w.tn_h if orient == HORIZONTAL else w.tn_v
        val = H[p, p]
        for q in range(d * d):
            if q == p or H[p, q] == 0.0:
                continue
            if psi is None:
                psi = w.contractor.trace(tn, a, orient)
            ta, tb = bond_variants(w.ftps, w.config, w.jw_h, w.jw_v, 'h' if orient == HORIZONTAL else 'v', a, b, q // d, q % d)
            val = val + H[p, q] * np.conj(w.contractor.replace_nn_site_trace(tn, a, b, orient, ta, tb) / psi)
        return val, psi

    def bond_energy(self, a, b, orient, w):
        e, psi = self._bond_energy(a, b, orient, w)
        if self.pin is not None and (a, b) == (tuple(self.pin[0]), tuple(self.pin[1])):
            pe, psi = self.table_value(self.pin[2], a, b, orient, w, psi)
            e = e + pe
        return e, psi

    def measure_bond_table(self, H, w):
        """(horizontal [rows][cols-1], vertical [rows-1][cols]) values of table_value on every NN bond."""
        saved = self._bond_energy, self.pin
        rec = {}
        self.pin = None
        self._bond_energy = lambda a, b, orient, ww: (rec.__setitem__((a, b), self.table_value(H, a, b, orient, ww)[0]) or 0.0,
                                                     ww.contractor.trace(ww.tn_h if orient == HORIZONTAL else ww.tn_v, a, orient))
        try:
            FermionModel.energy_and_holes(self, w, False)
        finally:
            del self._bond_energy
            self.pin = saved[1]
        dt = np.result_type(w.ftps.T[0][0][0].dtype, np.float64)
        h, v = np.zeros((w.rows, w.cols - 1), dt), np.zeros((w.rows - 1, w.cols), dt)
        for (a, b), val in rec.items():
            if a[0] == b[0]:
                h[a] = val
            else:
                v[a] = val
        return h, v

    def _bond_energy(self, a, b, orient, w):
        c1, c2 = int(w.config[a]), int(w.config[b])
        e = self.diag_nn(c1, c2)
        if c1 == c2:
            return e, None
        tn = w.tn_h if orient == HORIZONTAL else w.tn_v
        psi = w.contractor.trace(tn, a, orient)
        ta, tb, sg = hop_variants(w.ftps, w.config, w.jw_h, w.jw_v, 'h' if orient == HORIZONTAL else 'v', a, b)
        psi_ex = sg * w.contractor.replace_nn_site_trace(tn, a, b, orient, ta, tb)
        return e + self.offdiag_nn(c1, c2) * self._jr(w, a, b) * np.conj(psi_ex / psi), psi

    def nnn_energy(self, a, b, diagonal_dir, w, psi):
        c1, c2 = int(w.config[a]), int(w.config[b])
        if c1 == c2 or self.offdiag_nnn(c1, c2) == 0.0:
            return 0.0, psi
        left_up = a if diagonal_dir == 0 else (b[0], a[1])
        if psi is None:
            psi = w.contractor.replace_nnn_site_trace(w.tn_h, left_up, diagonal_dir, HORIZONTAL,
                                                      w.tn_h[a[0]][a[1]], w.tn_h[b[0]][b[1]])
        ta, tb, sg = hop_variants(w.ftps, w.config, w.jw_h, w.jw_v, 'dr' if diagonal_dir == 0 else 'ur', a, b)
        psi_ex = sg * w.contractor.replace_nnn_site_trace(w.tn_h, left_up, diagonal_dir, HORIZONTAL, ta, tb)
        return self.offdiag_nnn(c1, c2) * self._jr(w, a, b) * np.conj(psi_ex / psi), psi

    def energy_and_holes(self, w, calc_holes=True):
        """Returns (E_loc, O*[rows][cols] or None, psi_list).  O*(site) = conj(d psi / d T_site[cfg]) / conj(psi_site)
        with psi_site = <hole, T_site> rebuilt locally (utility/helpers.h:57-67)."""
        c = w.contractor
        rows, cols = w.rows, w.cols
        bond_e = []
        psi_list = []
        ostar = [[None] * cols for _ in range(rows)] if calc_holes else None
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(w.tn_h, UP)
        for row in range(rows):
            c.init_bten(w.tn_h, LEFT, row)
            c.grow_full_bten(w.tn_h, RIGHT, row, 1, True)
            psi_added = False
            for col in range(cols):
                if calc_holes:
                    hole = c.punch_hole(w.tn_h, (row, col), HORIZONTAL)
                    psi_site = np.sum(hole * w.tn_h[row][col])
                    sg = w.ftps.sign(row, col, HORIZONTAL, MU * int(w.jw_h[row, col]))
                    ostar[row][col] = np.conj(hole * sg) / np.conj(psi_site)
                if col < cols - 1:
                    e, psi = self.bond_energy((row, col), (row, col + 1), HORIZONTAL, w)
                    bond_e.append(e)
                    if psi is not None and not psi_added:
                        psi_list.append(psi)
                        psi_added = True
                    c.shift_bten_window(w.tn_h, RIGHT)
            if self.has_nnn and row < rows - 1:
                c.init_bten2(w.tn_h, LEFT, row)
                c.grow_full_bten2(w.tn_h, RIGHT, row, 2, True)
                for col in range(cols - 1):
                    e1, psi = self.nnn_energy((row, col), (row + 1, col + 1), 0, w, None)
                    e2, psi = self.nnn_energy((row + 1, col), (row, col + 1), 1, w, psi)
                    bond_e.append(e1 + e2)
                    c.shift_bten2_window(w.tn_h, RIGHT, row)
            if row < rows - 1:
                c.shift_bmps_window(w.tn_h, DOWN)
        c.generate_bmps_approach(w.tn_v, LEFT)
        for col in range(cols):
            c.init_bten(w.tn_v, UP, col)
            c.grow_full_bten(w.tn_v, DOWN, col, 2, True)
            psi_added = False
            for row in range(rows - 1):
                e, psi = self.bond_energy((row, col), (row + 1, col), VERTICAL, w)
                bond_e.append(e)
                if psi is not None and not psi_added:
                    psi_list.append(psi)
                    psi_added = True
                if row < rows - 2:
                    c.shift_bten_window(w.tn_v, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(w.tn_v, RIGHT)
        e = sum(bond_e[1:], bond_e[0]) if bond_e else 0.0
        return e + self.onsite_energy(w.config), ostar, psi_list


class SpinlessFermionModel(FermionModel):
    """SquareSpinlessFermion(t, t2, V): states 0 = occupied, 1 = empty (square_spinless_fermion.h:51-213)."""
    phys_par = (1, 0)

    def __init__(self, t, t2=0.0, V=0.0):
        self.t, self.t2, self.V = t, t2, V
        self.has_nnn = True                                 # SquareNNNModelEnergySolver<SquareSpinlessFermion>

    def diag_nn(self, c1, c2):
        return self.V * (1 - c1) * (1 - c2)

    def offdiag_nn(self, c1, c2):
        return -self.t

    def offdiag_nnn(self, c1, c2):
        return -self.t2


class tJModel(FermionModel):
    """SquaretJNNModel(t, J, mu) / SquaretJVModel: states 0 = up, 1 = down, 2 = empty (square_tJ_model.h:85-345)."""
    phys_par = (1, 1, 0)

    def __init__(self, t, J, mu=0.0, V=0.0, t2=0.0):
        self.t, self.J, self.mu, self.V, self.t2 = t, J, mu, V, t2
        self.has_nnn = t2 != 0.0

    def offdiag_nnn(self, c1, c2):
        """EvaluateNNNEnergy (square_tJ_model.h:424-460): t2 hopping only (one site empty)."""
        return -self.t2 if (c1 == 2 or c2 == 2) else 0.0

    def diag_nn(self, c1, c2):
        if c1 == c2:
            return 0.0 if c1 == 2 else self.V
        if c1 == 2 or c2 == 2:
            return 0.0
        return -0.5 * self.J + self.V

    def offdiag_nn(self, c1, c2):
        return -self.t if (c1 == 2 or c2 == 2) else 0.5 * self.J

    def onsite_energy(self, config):
        return -self.mu * float((np.asarray(config) != 2).sum()) if self.mu != 0 else 0.0


# ---------------------------------------------------------------------------------------------------------------
# exact summation (ExactSumEnergyEvaluatorMPI, exact_summation_energy_evaluator.h:150-300) for the K8 goldens
# ---------------------------------------------------------------------------------------------------------------
def exact_summation(ftps, model, configs, trunc):
    """Returns (energy, gradient[r][c][s])."""
    wsum = 0.0
    esum = 0.0
    zeros = lambda: [[[np.zeros_like(ftps.T[r][c][s]) for s in range(ftps.phys)] for c in range(ftps.cols)]
                     for r in range(ftps.rows)]
    osum, eosum = zeros(), zeros()
    for cfg in configs:
        w = FermionWalker(ftps, cfg, trunc)
        if w.amplitude == 0:
            continue
        e, ostar, _ = model.energy_and_holes(w, True)
        wt = abs(w.amplitude) ** 2
        wsum += wt
        esum += wt * e
        for r in range(ftps.rows):
            for c in range(ftps.cols):
                s = int(cfg[r][c])
                osum[r][c][s] += wt * ostar[r][c]
                eosum[r][c][s] += wt * np.conj(e) * ostar[r][c]
    energy = esum / wsum
    grad = [[[eosum[r][c][s] / wsum - np.conj(energy) * osum[r][c][s] / wsum for s in range(ftps.phys)]
             for c in range(ftps.cols)] for r in range(ftps.rows)]
    return energy, grad


def brute_force_energy(ftps, model, configs):
    """<H> from graded amplitudes and second-quantised matrix elements in row-major mode order; small lattices."""
    psi = {tuple(np.asarray(c).reshape(-1)): graded_amplitude(ftps, c) for c in configs}
    rows, cols = ftps.rows, ftps.cols
    num = 0.0
    den = 0.0
    for key, amp in psi.items():
        cfg = np.array(key).reshape(rows, cols)
        den += abs(amp) ** 2
        num += abs(amp) ** 2 * model.onsite_energy(cfg)
        pairs = []
        for r in range(rows):
            for c in range(cols):
                if c + 1 < cols: pairs.append(((r, c), (r, c + 1), 0))
                if r + 1 < rows: pairs.append(((r, c), (r + 1, c), 0))
                if model.has_nnn and r + 1 < rows and c + 1 < cols:
                    pairs.append(((r, c), (r + 1, c + 1), 1))
                    pairs.append(((r + 1, c), (r, c + 1), 1))
        for a, b, nnn in pairs:
            c1, c2 = int(cfg[a]), int(cfg[b])
            if not nnn:
                num += abs(amp) ** 2 * model.diag_nn(c1, c2)
            if c1 == c2:
                continue
            coef = model.offdiag_nnn(c1, c2) if nnn else model.offdiag_nn(c1, c2)
            if coef == 0.0:
                continue
            c2f = cfg.copy()
            c2f[a], c2f[b] = c2, c1
            k2 = tuple(c2f.reshape(-1))
            if k2 not in psi:
                continue
            sg = between_sign(ftps, cfg, a, b) if (ftps.phys_par[c1] ^ ftps.phys_par[c2]) else 1
            num += np.conj(psi[k2]) * coef * sg * amp
    return num / den

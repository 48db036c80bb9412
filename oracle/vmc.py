"""VMC sampling path restatement: walker state, NN-exchange sweep, XXZ/Heisenberg local energy + holes,
MC energy/gradient evaluator, SR S-matrix matvec.  Test infrastructure (see oracle/__init__.py).

Reference (relative to include/qlpeps/):
  * TPSWaveFunctionComponent          vmc_basic/wave_function_component.h:136-379
  * MCUpdateSquareNNUpdateBaseOBC     vmc_basic/configuration_update_strategies/square_nn_updater.h:29-81
  * MCUpdateSquareNNExchangeOBC       .../square_nn_updater.h:146-188
  * SquareNNNModelEnergySolver (NN)   algorithm/vmc_update/model_solvers/base/square_nnn_energy_solver.h:79-315
  * BondTraversalMixin (vertical)     .../model_solvers/base/bond_traversal_mixin.h:112-143
  * SquareSpinOneHalfXXZModelMixIn    .../model_solvers/square_spin_onehalf_xxz_obc.h:72-140
  * MCEnergyGradEvaluator::Evaluate   algorithm/vmc_update/mc_energy_grad_evaluator.h:152-330
  * MeanAndBinnedErrorSqrtNUniformBin vmc_basic/monte_carlo_tools/statistics.h:146-225
  * NormalizeStateOrder1              algorithm/vmc_update/monte_carlo_engine.h:206-240
  * SRSMatrix::operator*              optimizer/stochastic_reconfiguration_smatrix.h:45-91
  * SuwaTodoStateUpdate               vmc_basic/monte_carlo_tools/suwa_todo_update.h:53-113
  * MCUpdateSquareNNFullSpaceUpdateOBC .../square_nn_updater.h:253-293
  * MCUpdateSquareTNN3SiteExchange    .../square_3site_updater.h:23-160
  * TransverseFieldIsingSquareOBC     algorithm/vmc_update/model_solvers/transverse_field_ising_square_obc.h:149-247
"""
import math
import numpy as np
from .bmps import LEFT, DOWN, RIGHT, UP, HORIZONTAL, VERTICAL
from .contractor import BMPSContractor
from .mt19937 import MT19937


def project(tps, config):
    """TensorNetwork2D(sitps, config) (tensor_network_2d_basic_impl.h:24-73)."""
    return [[tps[r][c][int(config[r][c])] for c in range(len(tps[0]))] for r in range(len(tps))]


class Walker:
    """TPSWaveFunctionComponent (wave_function_component.h:136-379)."""

    def __init__(self, tps, config, trunc):
        self.config = np.array(config, dtype=np.int64)
        self.rows, self.cols = self.config.shape
        self.trunc = trunc
        self.tn = project(tps, self.config)
        self.contractor = BMPSContractor(self.rows, self.cols)
        self.contractor.init(self.tn)
        self.amplitude = 0.0
        self.evaluate_amplitude()

    def evaluate_amplitude(self):
        """EvaluateAmplitude (wave_function_component.h:187-212)."""
        c = self.contractor
        c.set_truncate_params(*self.trunc)
        c.grow_bmps_for_row(self.tn, 0)
        c.grow_full_bten(self.tn, RIGHT, 0, 2, True)
        c.init_bten(self.tn, LEFT, 0)
        self.amplitude = c.trace(self.tn, (0, 0), HORIZONTAL)
        return self.amplitude

    def update_local(self, tps, new_amplitude, *site_configs):
        """UpdateLocal / UpdateSingleSite_ (wave_function_component.h:345-378)."""
        for (site, cfg) in site_configs:
            self.config[site[0], site[1]] = cfg
            self.tn[site[0]][site[1]] = tps[site[0]][site[1]][cfg]
            self.contractor.erase_envs_after_update(site)
        self.amplitude = new_amplitude


class NNExchangeUpdater:
    """MCUpdateSquareNNExchangeOBC with the explicit-seed constructor
    (monte_carlo_sweep_updater_base.h:37, square_nn_updater.h:29-81, 146-188)."""

    def __init__(self, seed):
        self.rng = MT19937(seed)

    def two_site_update(self, site1, site2, bond_dir, tps, w):
        c1 = int(w.config[site1]); c2 = int(w.config[site2])
        if c1 == c2:
            return False
        psi_b = w.contractor.replace_nn_site_trace(w.tn, site1, site2, bond_dir,
                                                   tps[site1[0]][site1[1]][c2], tps[site2[0]][site2[1]][c1])
        psi_a = w.amplitude
        if abs(psi_b) >= abs(psi_a):
            pass
        else:
            div = abs(psi_b) / abs(psi_a)
            p = div * div
            if not (self.rng.uniform01() < p):
                return False
        w.update_local(tps, psi_b, (site1, c2), (site2, c1))
        return True

    def sweep(self, tps, w):
        """operator() (square_nn_updater.h:29-81). Returns accept_rates = [accepted / bond_num]."""
        tn, c = w.tn, w.contractor
        rows, cols = w.rows, w.cols
        accepted = 0
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(tn, UP)
        for row in range(rows):
            c.init_bten(tn, LEFT, row)
            c.grow_full_bten(tn, RIGHT, row, 2, True)
            for col in range(cols - 1):
                accepted += self.two_site_update((row, col), (row, col + 1), HORIZONTAL, tps, w)
                if col < cols - 2:
                    c.shift_bten_window(tn, RIGHT)
            if row < rows - 1:
                c.shift_bmps_window(tn, DOWN)
        c.delete_inner_bmps(LEFT)
        c.delete_inner_bmps(RIGHT)
        c.generate_bmps_approach(tn, LEFT)
        for col in range(cols):
            c.init_bten(tn, UP, col)
            c.grow_full_bten(tn, DOWN, col, 2, True)
            for row in range(rows - 1):
                accepted += self.two_site_update((row, col), (row + 1, col), VERTICAL, tps, w)
                if row < rows - 2:
                    c.shift_bten_window(tn, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(tn, RIGHT)
        c.delete_inner_bmps(UP)
        bond_num = cols * (rows - 1) + rows * (cols - 1)
        return [accepted / bond_num]


def suwa_todo_state_update(init_state, weights, rng):
    """SuwaTodoStateUpdate (suwa_todo_update.h:53-113): rejection-free geometric allocation. `weights` are doubles,
    the prefix sums and the draw are ``long double`` (x87 80-bit) exactly as in the reference."""
    ld = np.longdouble
    w = [float(x) for x in weights]
    n = len(w)
    max_id = max(range(n), key=lambda i: (w[i], -i))          # std::max_element: first maximum
    if max_id != 0:
        w[0], w[max_id] = w[max_id], w[0]
    if init_state == max_id:
        init_state = 0
    elif init_state == 0:
        init_state = max_id
    s = [ld(w[0])]
    for i in range(1, n):
        s.append(s[i - 1] + ld(w[i]))
    S = s[-1]
    s_im1 = ld(0.0) if init_state == 0 else s[init_state - 1]
    start = s_im1 + ld(w[0])
    if start >= S:
        start = start - S
    hi = np.nextafter(start + ld(w[init_state]), start)
    x = rng.uniform_long_double(start, hi)
    if x >= S:
        x = x - S
    final_state = n
    for i in range(n):                                        # std::upper_bound(s, x)
        if s[i] > x:
            final_state = i
            break
    if max_id != 0:
        if final_state == 0:
            final_state = max_id
        elif final_state == max_id:
            final_state = 0
    return final_state


def _std_norm(z):
    """std::norm: re^2 + im^2 (not |z|^2 through a hypot), so that complex weights round like the reference's."""
    if np.iscomplexobj(z):
        return float(z.real) ** 2 + float(z.imag) ** 2
    return abs(z) ** 2


class NNFullSpaceUpdater(NNExchangeUpdater):
    """MCUpdateSquareNNFullSpaceUpdateOBC (square_nn_updater.h:253-293): same bond traversal as the exchange updater,
    all d^2 local states of a bond weighted by |psi|^2, Suwa-Todo choice (one long double draw per bond)."""

    def two_site_update(self, site1, site2, bond_dir, tps, w):
        dim = len(tps[0][0])
        c1 = int(w.config[site1]); c2 = int(w.config[site2])
        init = c1 * dim + c2
        alt = [None] * (dim * dim)
        alt[init] = w.amplitude
        for a in range(dim):
            for b in range(dim):
                cfg = a * dim + b
                if cfg != init:
                    alt[cfg] = w.contractor.replace_nn_site_trace(w.tn, site1, site2, bond_dir,
                                                                  tps[site1[0]][site1[1]][a], tps[site2[0]][site2[1]][b])
        weights = [_std_norm(x / w.amplitude) for x in alt]
        final = suwa_todo_state_update(init, weights, self.rng)
        if final == init:
            return False
        w.update_local(tps, alt[final], (site1, final // dim), (site2, final % dim))
        return True


class TNN3SiteExchangeUpdater:
    """MCUpdateSquareTNN3SiteExchange (vmc_basic/configuration_update_strategies/square_3site_updater.h:23-160):
    permutations of the spins on three consecutive sites, Suwa-Todo choice; the cached amplitude is refreshed by a
    three-site trace at the start of every row / column."""

    def __init__(self, seed):
        self.rng = MT19937(seed)

    def three_site_update(self, s1, s2, s3, bond_dir, tps, w):
        spins = [int(w.config[s1]), int(w.config[s2]), int(w.config[s3])]
        if spins[0] == spins[1] == spins[2]:
            return False
        import itertools
        perms = sorted(set(itertools.permutations(sorted(spins))))          # std::next_permutation order
        init = perms.index(tuple(spins))
        psis = []
        for i, pm in enumerate(perms):
            if i != init:
                psis.append(w.contractor.replace_tnn_site_trace(w.tn, s1, bond_dir, tps[s1[0]][s1[1]][pm[0]],
                                                                tps[s2[0]][s2[1]][pm[1]], tps[s3[0]][s3[1]][pm[2]]))
            else:
                psis.append(w.amplitude)
        mx = max(abs(x) for x in psis)
        # complex / double divides the two components (std::complex operator/ with a real divisor)
        weights = [_std_norm(complex(x.real / mx, x.imag / mx) if np.iscomplexobj(x) else x / mx) for x in psis]
        final = suwa_todo_state_update(init, weights, self.rng)
        if final == init:
            return False
        pm = perms[final]
        w.update_local(tps, psis[final], (s1, pm[0]), (s2, pm[1]), (s3, pm[2]))
        return True

    def sweep(self, tps, w):
        tn, c = w.tn, w.contractor
        rows, cols = w.rows, w.cols
        accepted = 0
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(tn, UP)
        for row in range(rows):
            c.init_bten(tn, LEFT, row)
            c.grow_full_bten(tn, RIGHT, row, 3, True)
            w.amplitude = c.replace_tnn_site_trace(tn, (row, 0), HORIZONTAL, tn[row][0], tn[row][1], tn[row][2])
            for col in range(cols - 2):
                accepted += self.three_site_update((row, col), (row, col + 1), (row, col + 2), HORIZONTAL, tps, w)
                if col < cols - 3:
                    c.shift_bten_window(tn, RIGHT)
            if row < rows - 1:
                c.shift_bmps_window(tn, DOWN)
        c.delete_inner_bmps(LEFT)
        c.delete_inner_bmps(RIGHT)
        c.generate_bmps_approach(tn, LEFT)
        for col in range(cols):
            c.init_bten(tn, UP, col)
            c.grow_full_bten(tn, DOWN, col, 3, True)
            w.amplitude = c.replace_tnn_site_trace(tn, (0, col), VERTICAL, tn[0][col], tn[1][col], tn[2][col])
            for row in range(rows - 2):
                accepted += self.three_site_update((row, col), (row + 1, col), (row + 2, col), VERTICAL, tps, w)
                if row < rows - 3:
                    c.shift_bten_window(tn, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(tn, RIGHT)
        c.delete_inner_bmps(UP)
        total = cols * (rows - 2) + rows * (cols - 2)
        return [accepted / total]


class TFIMModel:
    """TransverseFieldIsingSquareOBC(h): H = -sum_<ij> sigma^z_i sigma^z_j - h sum_i sigma^x_i
    (transverse_field_ising_square_obc.h:149-247). Only the horizontal pass is needed (one-site terms)."""

    has_nnn = False

    def __init__(self, h):
        self.h = h

    def diag_energy(self, config):
        """CalDiagTermEnergy (:160-182)."""
        rows, cols = config.shape
        e = 0.0
        for row in range(rows):
            for col in range(cols - 1):
                e += -1 if config[row, col] == config[row, col + 1] else 1
        for col in range(cols):
            for row in range(rows - 1):
                e += -1 if config[row, col] == config[row + 1, col] else 1
        return e

    def measure(self, tps, w):
        """EvaluateObservables (transverse_field_ising_square_obc.h:60-137): energy, spin_z = config - 1/2, sigma_x(site) =
        -ex_term / h = conj(psi_flip / psi), SzSz_row along the middle row from (ly/2, lx/4)."""
        rec = {}
        e, _, _ = self.energy_and_holes(tps, w, False, rec=rec)
        rows, cols = w.rows, w.cols
        sz = w.config.astype(float) - 0.5
        row, c1 = rows // 2, cols // 4
        return {"energy": e, "spin_z": sz, "sigma_x": rec["sigma_x"],
                "SzSz_row": np.array([sz[row, c1] * sz[row, c1 + i] for i in range(1, cols // 2 + 1)])}

    def energy_and_holes(self, tps, w, calc_holes=True, rec=None):
        """CalEnergyAndHolesImplParsed (:208-247). Returns (E_loc, holes or None, psi_list)."""
        tn, c = w.tn, w.contractor
        rows, cols = w.rows, w.cols
        holes = [[None] * cols for _ in range(rows)] if calc_holes else None
        psi_list = []
        energy = 0.0
        if rec is not None:
            rec["sigma_x"] = np.zeros((rows, cols), dtype=np.result_type(tps[0][0][0].dtype, np.float64))
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(tn, UP)
        for row in range(rows):
            c.init_bten(tn, LEFT, row)
            c.grow_full_bten(tn, RIGHT, row, 1, True)
            psi = c.trace(tn, (row, 0), HORIZONTAL)
            inv_psi = 1.0 / psi
            psi_list.append(psi)
            for col in range(cols):
                if calc_holes:
                    holes[row][col] = np.conj(c.punch_hole(tn, (row, col), HORIZONTAL))
                cfg = int(w.config[row, col])
                psi_ex = c.replace_one_site_trace(tn, (row, col), tps[row][col][1 - cfg], HORIZONTAL)   # :191-204
                energy = energy + (-self.h) * np.conj(psi_ex * inv_psi)
                if rec is not None:
                    rec["sigma_x"][row, col] = np.conj(psi_ex * inv_psi)
                if col < cols - 1:
                    c.shift_bten_window(tn, RIGHT)
            if row < rows - 1:
                c.shift_bmps_window(tn, DOWN)
        return energy + self.diag_energy(w.config), holes, psi_list


class XXZModel:
    """SquareSpinOneHalfXXZModelOBC: H = sum_<ij> jz Sz Sz + jxy (Sx Sx + Sy Sy) - h00 Sz(0,0)
    (square_spin_onehalf_xxz_obc.h:64-159, 174-328). Heisenberg = XXZModel(1, 1, 0)."""

    def __init__(self, jz=1.0, jxy=1.0, pinning00=0.0, jz2=0.0, jxy2=0.0):
        """jz2 / jxy2 != 0 gives SquareSpinOneHalfJ1J2XXZModelOBC (square_spin_onehalf_j1j2_xxz_obc.h:34-113)."""
        self.jz, self.jxy, self.pinning00, self.jz2, self.jxy2 = jz, jxy, pinning00, jz2, jxy2
        self.has_nnn = (jz2 != 0.0 or jxy2 != 0.0)

    def nnn_energy(self, site1, site2, c1, c2, diagonal_dir, w, tps, inv_psi):
        """EvaluateNNNEnergy (square_spin_onehalf_xxz_obc.h:106-129)."""
        if c1 == c2:
            return 0.25 * self.jz2
        left_up = site1 if diagonal_dir == 0 else (site2[0], site1[1])
        psi_ex = w.contractor.replace_nnn_site_trace(w.tn, left_up, diagonal_dir, HORIZONTAL,
                                                     tps[site1[0]][site1[1]][c2], tps[site2[0]][site2[1]][c1])
        ratio = np.conj(psi_ex * inv_psi)
        return -0.25 * self.jz2 + ratio * 0.5 * self.jxy2

    def bond_energy(self, site1, site2, c1, c2, orient, w, tps, inv_psi):
        """EvaluateBondEnergy (square_spin_onehalf_xxz_obc.h:72-104)."""
        if c1 == c2:
            return 0.25 * self.jz
        psi_ex = w.contractor.replace_nn_site_trace(w.tn, site1, site2, orient,
                                                    tps[site1[0]][site1[1]][c2], tps[site2[0]][site2[1]][c1])
        ratio = np.conj(psi_ex * inv_psi)
        return -0.25 * self.jz + ratio * 0.5 * self.jxy

    def onsite_energy(self, config):
        """EvaluateTotalOnsiteEnergy (square_spin_onehalf_xxz_obc.h:131-135)."""
        return -self.pinning00 * (float(config[0, 0]) - 0.5)

    def measure(self, tps, w):
        """SquareNNNModelMeasurementSolver::EvaluateObservables (base/square_nnn_model_measurement_solver.h:33-214):
        the bond traversal of the energy solver without holes, every bond energy kept under its registry key."""
        rows, cols = w.rows, w.cols
        dt = np.result_type(tps[0][0][0].dtype, np.float64)                  # complex states: complex bond energies
        rec = {"h": np.zeros((rows, cols - 1), dt), "v": np.zeros((rows - 1, cols), dt)}
        if self.has_nnn:
            rec["dr"] = np.zeros((rows - 1, cols - 1), dt)
            rec["ur"] = np.zeros((rows - 1, cols - 1), dt)
        rec["row_corr"] = None
        e, _, _ = self.energy_and_holes(tps, w, False, rec=rec)
        out = {"energy": e, "spin_z": w.config.astype(float) - 0.5,          # CalSpinSzImpl: config - 0.5
               "bond_energy_h": rec["h"], "bond_energy_v": rec["v"]}
        if self.has_nnn:
            out["bond_energy_dr"], out["bond_energy_ur"] = rec["dr"], rec["ur"]
        # EvaluateOffDiagOrderInRow (square_spin_onehalf_xxz_obc.h:264-291): channel split by the spin at (ly/2, lx/4)
        corr = np.array(rec["row_corr"])
        zero = np.zeros_like(corr)
        if int(w.config[rows // 2, cols // 4]) == 0:
            out["SpSm_row"], out["SmSp_row"] = corr, zero
        else:
            out["SmSp_row"], out["SpSm_row"] = corr, zero
        sz = (w.config.astype(float) - 0.5).ravel()                           # SzSz_all2all, packed i <= j (:226-236)
        out["SzSz_all2all"] = np.array([sz[i] * sz[j] for i in range(sz.size) for j in range(i, sz.size)])
        return out

    def _row_corr(self, tps, w, row, inv_psi):
        """MeasureSpinOneHalfOffDiagOrderInRow (square_spin_onehalf_xxz_obc.h:22-60): flip (row, lx/4), walk right,
        one ReplaceOneSiteTrace per site whose spin differs. Leaves the tensor network as it found it."""
        tn, c = w.tn, w.contractor
        lx = w.cols
        s1 = (row, lx // 4)
        c1 = int(w.config[s1])
        tn[s1[0]][s1[1]] = tps[s1[0]][s1[1]][1 - c1]
        c.erase_envs_after_update(s1)
        c.grow_bten_step(tn, LEFT)
        c.grow_full_bten(tn, RIGHT, row, lx // 4 + 2, False)
        vals = []
        for i in range(1, lx // 2 + 1):
            s2 = (row, lx // 4 + i)
            c2 = int(w.config[s2])
            if c2 == c1:
                vals.append(0.0)
            else:
                psi_ex = c.replace_one_site_trace(tn, s2, tps[s2[0]][s2[1]][1 - c2], HORIZONTAL)
                vals.append(np.conj(psi_ex * inv_psi))
            c.shift_bten_window(tn, RIGHT)
        tn[s1[0]][s1[1]] = tps[s1[0]][s1[1]][c1]
        c.erase_envs_after_update(s1)
        return vals

    def energy_and_holes(self, tps, w, calc_holes=True, rec=None):
        """CalEnergyAndHolesImpl (square_nnn_energy_solver.h:79-101) for has_nnn=false.
        Returns (E_loc, holes[rows][cols] or None, psi_list). rec: optional per-bond record (see measure)."""
        tn, c = w.tn, w.contractor
        rows, cols = w.rows, w.cols
        bond_e = []
        psi_list = []
        holes = [[None] * cols for _ in range(rows)] if calc_holes else None
        # horizontal pass (square_nnn_energy_solver.h:104-266)
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(tn, UP)
        for row in range(rows):
            c.init_bten(tn, LEFT, row)
            c.grow_full_bten(tn, RIGHT, row, 1, True)
            psi = c.trace(tn, (row, 0), HORIZONTAL)
            if psi == 0:
                raise RuntimeError("Wavefunction amplitude is near zero, causing division by zero.")
            inv_psi = 1.0 / psi
            psi_list.append(psi)
            for col in range(cols):
                if calc_holes:
                    holes[row][col] = np.conj(c.punch_hole(tn, (row, col), HORIZONTAL))     # Dag(...)  :163
                if col < cols - 1:
                    s1, s2 = (row, col), (row, col + 1)
                    bond_e.append(self.bond_energy(s1, s2, int(w.config[s1]), int(w.config[s2]),
                                                   HORIZONTAL, w, tps, inv_psi))
                    if rec is not None:
                        rec["h"][row, col] = bond_e[-1]
                    c.shift_bten_window(tn, RIGHT)
            if self.has_nnn and row < rows - 1:               # square_nnn_energy_solver.h:203-265
                c.init_bten2(tn, LEFT, row)
                c.grow_full_bten2(tn, RIGHT, row, 2, True)
                for col in range(cols - 1):
                    s1, s2 = (row, col), (row + 1, col + 1)
                    e_nnn = self.nnn_energy(s1, s2, int(w.config[s1]), int(w.config[s2]), 0, w, tps, inv_psi)
                    s1, s2 = (row + 1, col), (row, col + 1)
                    e_ur = self.nnn_energy(s1, s2, int(w.config[s1]), int(w.config[s2]), 1, w, tps, inv_psi)
                    if rec is not None:
                        rec["dr"][row, col], rec["ur"][row, col] = e_nnn, e_ur
                    e_nnn = e_nnn + e_ur
                    bond_e.append(e_nnn)
                    c.shift_bten2_window(tn, RIGHT, row)
            if rec is not None and "row_corr" in rec and row == rows // 2:      # bond_traversal_mixin.h:96-98
                rec["row_corr"] = self._row_corr(tps, w, row, inv_psi)
            if row < rows - 1:
                c.shift_bmps_window(tn, DOWN)
        # vertical pass (bond_traversal_mixin.h:112-143)
        c.generate_bmps_approach(tn, LEFT)
        for col in range(cols):
            c.init_bten(tn, UP, col)
            c.grow_full_bten(tn, DOWN, col, 2, True)
            psi = c.trace(tn, (0, col), VERTICAL)
            if psi == 0:
                raise RuntimeError("Wavefunction amplitude is near zero, causing division by zero.")
            inv_psi = 1.0 / psi
            psi_list.append(psi)
            for row in range(rows - 1):
                s1, s2 = (row, col), (row + 1, col)
                bond_e.append(self.bond_energy(s1, s2, int(w.config[s1]), int(w.config[s2]),
                                               VERTICAL, w, tps, inv_psi))
                if rec is not None:
                    rec["v"][row, col] = bond_e[-1]
                if row < rows - 2:
                    c.shift_bten_window(tn, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(tn, RIGHT)
        e = sum(bond_e[1:], bond_e[0]) if bond_e else 0.0         # std::reduce, in order
        return e + self.onsite_energy(w.config), holes, psi_list


def measure_structure_factor(tps, w):
    """StructureFactorMeasurementMixin::MeasureStructureFactor (model_solvers/base/structure_factor_measurement_mixin.h:
    89-228), "excited state propagation": for every source (y1, x1) the UP boundary that has absorbed rows 0..y1-1 is
    forked, absorbs row y1 with S+ applied at x1 (spin-1 slice when the source spin is down), and is propagated through the
    rows y2 > y1; at each target (y2, x2) with spin up the amplitude with S- applied there is closed against the DOWN
    stack through LEFT / RIGHT environments (BMPSWalker::InitBTenLeft / TraceWithBTen / GrowBTenRightStep, bmps_walker.h).
    Returns the list of (y1, x1, y2, x2, value) in the reference's order; values are RAW overlaps (the caller divides by
    the amplitude), zero where S+ or S- annihilates the configuration."""
    from .bmps import multiply_mpo, vacuum_bmps, es
    tn, c = w.tn, w.contractor
    rows, cols = w.rows, w.cols
    dmin, dmax, terr = w.trunc
    c.set_truncate_params(*w.trunc)
    c.generate_bmps_approach(tn, UP)                 # full DOWN stack
    down = c.bmps_set[DOWN]
    main = vacuum_bmps(cols)
    out = []
    for y1 in range(rows - 1):
        for x1 in range(cols):
            src_down = int(w.config[y1, x1]) == 0
            row = [tn[y1][x] for x in range(cols)]
            if src_down:
                row[x1] = tps[y1][x1][1]
            exc = multiply_mpo(main, row, UP, dmin, dmax, terr)
            for y2 in range(y1 + 1, rows):
                bottom = down[rows - 1 - y2]
                std = [tn[y2][x] for x in range(cols)]
                # LEFT environments over the whole row (UP storage is column-reversed: AtLogicalCol)
                left = [np.ones((1, 1, 1))]
                for x in range(cols):
                    left.append(c.bten_step(left[-1], exc[cols - 1 - x], std[x], bottom[x], LEFT))
                right = np.ones((1, 1, 1))
                vals = [0.0] * cols
                for x2 in range(cols - 1, -1, -1):
                    if src_down and int(w.config[y2, x2]) == 1:
                        half = c.bten_step(left[x2], exc[cols - 1 - x2], tps[y2][x2][0], bottom[x2], LEFT)
                        vals[x2] = es("abc,cba->", half, right).item()
                    if x2 > 0:
                        right = c.bten_step(right, bottom[x2], std[x2], exc[cols - 1 - x2], RIGHT)
                for x2 in range(cols):
                    out.append((y1, x1, y2, x2, vals[x2]))
                if y2 < rows - 1:
                    exc = multiply_mpo(exc, std, UP, dmin, dmax, terr)
        main = multiply_mpo(main, [tn[y1][x] for x in range(cols)], UP, dmin, dmax, terr)
    return out


class TableModel:
    """Generic square-lattice model given by local Hamiltonian matrices, evaluated with the reference's bond traversal
    (square_nnn_energy_solver.h:79-315 + bond_traversal_mixin.h:112-143): what a user-defined `EvaluateBondEnergy /
    EvaluateNNNEnergy / EvaluateTotalOnsiteEnergy` mix-in (square_nnn_energy_solver.h:171-198) computes, with the
    matrix elements as data.  h2 / h2_nnn: (d*d, d*d) matrices in the basis p = c1*d + c2 (site1 = left / upper site of
    the bond; for the diagonals site1 = the LEFT site of the link); h1: (d, d).  E_loc = sum_terms sum_p' H[p, p'] psi(p')/psi."""

    def __init__(self, phys, h2=None, h2_nnn=None, h1=None):
        self.d, self.h2, self.h2n, self.h1 = phys, h2, h2_nnn, h1

    def _two(self, H, p, amp_of, inv_psi):
        e = H[p, p]
        for q in range(self.d * self.d):
            if q != p and H[p, q] != 0.0:
                # complex states: the reference's mix-ins conjugate every ratio (square_spin_onehalf_xxz_obc.h:72-104)
                e = e + H[p, q] * np.conj(amp_of(q // self.d, q % self.d) * inv_psi)
        return e

    def energy_and_holes(self, tps, w, calc_holes=True):
        tn, c = w.tn, w.contractor
        rows, cols, d = w.rows, w.cols, self.d
        e_tot, psi_list = 0.0, []
        holes = [[None] * cols for _ in range(rows)] if calc_holes else None
        c.set_truncate_params(*w.trunc)
        c.generate_bmps_approach(tn, UP)
        for row in range(rows):
            c.init_bten(tn, LEFT, row)
            c.grow_full_bten(tn, RIGHT, row, 1, True)
            psi = c.trace(tn, (row, 0), HORIZONTAL)
            inv_psi = 1.0 / psi
            psi_list.append(psi)
            for col in range(cols):
                if calc_holes:
                    holes[row][col] = np.conj(c.punch_hole(tn, (row, col), HORIZONTAL))
                s = (row, col)
                if self.h1 is not None:
                    p = int(w.config[s])
                    e_tot = e_tot + self.h1[p, p]
                    for q in range(d):
                        if q != p and self.h1[p, q] != 0.0:
                            e_tot = e_tot + self.h1[p, q] * np.conj(c.replace_one_site_trace(tn, s, tps[row][col][q], HORIZONTAL) * inv_psi)
                if col < cols - 1:
                    s2 = (row, col + 1)
                    if self.h2 is not None:
                        amp = lambda a, b: c.replace_nn_site_trace(tn, s, s2, HORIZONTAL, tps[row][col][a], tps[row][col + 1][b])
                        e_tot = e_tot + self._two(self.h2, int(w.config[s]) * d + int(w.config[s2]), amp, inv_psi)
                    c.shift_bten_window(tn, RIGHT)
            if self.h2n is not None and row < rows - 1:
                c.init_bten2(tn, LEFT, row)
                c.grow_full_bten2(tn, RIGHT, row, 2, True)
                for col in range(cols - 1):
                    for nnn_dir, (s1, s2) in enumerate((((row, col), (row + 1, col + 1)), ((row + 1, col), (row, col + 1)))):
                        amp = lambda a, b: c.replace_nnn_site_trace(tn, (row, col), nnn_dir, HORIZONTAL,
                                                                    tps[s1[0]][s1[1]][a], tps[s2[0]][s2[1]][b])
                        e_tot = e_tot + self._two(self.h2n, int(w.config[s1]) * d + int(w.config[s2]), amp, inv_psi)
                    c.shift_bten2_window(tn, RIGHT, row)
            if row < rows - 1:
                c.shift_bmps_window(tn, DOWN)
        c.generate_bmps_approach(tn, LEFT)
        for col in range(cols):
            c.init_bten(tn, UP, col)
            c.grow_full_bten(tn, DOWN, col, 2, True)
            psi = c.trace(tn, (0, col), VERTICAL)
            inv_psi = 1.0 / psi
            psi_list.append(psi)
            for row in range(rows - 1):
                s1, s2 = (row, col), (row + 1, col)
                if self.h2 is not None:
                    amp = lambda a, b: c.replace_nn_site_trace(tn, s1, s2, VERTICAL, tps[row][col][a], tps[row + 1][col][b])
                    e_tot = e_tot + self._two(self.h2, int(w.config[s1]) * d + int(w.config[s2]), amp, inv_psi)
                if row < rows - 2:
                    c.shift_bten_window(tn, DOWN)
            if col < cols - 1:
                c.shift_bmps_window(tn, RIGHT)
        return e_tot, holes, psi_list


def mean_and_binned_error(samples):
    """MeanAndBinnedErrorSqrtNUniformBin for one rank (statistics.h:146-225)."""
    n = len(samples)
    if n == 0:
        return 0.0, 0.0
    bin_size = max(1, int(math.sqrt(n)))
    num_bins = n // bin_size
    means = [sum(samples[i * bin_size + 1:(i + 1) * bin_size], samples[i * bin_size]) / bin_size
             for i in range(num_bins)]
    return combine_bin_means(means)


def combine_bin_means(means):
    """Master-side part of MeanAndBinnedErrorSqrtNUniformBin (statistics.h:209-223)."""
    nb = len(means)
    if nb == 0:
        return 0.0, 0.0
    mean = sum(means[1:], means[0]) / nb
    if nb == 1:
        return mean, float("inf")
    var = sum(abs(m - mean) ** 2 for m in means) / (nb - 1)
    return mean, math.sqrt(var / nb)


def tps_like_zeros(tps):
    return [[[np.zeros_like(t) for t in site] for site in row] for row in tps]


class EnergyGradEvaluator:
    """MCEnergyGradEvaluator::Evaluate for one rank == one walker (mc_energy_grad_evaluator.h:152-330).
    Multi-rank results are the rank-mean of per-rank gradients (``:296-309``) and the mean over all
    ranks' bin means for the energy; ``evaluate_ranks`` below does that combination."""

    def __init__(self, tps, model, trunc, sweeps_between_samples=1):
        self.tps, self.model, self.trunc = tps, model, trunc
        self.sweeps_between_samples = sweeps_between_samples

    def sample_rank(self, walker, updater, num_samples, collect_sr=False):
        tps = self.tps
        ostar_sum = tps_like_zeros(tps)
        eloc_ostar_sum = tps_like_zeros(tps)
        energies, ostar_samples, accept = [], [], 0.0
        for _ in range(num_samples):
            for _ in range(self.sweeps_between_samples):
                rates = updater.sweep(tps, walker)
            accept += rates[0]
            e_loc, holes, _ = self.model.energy_and_holes(tps, walker, True)
            e_conj = np.conj(e_loc)
            inv_amp = np.conj(1.0 / walker.amplitude)
            energies.append(e_loc)
            sample = {} if collect_sr else None
            for r in range(walker.rows):
                for c in range(walker.cols):
                    b = int(walker.config[r, c])
                    ostar = inv_amp * holes[r][c]                       # :259-262
                    ostar_sum[r][c][b] = ostar_sum[r][c][b] + ostar
                    eloc_ostar_sum[r][c][b] = eloc_ostar_sum[r][c][b] + e_conj * ostar
                    if collect_sr:
                        sample[(r, c)] = (b, ostar)
            if collect_sr:
                ostar_samples.append(sample)
        return dict(energies=energies, ostar_sum=ostar_sum, eloc_ostar_sum=eloc_ostar_sum,
                    accept=accept / max(1, num_samples), ostar_samples=ostar_samples, n=num_samples)

    @staticmethod
    def combine(rank_results):
        """Energy = mean of all ranks' bin means; grad = mean over ranks of
        (sum E*O*/N - conj(E) sum O*/N)  (mc_energy_grad_evaluator.h:292-309)."""
        n = rank_results[0]["n"]
        bin_size = max(1, int(math.sqrt(n)))
        means = []
        for rr in rank_results:
            e = rr["energies"]
            for i in range(len(e) // bin_size):
                chunk = e[i * bin_size:(i + 1) * bin_size]
                means.append(sum(chunk[1:], chunk[0]) / bin_size)
        energy, err = combine_bin_means(means)
        nr = len(rank_results)
        grad = tps_like_zeros(rank_results[0]["ostar_sum"])
        for rr in rank_results:
            for r, row in enumerate(grad):
                for c, site in enumerate(row):
                    for s in range(len(site)):
                        g = rr["eloc_ostar_sum"][r][c][s] * (1.0 / n) + np.conj(-energy) * (rr["ostar_sum"][r][c][s] * (1.0 / n))
                        site[s] = site[s] + g / nr
        return energy, err, grad


def normalize_state_order1(tps, amplitudes):
    """NormalizeStateOrder1 (monte_carlo_engine.h:206-240): scale = 1/max|psi|, every site tensor
    multiplied by scale^(1/(Lx*Ly)). Returns the rescaled TPS and the per-site factor."""
    rows, cols = len(tps), len(tps[0])
    scale = 1.0 / max(abs(a) for a in amplitudes)
    f = scale ** (1.0 / (rows * cols))
    return [[[t * f for t in site] for site in row] for row in tps], f


def normalize_all_site(tps):
    """SplitIndexTPS::NormalizeAllSite (split_index_tps_impl.h:209-255)."""
    out = []
    for row in tps:
        orow = []
        for site in row:
            nrm = math.sqrt(sum(float(np.sum(np.abs(t) ** 2)) for t in site))
            orow.append([t / nrm for t in site])
        out.append(orow)
    return out


def sr_matvec(ostar_samples_flat, ostar_mean_flat, v, diag_shift=0.0):
    """SRSMatrix::operator* for one rank group (stochastic_reconfiguration_smatrix.h:45-91):
    S v = (1/N) sum_i conj(O*_i . v - Obar . v) ... written for flat real/complex vectors:
    S v = mean_i O*_i (O*_i^dagger v) - Obar (Obar^dagger v) + diag_shift v."""
    n = len(ostar_samples_flat)
    acc = np.zeros_like(v)
    for o in ostar_samples_flat:
        acc = acc + o * np.vdot(o, v)
    acc = acc / n - ostar_mean_flat * np.vdot(ostar_mean_flat, v)
    return acc + diag_shift * v


def random_tps(rows, cols, phys, D, seed, signed=False, dtype=np.float64):
    """Synthetic TPS of SURVEY.md section 8d.1: i.i.d. uniform [0,1) (or [-1,1)) entries, edge legs of
    dimension 1, NormalizeAllSite applied."""
    rng = np.random.default_rng(seed)
    tps = []
    for r in range(rows):
        row = []
        for c in range(cols):
            shape = (1 if c == 0 else D, 1 if r == rows - 1 else D, 1 if c == cols - 1 else D, 1 if r == 0 else D)
            site = []
            for _ in range(phys):
                t = rng.random(shape)
                if signed:
                    t = 2.0 * t - 1.0
                site.append(t.astype(dtype))
            row.append(site)
        tps.append(row)
    return normalize_all_site(tps)


def neel_config(rows, cols):
    return np.array([[(r + c) % 2 for c in range(cols)] for r in range(rows)], dtype=np.int64)


def shuffled_half_filled_config(rows, cols, seed):
    """Per-walker initial configuration: Fisher-Yates shuffle (oracle.mt19937.shuffle_std) of the
    half-filled list with mt19937(seed)."""
    from .mt19937 import shuffle_std
    n = rows * cols
    base = [i % 2 for i in range(n)]
    return np.array(shuffle_std(base, MT19937(seed)), dtype=np.int64).reshape(rows, cols)

"""BMPSContractor restatement (dense, bosonic).  Test infrastructure (see oracle/__init__.py).

Follows two_dim_tn/tensor_network_2d/bmps/ of the reference:
  * Init / InitBMPS / InitBTen / TruncateBTen     impl/bmps_contractor_init.h:25-128
  * GenerateBMPSApproach / GrowBMPSStep / GrowFullBMPS / GrowBMPSForRow / GrowBMPSForCol /
    ShiftBMPSWindow / DeleteInnerBMPS             impl/bmps_contractor_grow.h:11-148, bmps_contractor.h:320-324
  * PunchHole                                     impl/bmps_contractor_grow.h:150-183
  * GrowFullBTen / GrowBTenStep / ShiftBTenWindow impl/bmps_contractor_grow.h:243-373, 517-582
  * Trace / ReplaceOneSiteTrace / ReplaceNNSiteTrace   impl/bmps_contractor_trace.h:11-205
  * EraseEnvsAfterUpdate                          impl/bmps_contractor_trace.h:538-589
  * BMPSAtSlice_ / BTenAtSlice_ / AtLogicalCol    bmps_contractor.h:985-1018, bmps.h:214-227

``tn`` is a rows x cols nested list of site tensors with legs (L, D, R, U)
(tensor_network_2d.h:38-46).  All three chained contractions of a BTen step are written once in the
"(pre_post, post, next_post, opposite)" frame of bmps.mpo_perm; the four position-specific code
blocks of the reference are that one pattern with the leg numbers substituted.
"""
import numpy as np
from .bmps import (LEFT, DOWN, RIGHT, UP, HORIZONTAL, VERTICAL, opposite, mpo_perm, es,
                   vacuum_bmps, multiply_mpo)


class BMPSContractor:
    def __init__(self, rows, cols):
        self.rows, self.cols = rows, cols
        self.bmps_set = {p: [] for p in range(4)}
        self.bten_set = {p: [] for p in range(4)}
        self.bten_set2 = {p: [] for p in range(4)}
        self.trunc = None
        self.n_multiply = 0          # instrumentation: number of MultiplyMPO calls

    # ---- parameters (bmps_contractor.h:216-230)
    def set_compress_scheme(self, scheme, tol=1e-12, max_iter=10):
        """BMPSTruncateParams::Variational2Site / Variational1Site (bmps.h:81-97): 0 SVD, 1 two-site, 2 one-site."""
        self.scheme, self.var_tol, self.var_iter = scheme, tol, max_iter

    def set_truncate_params(self, dmin, dmax, trunc_err):
        self.trunc = (int(dmin), int(dmax), float(trunc_err))

    # ---- init (impl/bmps_contractor_init.h:25-70)
    def init(self, tn):
        for p in range(4):
            n = self.cols if p in (UP, DOWN) else self.rows
            self.bmps_set[p] = [vacuum_bmps(n)]
            self.bten_set[p] = []
            self.bten_set2[p] = []

    # ---- accessors (bmps_contractor.h:985-1018)
    def bmps_at_slice(self, pos, logical_idx):
        if pos == DOWN:
            return self.bmps_set[DOWN][self.rows - 1 - logical_idx]
        if pos == RIGHT:
            return self.bmps_set[RIGHT][self.cols - 1 - logical_idx]
        return self.bmps_set[pos][logical_idx]

    def bten_at_slice(self, pos, logical_idx):
        if pos == DOWN:
            return self.bten_set[DOWN][self.rows - 1 - logical_idx]
        if pos == RIGHT:
            return self.bten_set[RIGHT][self.cols - 1 - logical_idx]
        return self.bten_set[pos][logical_idx]

    @staticmethod
    def at_logical_col(bmps, pos, col):
        """BMPS::AtLogicalCol (bmps.h:214-227)."""
        return bmps[len(bmps) - 1 - col] if pos in (UP, RIGHT) else bmps[col]

    # ---- BMPS growth (impl/bmps_contractor_grow.h:11-148)
    def _slice(self, tn, num, orient):
        """TenMatrix::get_slice (framework/duomatrix.h:312-318)."""
        if orient == HORIZONTAL:
            return [tn[num][c] for c in range(self.cols)]
        return [tn[r][num] for r in range(self.rows)]

    def _grow_with_mpo(self, pos, mpo):
        dmin, dmax, terr = self.trunc
        stack = self.bmps_set[pos]
        if getattr(self, "scheme", 0) == 0:
            stack.append(multiply_mpo(stack[-1], mpo, pos, dmin, dmax, terr))
        else:                                                # CompressMPSScheme::VARIATION2Site (1) / VARIATION1Site (2)
            from .bmps import multiply_mpo_variational
            stack.append(multiply_mpo_variational(stack[-1], mpo, pos, dmin, dmax, terr, self.var_tol, self.var_iter,
                                                  one_site=self.scheme == 2))
        self.n_multiply += 1
        return len(stack)

    def grow_bmps_step(self, tn, pos):
        """GrowBMPSStep(tn, position) (grow.h:32-48)."""
        existed = len(self.bmps_set[pos])
        assert existed > 0
        if pos in (UP, LEFT):
            mpo_num = existed - 1
        elif pos == DOWN:
            mpo_num = self.rows - existed
        else:
            mpo_num = self.cols - existed
        orient = HORIZONTAL if pos in (UP, DOWN) else VERTICAL   # Rotate(Orientation(position))
        return self._grow_with_mpo(pos, self._slice(tn, mpo_num, orient))

    def grow_full_bmps(self, tn, pos):
        """GrowFullBMPS (grow.h:50-86)."""
        existed = len(self.bmps_set[pos])
        assert existed > 0
        if pos == DOWN:
            for row in range(self.rows - existed, 0, -1):
                self._grow_with_mpo(pos, self._slice(tn, row, HORIZONTAL))
        elif pos == UP:
            for row in range(existed - 1, self.rows - 1):
                self._grow_with_mpo(pos, self._slice(tn, row, HORIZONTAL))
        elif pos == LEFT:
            for col in range(existed - 1, self.cols - 1):
                self._grow_with_mpo(pos, self._slice(tn, col, VERTICAL))
        else:
            for col in range(self.cols - existed, 0, -1):
                self._grow_with_mpo(pos, self._slice(tn, col, VERTICAL))

    def delete_inner_bmps(self, pos):
        """DeleteInnerBMPS (bmps_contractor.h:320-324)."""
        if self.bmps_set[pos]:
            del self.bmps_set[pos][1:]

    def generate_bmps_approach(self, tn, post):
        """GenerateBMPSApproach (grow.h:11-17)."""
        self.delete_inner_bmps(post)
        self.grow_full_bmps(tn, opposite(post))

    def grow_bmps_for_row(self, tn, row):
        """GrowBMPSForRow (grow.h:88-104)."""
        for row_bmps in range(self.rows - len(self.bmps_set[DOWN]), row, -1):
            self._grow_with_mpo(DOWN, self._slice(tn, row_bmps, HORIZONTAL))
        for row_bmps in range(len(self.bmps_set[UP]) - 1, row):
            self._grow_with_mpo(UP, self._slice(tn, row_bmps, HORIZONTAL))

    def grow_bmps_for_col(self, tn, col):
        """GrowBMPSForCol (grow.h:106-122)."""
        for col_bmps in range(self.cols - len(self.bmps_set[RIGHT]), col, -1):
            self._grow_with_mpo(RIGHT, self._slice(tn, col_bmps, VERTICAL))
        for col_bmps in range(len(self.bmps_set[LEFT]) - 1, col):
            self._grow_with_mpo(LEFT, self._slice(tn, col_bmps, VERTICAL))

    def shift_bmps_window(self, tn, pos):
        """ShiftBMPSWindow (grow.h:143-148)."""
        self.bmps_set[pos].pop()
        self.grow_bmps_step(tn, opposite(pos))

    # ---- BTen (impl/bmps_contractor_init.h:72-128, grow.h:243-373, 517-582)
    def init_bten(self, tn, pos, slice_num):
        self.bten_set[pos] = [np.ones((1, 1, 1))]

    def truncate_bten(self, pos, length):
        if len(self.bten_set[pos]) > length:
            del self.bten_set[pos][length:]

    @staticmethod
    def bten_step(bten, mps1, site, mps2, post):
        """The three chained contractions of a BTen step (grow.h:577-579):
        Contract(mps1, bten, 2,0,1); Contract(tmp1, site, 1, pre_post, 2); Contract(tmp2,{0,2}, mps2,{0,1}).
        Result legs (mps1[0], site[opposite(post)], mps2[2])."""
        m = np.transpose(site, mpo_perm(post))               # m[p1(pre_post), y(post), f(next_post), o(opposite)]
        tmp1 = es("apx,xyz->apyz", mps1, bten)
        tmp2 = es("apyz,pyfo->zafo", tmp1, m)
        return es("zafo,zfb->aob", tmp2, mps2)

    def _bten_operands(self, tn, post, slice_num, bten_size):
        """mps_ten1 / mps_ten2 / grown site of one BTen step at cache size ``bten_size`` on slice
        ``slice_num`` (GrowFullBTen grow.h:262-372 == GrowBTenStep grow.h:541-575)."""
        pre_post, next_post = (post + 3) % 4, (post + 1) % 4
        n = self.cols if post in (LEFT, RIGHT) else self.rows
        b1 = self.bmps_at_slice(pre_post, slice_num)
        b2 = self.bmps_at_slice(next_post, slice_num)
        mps1 = b1[n - bten_size]
        mps2 = b2[bten_size - 1]
        if post == LEFT:
            site = (slice_num, bten_size - 1)
        elif post == RIGHT:
            site = (slice_num, n - bten_size)
        elif post == UP:
            site = (bten_size - 1, slice_num)
        else:
            site = (n - bten_size, slice_num)
        return mps1, mps2, site

    def grow_full_bten(self, tn, pos, slice_num, remain_sites, init):
        """GrowFullBTen (grow.h:243-373)."""
        if init:
            self.init_bten(tn, pos, slice_num)
        btens = self.bten_set[pos]
        n = self.cols if pos in (LEFT, RIGHT) else self.rows
        start_idx = len(btens) - 1
        for i in range(start_idx, n - remain_sites):
            mps1, mps2, site = self._bten_operands(tn, pos, slice_num, i + 1)
            btens.append(self.bten_step(btens[-1], mps1, tn[site[0]][site[1]], mps2, pos))

    def grow_bten_step(self, tn, post):
        """GrowBTenStep (grow.h:529-582): slice is implied by the BMPS stack sizes."""
        if post in (LEFT, RIGHT):
            slice_num = len(self.bmps_set[UP]) - 1
            assert len(self.bmps_set[UP]) + len(self.bmps_set[DOWN]) == self.rows + 1
        else:
            slice_num = len(self.bmps_set[LEFT]) - 1
            assert len(self.bmps_set[LEFT]) + len(self.bmps_set[RIGHT]) == self.cols + 1
        btens = self.bten_set[post]
        mps1, mps2, site = self._bten_operands(tn, post, slice_num, len(btens))
        btens.append(self.bten_step(btens[-1], mps1, tn[site[0]][site[1]], mps2, post))

    def shift_bten_window(self, tn, pos):
        """ShiftBTenWindow (grow.h:517-521)."""
        self.bten_set[pos].pop()
        self.grow_bten_step(tn, opposite(pos))

    # ---- traces (impl/bmps_contractor_trace.h:11-205)
    def trace(self, tn, site_a, bond_dir):
        """Trace(tn, site_a, bond_dir) (trace.h:11-28)."""
        r, c = site_a
        site_b = (r, c + 1) if bond_dir == HORIZONTAL else (r + 1, c)
        return self.replace_nn_site_trace(tn, site_a, site_b, bond_dir, tn[r][c], tn[site_b[0]][site_b[1]])

    def replace_nn_site_trace(self, tn, site_a, site_b, bond_dir, ten_a, ten_b):
        """ReplaceNNSiteTrace (trace.h:90-205): two half-environments and a rank-3 inner product."""
        if bond_dir == HORIZONTAL:
            row, col_a, col_b = site_a[0], site_a[1], site_b[1]
            first, second, slice_num, ia, ib = LEFT, RIGHT, row, col_a, col_b
            n = self.cols
        else:
            col, row_a, row_b = site_a[1], site_a[0], site_b[0]
            first, second, slice_num, ia, ib = UP, DOWN, col, row_a, row_b
            n = self.rows
        # first half: <bten[first][ia]| absorbing ten_a   (trace.h:129-131 / 165-167)
        mps1, mps2, _ = self._bten_operands(tn, first, slice_num, ia + 1)
        half_a = self.bten_step(self.bten_set[first][ia], mps1, ten_a, mps2, first)
        # second half: |bten[second] at ib> absorbing ten_b  (trace.h:149-156 / 184-191)
        mps1, mps2, _ = self._bten_operands(tn, second, slice_num, n - ib)
        half_b = self.bten_step(self.bten_at_slice(second, ib), mps1, ten_b, mps2, second)
        return es("abc,cba->", half_a, half_b).item()        # trace.h:202

    def replace_one_site_trace(self, tn, site, replace_ten, mps_orient):
        """ReplaceOneSiteTrace (trace.h:30-88)."""
        row, col = site
        if mps_orient == HORIZONTAL:
            first, second, slice_num, i = LEFT, RIGHT, row, col
        else:
            first, second, slice_num, i = UP, DOWN, col, row
        mps1, mps2, _ = self._bten_operands(tn, first, slice_num, i + 1)
        half_a = self.bten_step(self.bten_set[first][i], mps1, replace_ten, mps2, first)
        return es("abc,cba->", half_a, self.bten_at_slice(second, i)).item()

    def replace_tnn_site_trace(self, tn, site0, mps_orient, ten0, ten1, ten2):
        """ReplaceTNNSiteTrace (trace.h:326-420): three consecutive sites of a row (column) replaced; three BTen steps
        from the LEFT (UP) environment closed with the RIGHT (DOWN) one."""
        row, col = site0
        if mps_orient == HORIZONTAL:
            first, second, slice_num, i = LEFT, RIGHT, row, col
        else:
            first, second, slice_num, i = UP, DOWN, col, row
        cur = self.bten_set[first][i]
        for step, ten in enumerate((ten0, ten1, ten2)):
            mps1, mps2, _ = self._bten_operands(tn, first, slice_num, i + step + 1)
            cur = self.bten_step(cur, mps1, ten, mps2, first)
        return es("abc,cba->", cur, self.bten_at_slice(second, i + 2)).item()

    def punch_hole(self, tn, site, mps_orient):
        """PunchHole (grow.h:150-183), bosonic branch; result legs (L, D, R, U)."""
        row, col = site
        if mps_orient == HORIZONTAL:
            up_ten = self.at_logical_col(self.bmps_at_slice(UP, row), UP, col)
            down_ten = self.at_logical_col(self.bmps_at_slice(DOWN, row), DOWN, col)
            left_ten = self.bten_set[LEFT][col]
            right_ten = self.bten_at_slice(RIGHT, col)
        else:
            up_ten = self.bten_set[UP][row]
            down_ten = self.bten_at_slice(DOWN, row)
            left_ten = self.at_logical_col(self.bmps_at_slice(LEFT, col), LEFT, row)
            right_ten = self.at_logical_col(self.bmps_at_slice(RIGHT, col), RIGHT, row)
        tmp1 = es("xlz,zdb->xldb", left_ten, down_ten)       # Contract(left,{2}, down,{0})
        tmp2 = es("brz,zux->brux", right_ten, up_ten)        # Contract(right,{2}, up,{0})
        return es("xldb,brux->ldru", tmp1, tmp2)             # Contract(tmp1,{0,3}, tmp2,{3,0})

    # ---- invalidation (impl/bmps_contractor_trace.h:538-589)
    def erase_envs_after_update(self, site):
        row, col = site
        del self.bmps_set[LEFT][col + 1:]
        del self.bmps_set[UP][row + 1:]
        del self.bmps_set[DOWN][self.rows - row:]
        del self.bmps_set[RIGHT][self.cols - col:]
        del self.bten_set[LEFT][col + 1:]
        del self.bten_set[UP][row + 1:]
        del self.bten_set[RIGHT][self.cols - col:]
        del self.bten_set[DOWN][self.rows - row:]
        del self.bten_set2[LEFT][col + 1:]
        del self.bten_set2[UP][row + 1:]
        del self.bten_set2[RIGHT][self.cols - col:]
        del self.bten_set2[DOWN][self.rows - row:]

    # ---- two-row environments for next-nearest-neighbour terms
    #      (impl/bmps_contractor_init.h:130-186, grow.h:375-527, helpers.h:151-180, trace.h:207-324)
    def init_bten2(self, tn, pos, slice_num1):
        self.bten_set2[pos] = [np.ones((1, 1, 1, 1))]

    @staticmethod
    def bten2_step(bten2, mps1, site1, site2, mps2, post):
        """GrowBTen2StepAfterTransposedMPOTens (helpers.h:151-180) == the loop body of GrowFullBTen2 (grow.h:497-511).
        site1 is the tensor adjacent to the BMPS at pre_post, site2 the one adjacent to the BMPS at next_post.
        Result legs (mps1[0], site1[opposite], site2[opposite], mps2[2])."""
        pre, nxt, opp = (post + 3) % 4, (post + 1) % 4, (post + 2) % 4
        s1 = np.transpose(site1, (pre, post, opp, nxt))          # GenMpoTen1TransposeAxesForBrowBTen2
        s2 = np.transpose(site2, (pre, post, nxt, opp))
        tmp1 = es("apx,xyvz->apyvz", mps1, bten2)
        tmp2 = es("apyvz,pyon->vzaon", tmp1, s1)
        tmp3 = es("vzaon,nvfq->zaofq", tmp2, s2)
        return es("zaofq,zfb->aoqb", tmp3, mps2)

    def _bten2_operands(self, tn, post, slice_num1, bten_size):
        """(mps1, mps2, site1, site2) of one BTen2 step (SetUpCoordInfoForGrowBTen2 / GrowFullBTen2 grow.h:389-437)."""
        pre_post, next_post = (post + 3) % 4, (post + 1) % 4
        n = self.cols if post in (LEFT, RIGHT) else self.rows
        s1, s2 = slice_num1, slice_num1 + 1
        if post == LEFT:      # pre = UP (row1), next = DOWN (row2)
            b1, b2 = self.bmps_at_slice(UP, s1), self.bmps_at_slice(DOWN, s2)
            site1, site2 = (s1, bten_size - 1), (s2, bten_size - 1)
        elif post == RIGHT:   # pre = DOWN (row2), next = UP (row1)
            b1, b2 = self.bmps_at_slice(DOWN, s2), self.bmps_at_slice(UP, s1)
            site1, site2 = (s2, n - bten_size), (s1, n - bten_size)
        elif post == UP:      # pre = RIGHT (col2), next = LEFT (col1)
            b1, b2 = self.bmps_at_slice(RIGHT, s2), self.bmps_at_slice(LEFT, s1)
            site1, site2 = (bten_size - 1, s2), (bten_size - 1, s1)
        else:                 # DOWN: pre = LEFT (col1), next = RIGHT (col2)
            b1, b2 = self.bmps_at_slice(LEFT, s1), self.bmps_at_slice(RIGHT, s2)
            site1, site2 = (n - bten_size, s1), (n - bten_size, s2)
        return b1[n - bten_size], b2[bten_size - 1], site1, site2

    def grow_full_bten2(self, tn, post, slice_num1, remain_sites, init):
        """GrowFullBTen2 (grow.h:375-515)."""
        if init:
            self.init_bten2(tn, post, slice_num1)
        btens = self.bten_set2[post]
        n = self.cols if post in (LEFT, RIGHT) else self.rows
        for i in range(len(btens) - 1, n - remain_sites):
            m1, m2, s1, s2 = self._bten2_operands(tn, post, slice_num1, i + 1)
            btens.append(self.bten2_step(btens[-1], m1, tn[s1[0]][s1[1]], tn[s2[0]][s2[1]], m2, post))

    def grow_bten2_step(self, tn, post, slice_num1):
        """GrowBTen2Step (grow.h:447-493)."""
        btens = self.bten_set2[post]
        m1, m2, s1, s2 = self._bten2_operands(tn, post, slice_num1, len(btens))
        btens.append(self.bten2_step(btens[-1], m1, tn[s1[0]][s1[1]], tn[s2[0]][s2[1]], m2, post))

    def shift_bten2_window(self, tn, pos, slice_num1):
        """ShiftBTen2Window (grow.h:523-527)."""
        self.bten_set2[pos].pop()
        self.grow_bten2_step(tn, opposite(pos), slice_num1)

    def bten2_at_slice(self, pos, logical_idx):
        if pos == DOWN:
            return self.bten_set2[DOWN][self.rows - 1 - logical_idx]
        if pos == RIGHT:
            return self.bten_set2[RIGHT][self.cols - 1 - logical_idx]
        return self.bten_set2[pos][logical_idx]

    def replace_nnn_site_trace(self, tn, left_up_site, nnn_dir, mps_orient, ten_left, ten_right):
        """ReplaceNNNSiteTrace (trace.h:207-324), both MPS orientations. nnn_dir 0 = LEFTUP_TO_RIGHTDOWN (ten_left
        replaces (row1,col1), ten_right replaces (row2,col2)); 1 = LEFTDOWN_TO_RIGHTUP (ten_left replaces (row2,col1),
        ten_right replaces (row1,col2)). HORIZONTAL closes the two-row environments LEFT | RIGHT of rows row1,row2
        (:218-281); VERTICAL the two-column environments UP | DOWN of columns col1,col2 (:282-324)."""
        row1, col1 = left_up_site
        row2, col2 = row1 + 1, col1 + 1
        t = {(row1, col1): tn[row1][col1], (row2, col1): tn[row2][col1],
             (row1, col2): tn[row1][col2], (row2, col2): tn[row2][col2]}
        if nnn_dir == 0:
            t[(row1, col1)], t[(row2, col2)] = ten_left, ten_right
        else:
            t[(row2, col1)], t[(row1, col2)] = ten_left, ten_right
        if mps_orient == HORIZONTAL:
            n = self.cols
            m1, m2, _, _ = self._bten2_operands(tn, LEFT, row1, col1 + 1)
            half_a = self.bten2_step(self.bten_set2[LEFT][col1], m1, t[(row1, col1)], t[(row2, col1)], m2, LEFT)
            m1, m2, _, _ = self._bten2_operands(tn, RIGHT, row1, n - col2)
            half_b = self.bten2_step(self.bten2_at_slice(RIGHT, col2), m1, t[(row2, col2)], t[(row1, col2)], m2, RIGHT)
        else:
            n = self.rows
            m1, m2, _, _ = self._bten2_operands(tn, UP, col1, row1 + 1)
            half_a = self.bten2_step(self.bten_set2[UP][row1], m1, t[(row1, col2)], t[(row1, col1)], m2, UP)
            m1, m2, _, _ = self._bten2_operands(tn, DOWN, col1, n - row2)
            half_b = self.bten2_step(self.bten2_at_slice(DOWN, row2), m1, t[(row2, col1)], t[(row2, col2)], m2, DOWN)
        return es("aoqb,bqoa->", half_a, half_b).item()      # Contract(tmp[3],{0,1,2,3}, tmp[7],{3,2,1,0})

    def replace_sqrt5_dist_two_site_trace(self, tn, left_up_site, link_dir, mps_orient, ten_left, ten_right):
        """ReplaceSqrt5DistTwoSiteTrace (trace.h:426-536): the two sites at the far corners of a 2 x 3 (HORIZONTAL) or
        3 x 2 (VERTICAL) plaquette replaced. link_dir 0 = LEFTUP_TO_RIGHTDOWN: ten_left at (row1,col1), ten_right at the
        opposite lower-right corner; 1 = LEFTDOWN_TO_RIGHTUP: ten_left at the lower-left corner, ten_right at the upper
        right one. Two environment steps from the first side (:463-466 + :473-476), one from the other (:468-471),
        closed by Contract(tmp[11],{0,1,2,3}, tmp[7],{3,2,1,0}) (:534)."""
        row1, col1 = left_up_site
        if mps_orient == HORIZONTAL:
            row2, col2, col3 = row1 + 1, col1 + 1, col1 + 2
            t = {(r, c): tn[r][c] for r in (row1, row2) for c in (col1, col2, col3)}
            if link_dir == 0:
                t[(row1, col1)], t[(row2, col3)] = ten_left, ten_right
            else:
                t[(row2, col1)], t[(row1, col3)] = ten_left, ten_right
            n = self.cols
            m1, m2, _, _ = self._bten2_operands(tn, LEFT, row1, col1 + 1)
            a = self.bten2_step(self.bten_set2[LEFT][col1], m1, t[(row1, col1)], t[(row2, col1)], m2, LEFT)
            m1, m2, _, _ = self._bten2_operands(tn, LEFT, row1, col2 + 1)
            a = self.bten2_step(a, m1, t[(row1, col2)], t[(row2, col2)], m2, LEFT)
            m1, m2, _, _ = self._bten2_operands(tn, RIGHT, row1, n - col3)
            b = self.bten2_step(self.bten2_at_slice(RIGHT, col3), m1, t[(row2, col3)], t[(row1, col3)], m2, RIGHT)
        else:
            row2, row3, col2 = row1 + 1, row1 + 2, col1 + 1
            t = {(r, c): tn[r][c] for r in (row1, row2, row3) for c in (col1, col2)}
            if link_dir == 0:
                t[(row1, col1)], t[(row3, col2)] = ten_left, ten_right
            else:
                t[(row3, col1)], t[(row1, col2)] = ten_left, ten_right
            n = self.rows
            m1, m2, _, _ = self._bten2_operands(tn, UP, col1, row1 + 1)
            a = self.bten2_step(self.bten_set2[UP][row1], m1, t[(row1, col2)], t[(row1, col1)], m2, UP)
            m1, m2, _, _ = self._bten2_operands(tn, UP, col1, row2 + 1)
            a = self.bten2_step(a, m1, t[(row2, col2)], t[(row2, col1)], m2, UP)
            m1, m2, _, _ = self._bten2_operands(tn, DOWN, col1, n - row3)
            b = self.bten2_step(self.bten2_at_slice(DOWN, row3), m1, t[(row3, col1)], t[(row3, col2)], m2, DOWN)
        return es("aoqb,bqoa->", a, b).item()

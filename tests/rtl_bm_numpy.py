"""Second, independent reading of the FPGA block matcher in vectorised numpy.

TEST INFRASTRUCTURE.  Written from the Verilog alone (not from oracle/u96_oracle.c) so that two readers of the RTL can be
compared at every disparity range the build supports -- the reference ships no disparity dump, so this is a cross-check,
not a pin.  Every block cites the RTL lines it restates (paths relative to /root/reference/src/dvp/rtl/).

Differences in METHOD from the C oracle on purpose: whole image rows at a time instead of pixel by pixel, one global
array of D+2 disparity lanes instead of 34-lane dphases (the guard lanes 0 / 33 of a dphase are the regular lanes of its
neighbours: the column sum of an (x, d) pair does not depend on the pass it is computed in), horizontal window sums by
prefix sums instead of the sliding add/subtract, neighbours of the winner gathered by index instead of carried through
the tournament, dividers by their closed forms.
"""
import numpy as np


def _tournament(S):
    """bm_calc_det.v:124-426.  S: (32, n) window sums of lanes 1..32 of one dphase.
    Returns idx1, min1, idx2, min2 (the approximate second minimum)."""
    n = S.shape[1]
    ar = np.arange(n)
    # stage 1 (:124-141): pairs, right wins only when strictly smaller
    c1 = S[1::2] < S[0::2]                                         # (16, n)
    v1 = np.where(c1, S[1::2], S[0::2])
    i1 = 2 * np.arange(16)[:, None] + c1
    # stage 2 (:168-229)
    c2 = v1[1::2] < v1[0::2]                                       # (8, n)
    v2 = np.where(c2, v1[1::2], v1[0::2])
    i2 = np.where(c2, i1[1::2], i1[0::2])
    # stage 3 (:234-268)
    c3 = v2[1::2] < v2[0::2]                                       # (4, n)
    v3 = np.where(c3, v2[1::2], v2[0::2])
    i3 = np.where(c3, i2[1::2], i2[0::2])
    # stage 4 (:273-316): winners and losers of the two halves
    c4 = v3[1::2] < v3[0::2]                                       # (2, n)
    w4 = np.where(c4, v3[1::2], v3[0::2]); wi4 = np.where(c4, i3[1::2], i3[0::2])
    l4 = np.where(c4, v3[0::2], v3[1::2]); li4 = np.where(c4, i3[0::2], i3[1::2])
    # stage 5 (:321-378): final winner; candidate 0 = loser of the final, candidate 1 = smaller stage-4 loser
    c5 = w4[1] < w4[0]
    min1 = np.where(c5, w4[1], w4[0]); idx1 = np.where(c5, wi4[1], wi4[0])
    m2a = np.where(c5, w4[0], w4[1]); i2a = np.where(c5, wi4[0], wi4[1])
    cl = l4[1] < l4[0]
    m2b = np.where(cl, l4[1], l4[0]); i2b = np.where(cl, li4[1], li4[0])
    # stage 6 (:382-416): 6-bit "+1" never wraps, adjacency = indices differ by exactly one
    adj_a = (i2a == idx1 + 1) | (idx1 == i2a + 1)
    adj_b = (i2b == idx1 + 1) | (idx1 == i2b + 1)
    take_b = ((m2b < m2a) & ~adj_b) | adj_a
    del ar
    return idx1, min1, np.where(take_b, i2b, i2a), np.where(take_b, m2b, m2a)


def _frac(L, R, C):
    """bm_calc_frac.v:63-173 with diven#(18,18,8,17) == floor(128*num/den).  Returns the 8-bit two's
    complement fraction as a signed integer in [-64, 64]."""
    cmp = L < R
    neg = (L < C) | (R < C)                                        # borrow bits of L-C, R-C (:67-72)
    num = np.where(neg, 0, L - R)
    den = 2 * np.where(cmp, R - C, L - C)                          # :100-113
    safe = np.where(den == 0, 1, den)
    q = np.floor_divide(128 * num, safe)                           # numpy floor_divide floors toward -inf, like the divider
    q = np.where(num == 0, 0, q)                                   # 0 / negative divisor -> 0
    return np.where(den == 0, np.where(cmp, 64, -64), q)           # :156-163 (checked on the divisor alone)


def bm_rtl_numpy(xl, xr, wsz=21, ndisp=64, uni_enb=0, uni_mode=0, uni_thr=0, x_store_offset=1, rtl_extended=0):
    xl = (np.asarray(xl).astype(np.int64)) & 0x3F                   # lr_din is 6 bit (bm_calc_sad.v:82-101)
    xr = (np.asarray(xr).astype(np.int64)) & 0x3F
    H, W = xl.shape
    h = wsz >> 1                                                   # bm.v:246
    x0, x1 = ndisp, W - 2                                          # HSAD columns, hsad_wdt = W - ndisp - 1 (bm.v:249)
    ncol = x1 - x0 + 1
    nctr = ncol - 2 * h                                            # sad_wdt (bm.v:252)
    rows = H - 2 * h                                               # sad_hgt (bm.v:255)
    out = np.full((H, W), -1, np.int16)                            # firmware memset 0xFF (fpga.c:105-106)
    if nctr <= 0 or rows <= 0:
        return out
    lanes = np.arange(-1, ndisp + 1)                               # d = -1 .. D (guard lanes at both ends of every dphase)
    src = np.arange(x0, x1 + 1)[None, :] - lanes[:, None]          # R column of (lane, x): x - d

    def AD(y):                                                     # bm_calc_sad.v:353-418
        return np.abs(xl[y, x0:x1 + 1][None, :] - xr[y][src])

    col = AD(0)                                                    # first_line (:454)
    for y in range(1, wsz):
        col = np.minimum(1023, col + AD(y))                        # upper_lim10 (:455-457)
    nd = ndisp // 32
    for i in range(rows):
        if i > 0:                                                  # op 1 then op 0/2 (bm_ibuf.v:195-248)
            col = np.maximum(0, col - AD(i - 1))                   # lower_lim10 (:459-462)
            col = np.minimum(1023, col + AD(i + wsz - 1))
        cs = np.concatenate([np.zeros((col.shape[0], 1), np.int64), np.cumsum(col, axis=1)], axis=1)
        sad = cs[:, wsz:wsz + nctr] - cs[:, 0:nctr]                # (D+2, nctr) window sums (bm_calc_sad.v:501-605)
        ar = np.arange(nctr)
        s_min1 = s_min2 = s_d1 = s_d2 = s_q = None
        for p in range(nd):
            S = sad[32 * p + 1:32 * p + 33]                        # lanes 1..32 <-> d = 32p .. 32p+31
            idx1, min1, idx2, min2 = _tournament(S)
            L = sad[32 * p + idx1, ar]                             # lane idx1+1-1 of the global array = d-1
            R = sad[32 * p + idx1 + 2, ar]                         # d+1
            q = _frac(L, R, min1)
            d1 = (p << 5) | idx1; d2 = (p << 5) | idx2             # bm_calc_upd.v:119-123
            if p == 0:                                             # ~mode (:146-153)
                s_min1, s_min2, s_d1, s_d2, s_q = min1, min2, d1, d2, q
                continue
            a = min1 < s_min1; b = min2 < s_min1; c = min1 < s_min2; e = min2 < s_min2     # :134-137
            adj = (d1 & 0xFF) == ((s_d1 + 1) & 0xFF)               # :138, 8-bit add
            # casex table :155-199
            r11 = a & b
            r10_1 = a & ~b & e
            r10_0 = a & ~b & ~e
            r0_11 = ~a & c & e
            r0_10 = ~a & c & ~e
            n_min2 = np.select([r11, r10_1, r10_0, r0_11, r0_10],
                               [min2, np.where(adj, min2, s_min1), np.where(adj, s_min2, s_min1),
                                np.where(adj, min2, min1), np.where(adj, s_min2, min1)], s_min2)
            n_d2 = np.select([r11, r10_1, r10_0, r0_11, r0_10],
                             [d2, np.where(adj, d2, s_d1), np.where(adj, s_d2, s_d1),
                              np.where(adj, d2, d1), np.where(adj, s_d2, d1)], s_d2)
            s_q = np.where(a, q, s_q)                              # bm_calc.v:313: fraction follows min1
            s_d1 = np.where(a, d1, s_d1)
            s_min1 = np.where(a, min1, s_min1)
            s_min2, s_d2 = n_min2, n_d2
        disp, frac = s_d1.copy(), s_q.copy()
        if uni_enb:                                                # bm_calc_uni.v:120-134, bm_calc.v:315-328 (last dphase only)
            ratio = np.where(s_min2 == 0, 2047, (1024 * s_min1) // np.where(s_min2 == 0, 1, s_min2)) & 0x3FF
            act = ratio > uni_thr
            disp = np.where(act, 0xFF if uni_mode else 0, disp)
            frac = np.where(act, -1 if uni_mode else 0, frac)      # 8'hFF as a signed fraction
        depth = disp * 256 + frac                                  # bm_obuf2.v:134-143 (17 bit signed)
        if rtl_extended:
            val = depth >> 4
        else:                                                      # {{4{depth[15]}}, depth[15:4]} (:150)
            val = ((depth >> 4) & 0x0FFF) | np.where(depth & 0x8000, 0xF000, 0)
            val = np.where(val >= 0x8000, val - 0x10000, val)
        val = np.where(depth <= 0, -1, val)                        # negative or zero -> 0xFFFF (:146-149)
        xs = ndisp + h + x_store_offset                            # first stored column (bm_obuf2.v:125-127; A1)
        out[h + i, xs:xs + nctr] = val.astype(np.int16)[:max(0, min(nctr, W - xs))]
    return out

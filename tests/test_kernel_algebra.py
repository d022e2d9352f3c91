"""CPU checks of two arithmetic identities the fused BM kernel (u96_slam_b200/csrc/bm_fused.cuh) rests on.  They say nothing about the
CUDA code itself (the GPU parity tests do); they pin the mathematics, exhaustively where the domain is small.

1. Integers below 2048 read as IEEE fp16 bit patterns are the subnormals and the first binade, all spaced 2^-24: fp16 subtraction of
   the patterns is exact integer subtraction and clamping at +0 is max(., 0) -- the kernel's `HADD2.SAT c, -|l - r|`
   (bm_calc_sad.v:449-457: c <- c - min(c, o)).
2. One row of the saturating column-sum chain, c <- min(max(c - o, 0) + n, 1023) (bm_calc_sad.v:449-466), is a clamp-add map
   c -> min(max(c + a, lo), hi); such maps are closed under composition, and a band of rows composes to (chain from 0, chain from 1023,
   sum of n - o).  The kernel cuts the chain of a handful of pairs into y-bands that way (MODE 1 / k_bm_chain / MODE 2)."""
import numpy as np


def _as_f16(u):
    return np.asarray(u, np.uint16).view(np.float16)


def _bits(f):
    return np.asarray(f, np.float16).view(np.uint16).astype(np.int64)


def test_fp16_patterns_below_2048_subtract_like_integers():
    c = np.arange(2048, dtype=np.uint16)[:, None]
    o = np.arange(256, dtype=np.uint16)[None, :]
    with np.errstate(all="raise"):
        d = _as_f16(c) - _as_f16(o)                                   # exact: both are multiples of 2^-24 below 2^-13
    sat = np.where(d < 0, np.float16(0), d)                           # .sat clamps a negative difference to +0 (the upper bound 1.0 is never reached)
    want = np.maximum(c.astype(np.int64) - o.astype(np.int64), 0)
    assert np.array_equal(_bits(sat), want)
    # |l - r| of two pixels as a pattern: the sign bit cleared
    l = np.arange(256, dtype=np.uint16)[:, None]
    r = np.arange(256, dtype=np.uint16)[None, :]
    m = np.abs(_as_f16(l) - _as_f16(r))
    assert np.array_equal(_bits(m), np.abs(l.astype(np.int64) - r.astype(np.int64)))
    # the largest pattern the kernel ever forms is 2046 (window 31 x 2 cap 33 in the cv::StereoBM variants): still on the uniform grid
    assert float(_as_f16(np.uint16(2047))) - float(_as_f16(np.uint16(2046))) == 2.0 ** -24


def _row(c, o, n):
    return np.minimum(np.maximum(c - o, 0) + n, 1023)


def test_saturating_chain_composes_to_three_numbers_per_band():
    rng = np.random.default_rng(7)
    lanes, rows, band = 4096, 96, 16
    # differences as the RTL sees them (6-bit pixels), with stretches of full contrast so that the chain saturates and recovers
    n = rng.integers(0, 64, (rows, lanes))
    n[20:50, : lanes // 2] = 63
    o = np.vstack([np.zeros((21, lanes), np.int64), n[:-21]])          # the row that leaves a 21-row window (nothing leaves while it fills)
    # reference: the sequential chain from 0
    c = np.zeros(lanes, np.int64)
    seq = []
    for y in range(rows):
        c = _row(c, o[y], n[y])
        seq.append(c.copy())
    assert (np.array(seq) == 1023).any() and (np.array(seq)[60:] < 1023).any()
    # band functions: chain from 0, chain from 1023, plain sum of n - o  ->  start states band after band  ->  bands from their states
    state = np.zeros(lanes, np.int64)
    for b0 in range(0, rows, band):
        lo, hi, a = np.zeros(lanes, np.int64), np.full(lanes, 1023, np.int64), np.zeros(lanes, np.int64)
        cur = state.copy()
        for y in range(b0, min(b0 + band, rows)):
            lo, hi, a = _row(lo, o[y], n[y]), _row(hi, o[y], n[y]), a + n[y] - o[y]
            cur = _row(cur, o[y], n[y])
            assert np.array_equal(cur, seq[y])                         # a band run from its exact start state reproduces the chain
        nxt = np.minimum(np.maximum(state + a, lo), hi)                # what k_bm_chain computes
        assert np.array_equal(nxt, cur)
        # ... and for EVERY possible start state, not only the one that occurred
        for s0 in (0, 1, 37, 511, 1000, 1023):
            t = np.full(lanes, s0, np.int64)
            for y in range(b0, min(b0 + band, rows)):
                t = _row(t, o[y], n[y])
            assert np.array_equal(t, np.minimum(np.maximum(s0 + a, lo), hi))
        state = nxt
        # the sum of n - o fits the kernel's signed 16-bit lanes by a wide margin
        assert np.abs(a).max() <= 63 * band

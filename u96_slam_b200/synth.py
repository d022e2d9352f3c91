"""Synthetic stereo pairs: deterministic, counter-based, numpy only.

splitmix64 is counter based (output i = mix(s0 + (i+1)*GAMMA)), so whole images are drawn with
vectorised uint64 arithmetic.  L0 is a textured strip of width W+2D (coarse random grid every
4 px, bilinearly upsampled with integer weights /16, plus per-pixel noise); the disparity field
is piecewise constant on 64x32 blocks; R[y][x] = L0[y][x+D+delta(y,x)] + small noise, so
R[y][x-delta] == L[y][x] inside a block and block edges give occlusion seams.
"""
import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _rand(s0, start, count):
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + 1 + count, dtype=np.uint64)
        return _mix(np.uint64(s0) + i * GAMMA)


def synth_pair(seed, frame, W, H, D, x_drift=0):
    """Returns (L, R) uint8 [H, W]."""
    with np.errstate(over="ignore"):
        s0 = np.uint64(seed) + GAMMA * np.uint64(frame)
    W0 = W + 2 * D + 8
    gw, gh = W0 // 4 + 2, H // 4 + 2
    grid = (_rand(s0, 0, gw * gh) & np.uint64(0xFF)).astype(np.int32).reshape(gh, gw)
    ys, xs = np.arange(H), np.arange(W0) + x_drift * frame
    gy, fy = ys // 4, (ys % 4)
    gx, fx = (xs // 4) % (gw - 1), (xs % 4)
    g00 = grid[np.ix_(gy, gx)]; g01 = grid[np.ix_(gy, gx + 1)]
    g10 = grid[np.ix_(gy + 1, gx)]; g11 = grid[np.ix_(gy + 1, gx + 1)]
    wy, wx = fy[:, None], fx[None, :]
    up = (g00 * (4 - wx) * (4 - wy) + g01 * wx * (4 - wy) + g10 * (4 - wx) * wy + g11 * wx * wy) // 16
    n0 = (_rand(s0, gw * gh, H * W0) & np.uint64(15)).astype(np.int32).reshape(H, W0) - 8
    L0 = np.clip(up + n0, 0, 255).astype(np.uint8)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    delta = 4 + (((xx >> 6) * 7 + (yy >> 5) * 13 + frame) % max(1, D - 12))
    L = L0[yy, xx + D]
    n1 = (_rand(s0, gw * gh + H * W0, H * W) & np.uint64(3)).astype(np.int32).reshape(H, W) - 1
    R = np.clip(L0[yy, xx + D + delta].astype(np.int32) + n1, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(L), R


def synth_batch(seed, frame0, n, W, H, D, x_drift=0):
    Ls, Rs = zip(*(synth_pair(seed, frame0 + i, W, H, D, x_drift) for i in range(n)))
    return np.stack(Ls), np.stack(Rs)


def identity_rect_params(W, H, f=None):
    """Near-identity rectification: identity rotation, f = f', c = (W/2, H/2)."""
    f = float(f if f is not None else W)
    fx = int(round(f * 65536))
    one = 1 << 24
    rot = [[one, 0, 0], [0, one, 0], [0, 0, one]]
    return dict(f=[[fx, fx], [fx, fx]], c=[W // 2, H // 2],
                f2inv=[int(round(2**32 / f)), int(round(2**32 / f))],
                c2_f2=[int(round((W / 2) / f * 2**24)), int(round((H / 2) / f * 2**24))],
                rot=[rot, rot])

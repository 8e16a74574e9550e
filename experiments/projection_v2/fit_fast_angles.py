#!/usr/bin/env python
"""Developer tool: coefficients and error bounds of the projection kernel's fast-path
angle approximations (csrc/project.cu: atan2_fast / asin_fast).

The fast path only has to land within the guard band of the exact float32 chain
(projection.py:54-85 with correctly rounded arctan2 / arcsin); points inside the band are
re-evaluated exactly.  This script fits the polynomials (Lawson-reweighted least squares,
near minimax), emulates the kernel's float32 operation sequence (fma = one rounding) on many
inputs and prints the largest |fast - exact chain| in pixels, which must stay below half the
guard bands tol_x / tol_y set in project.cu."""
import numpy as np
from numpy.polynomial import polynomial as P

f32 = np.float32


def fit(g, hi, n, k=4000, iters=300):
    x = np.cos(np.pi * (np.arange(k) + 0.5) / k)
    s = 1e-12 + (hi - 1e-12) * (x + 1) / 2
    t = np.sqrt(s)
    V = np.vander(s, n, increasing=True)
    y = g(t) / t
    wt = np.ones(k)
    for _ in range(iters):
        w = t * np.sqrt(wt)
        c = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
        e = np.abs((V @ c - y) * t)
        wt = wt * (e / e.max() + 1e-3)
        wt /= wt.sum()
    return c


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def horner(c, s):
    r = np.full_like(s, f32(c[-1]))
    for k in c[-2::-1]:
        r = fma(r, s, np.full_like(s, f32(k)))
    return r


def atan2_fast(y, x, c):
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    t = (mn * (f32(1) / mx)).astype(f32)          # MUFU.RCP (<= 1 ulp) then FMUL
    s = t * t
    a = t * horner(c, s)
    a = np.where(ay > ax, f32(1.5707963267948966) - a, a)
    a = np.where(x < 0, f32(3.141592653589793) - a, a)
    return np.where(y < 0, -a, a).astype(f32)


def asin_fast(q, c):
    s = q * q
    return (q * horner(c, s)).astype(f32)


def main():
    ca = fit(np.arctan, 1.0, 7)
    cs = fit(np.arcsin, 0.71 ** 2, 7)
    print("atan coeffs:", ", ".join("%.9ef" % v for v in ca.astype(f32)))
    print("asin coeffs:", ", ".join("%.9ef" % v for v in cs.astype(f32)))
    rng = np.random.default_rng(0)
    n = 4_000_000
    for W, H, up, down in ((2048, 64, 3.0, -25.0), (1024, 32, 10.0, -30.0), (1800, 40, 15.0, -25.0)):
        ang = rng.uniform(-np.pi, np.pi, n)
        r = np.exp(rng.uniform(np.log(0.5), np.log(120), n))
        pit = np.deg2rad(rng.uniform(down - 3, up + 3, n))
        x = (r * np.cos(pit) * np.cos(ang)).astype(f32)
        y = (r * np.cos(pit) * np.sin(ang)).astype(f32)
        z = (r * np.sin(pit)).astype(f32)
        depth = np.sqrt((x * x + y * y) + z * z)
        absl, fovh = f32(np.pi), f32(2 * np.pi)
        absd, fovv = f32(abs(np.deg2rad(down))), f32(abs(np.deg2rad(down)) + abs(np.deg2rad(up)))
        # exact chain (correctly rounded transcendental, then the reference's f32 ops)
        yaw = (-np.arctan2(y.astype(np.float64), x.astype(np.float64))).astype(f32)
        qx = z / depth
        pitch = np.arcsin(qx.astype(np.float64)).astype(f32)
        fx = ((yaw + absl) / fovh) * f32(W)
        fy = (f32(1) - (pitch + absd) / fovv) * f32(H)
        # fast chain as in the kernel
        sx, sy = f32(W) / fovh, f32(H) / fovv
        yf = -atan2_fast(y, x, ca)
        fxf = fma(yf, np.full(n, sx), np.full(n, absl * sx))
        qf = (z * (f32(1) / depth)).astype(f32)
        ok = np.abs(qf) <= f32(0.7)
        pf = asin_fast(qf, cs)
        fyf = fma(-pf, np.full(n, sy), np.full(n, f32(H) - absd * sy))
        ex = np.abs(fxf.astype(np.float64) - fx.astype(np.float64)).max() / W
        ey = (np.abs(fyf.astype(np.float64) - fy.astype(np.float64))[ok]).max() / H
        print("W=%d H=%d fov_v=%.3f: max|dfx|/W = %.3g  max|dfy|/H = %.3g  (dfy*fov_v/H = %.3g rad)"
              % (W, H, fovv, ex, ey, ey * fovv))


if __name__ == "__main__":
    main()

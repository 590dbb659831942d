#!/usr/bin/env python
"""Deterministic synthetic y4m generator (SURVEY.md section 8d recipe).

Luma canvas = smooth sinusoid field + 2x2-box-filtered noise; the view pans by
(2t + t%3, t) per frame; a 64x64 inverted-luma square moves by (17t, 9t); sensor
noise [-3,3]; hard scene cut (canvas inverted) every `cut` frames.  Output is the
exact y4m subset the reference parser accepts (util.c:184-307):
"YUV4MPEG2 W.. H.. F..:1 A1:1 Ip C420|C444" + "FRAME\n" per frame.

Used by tests/ and bench.py; no dependency on the reference.
"""
import argparse
import sys

import numpy as np


def _canvas(w, h, rng, noise_amp):
    cw, ch = w * 2 + 256, h * 2 + 256
    x = np.arange(cw, dtype=np.float64)[None, :]
    y = np.arange(ch, dtype=np.float64)[:, None]
    luma = 128 + 60 * np.sin(x / 37.0) * np.cos(y / 23.0) + 30 * np.sin((x + y) / 11.0)
    if noise_amp > 0:
        n = rng.uniform(-noise_amp, noise_amp, size=(ch + 1, cw + 1))
        n = (n[:-1, :-1] + n[1:, :-1] + n[:-1, 1:] + n[1:, 1:]) / 4.0
        luma = luma + n
    cb = 128 + 50 * np.sin(x / 53.0) + 0 * y
    cr = 128 + 50 * np.cos(y / 41.0) + 0 * x
    return luma, cb, cr


def _tri(v, period, amp):
    """integer triangle wave of v: values in [-amp, amp], exact on every platform"""
    p = np.mod(v, 2 * period)
    t = np.where(p < period, p, 2 * period - p)  # 0 .. period
    return (2 * amp * t) // period - amp


def _canvas_int(w, h, rng, noise_amp):
    """the same kind of content as _canvas from integer arithmetic only (no libm):
    used for the committed golden fixtures, whose input must be bit-identical on
    every machine"""
    cw, ch = w * 2 + 256, h * 2 + 256
    x = np.arange(cw, dtype=np.int64)[None, :]
    y = np.arange(ch, dtype=np.int64)[:, None]
    luma = 128 + (_tri(x, 58, 60) * _tri(y + 18, 36, 64)) // 64 + _tri(x + y, 17, 30)
    na = int(noise_amp)
    if na > 0:
        n = rng.integers(-na, na + 1, size=(ch + 1, cw + 1), dtype=np.int64)
        luma = luma + (n[:-1, :-1] + n[1:, :-1] + n[:-1, 1:] + n[1:, 1:] + 2 * na * 4) // 4 - 2 * na
    cb = 128 + _tri(x, 83, 50) + 0 * y
    cr = 128 + _tri(y + 32, 64, 50) + 0 * x
    return luma.astype(np.float64), cb.astype(np.float64), cr.astype(np.float64)


def frames(w, h, nfr, fmt="420", seed=1234, noise=24.0, sensor=3, cut=40, kind="sin"):
    """Yield (Y, U, V) uint8 arrays for nfr frames.  kind="tri": integer-only
    canvas (platform-independent bytes); every later step is exact in float64
    (sums of integers, /4 and /2 of integers, rint)."""
    if (w | h) & 1 and fmt in ("420", "422"):
        raise ValueError("odd luma dimensions are not valid input for the reference CLI (dsv_main.c:622) "
                         "and not supported by this generator")
    rng = np.random.default_rng(seed)
    luma, cb, cr = (_canvas_int if kind == "tri" else _canvas)(w, h, rng, noise)
    ch_h, ch_w = luma.shape
    sub = 2 if fmt == "420" else 1
    for t in range(nfr):
        ox = (2 * t + t % 3) % (ch_w - w)
        oy = t % (ch_h - h)
        Y = luma[oy:oy + h, ox:ox + w].copy()
        U = cb[oy:oy + h, ox:ox + w]
        V = cr[oy:oy + h, ox:ox + w]
        if cut > 0 and (t // cut) % 2 == 1:
            Y = 255.0 - Y
        sx = (17 * t) % max(1, w - 64)
        sy = (9 * t) % max(1, h - 64)
        Y[sy:sy + 64, sx:sx + 64] = 255.0 - Y[sy:sy + 64, sx:sx + 64]
        U = U.copy()
        V = V.copy()
        U[sy:sy + 64, sx:sx + 64] = 90.0
        V[sy:sy + 64, sx:sx + 64] = 170.0
        if sensor > 0:
            Y = Y + rng.integers(-sensor, sensor + 1, size=Y.shape)
        Y8 = np.clip(np.rint(Y), 0, 255).astype(np.uint8)
        if sub == 2:
            U = (U[0::2, 0::2] + U[1::2, 0::2] + U[0::2, 1::2] + U[1::2, 1::2]) / 4.0
            V = (V[0::2, 0::2] + V[1::2, 0::2] + V[0::2, 1::2] + V[1::2, 1::2]) / 4.0
        elif fmt == "422":
            U = (U[:, 0::2] + U[:, 1::2]) / 2.0
            V = (V[:, 0::2] + V[:, 1::2]) / 2.0
        elif fmt in ("411", "410"):
            cw_, ch_ = (w + 3) // 4, (h if fmt == "411" else (h + 3) // 4)
            U = U[::(1 if fmt == "411" else 4), ::4][:ch_, :cw_]
            V = V[::(1 if fmt == "411" else 4), ::4][:ch_, :cw_]
        U8 = np.clip(np.rint(U), 0, 255).astype(np.uint8)
        V8 = np.clip(np.rint(V), 0, 255).astype(np.uint8)
        yield Y8, U8, V8


def write_y4m(path, w, h, nfr, fmt="420", fps=30, **kw):
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F%d:1 A1:1 Ip C%s\n" % (w, h, fps, fmt)).encode())
        for Y, U, V in frames(w, h, nfr, fmt, **kw):
            f.write(b"FRAME\n")
            f.write(Y.tobytes())
            f.write(U.tobytes())
            f.write(V.tobytes())


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("-W", type=int, default=352)
    ap.add_argument("-H", type=int, default=288)
    ap.add_argument("-n", type=int, default=60)
    ap.add_argument("--fmt", default="420", choices=["420", "422", "444", "411", "410"])
    ap.add_argument("--fps", type=int, default=30)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--noise", type=float, default=24.0)
    ap.add_argument("--sensor", type=int, default=3)
    ap.add_argument("--cut", type=int, default=40)
    a = ap.parse_args(argv)
    write_y4m(a.out, a.W, a.H, a.n, a.fmt, a.fps, seed=a.seed, noise=a.noise,
              sensor=a.sensor, cut=a.cut)
    return 0


if __name__ == "__main__":
    sys.exit(main())

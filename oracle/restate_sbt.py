"""oracle/restate_sbt.py -- TEST INFRASTRUCTURE ONLY (never imported by the
product).  A numpy restatement of the reference's LOSSLESS subband transform
chain, the part of the pixel path whose defining property -- perfect
reconstruction -- can be checked at any size without the reference:

  dsv_fwd_sbt / dsv_inv_sbt in lossless mode (reference src/sbt.c:847-934):
    levels 1 .. lvls-2   reversible 5/3 lifting, rows then columns
                         (filterLOSSLESS / ifilterLOSSLESS :430-447, macros
                         DO_SIMPLE_HI :190-197, DO_SIMPLE_LO :199-203,
                         SCALE_PACK / UNSCALE_UNPACK :151-168, fwd_2d/inv_2d :449-473)
    levels lvls-1, lvls  plain 2x2 Haar without the overflow guard
                         (fwd :546-612, inv_simple :615-682)
    p2sbc / sbc2p        u8 <-> centred int32 (:798-831)

Pinned: tests/test_oracle.py checks every function here against the UNMODIFIED
reference (oracle/_ref/librefops.so -> dsv_fwd_sbt / dsv_inv_sbt) on random and
synthetic planes including odd sub-image sizes; the GPU tests then use it at
1920x1080 where it also serves as a second, reference-independent checker.
Everything else on the path is checked against the compiled reference directly
(oracle/_ref, see oracle/Makefile and oracle/refops.c).
"""
import numpy as np


def lb2(n):
    l, i = 0, 1
    while i < n:
        i <<= 1
        l += 1
    return l


def nlevels(w, h):
    """sbt.c:834-845"""
    return lb2(max(w, h))


def _rshift_up(x, s):
    return (x + (1 << s) - 1) >> s


def _tdiv(a, b):
    """C division (truncation toward zero) on int64 arrays"""
    q = np.abs(a) // b
    return np.where(a < 0, -q, q)


def lift_fwd_1d(v):
    """filterLOSSLESS on the LAST axis of v (int64, length n >= 2); returns packed [L | H]"""
    v = v.copy()
    n = v.shape[-1]
    even_n = n & ~1
    # DO_SIMPLE_HI: odd -= (l + r + 1) >> 1 ; last odd (n even) -= left
    idx = np.arange(1, n - 1, 2)
    if len(idx):
        v[..., idx] -= (v[..., idx - 1] + v[..., idx + 1] + 1) >> 1
    if not (n & 1):
        v[..., n - 1] -= v[..., n - 2]
    # DO_SIMPLE_LO: v[0] += v[1] >> 1 ; even i in [2, even_n) += (l + r + 2) >> 2
    v[..., 0] += v[..., 1] >> 1
    idx = np.arange(2, even_n, 2)
    if len(idx):
        v[..., idx] += (v[..., idx - 1] + v[..., idx + 1] + 2) >> 2
    # SCALE_PACK: lows to [0, ceil(n/2)), highs after
    h = (n + (n & 1)) // 2
    out = np.empty_like(v)
    out[..., :h] = v[..., 0::2]
    out[..., h:] = v[..., 1::2]
    return out


def lift_inv_1d(p):
    """ifilterLOSSLESS on the last axis of packed p"""
    n = p.shape[-1]
    even_n = n & ~1
    h = (n + (n & 1)) // 2
    v = np.empty_like(p)
    v[..., 0::2] = p[..., :h]
    v[..., 1::2] = p[..., h:]
    v[..., 0] -= v[..., 1] >> 1
    idx = np.arange(2, even_n, 2)
    if len(idx):
        v[..., idx] -= (v[..., idx - 1] + v[..., idx + 1] + 2) >> 2
    idx = np.arange(1, n - 1, 2)
    if len(idx):
        v[..., idx] += (v[..., idx - 1] + v[..., idx + 1] + 1) >> 1
    if not (n & 1):
        v[..., n - 1] += v[..., n - 2]
    return v


def haar_fwd(sub):
    """fwd() without overflow guard on an sh x sw sub-image; returns packed quadrants"""
    sh, sw = sub.shape
    ch, cw = (sh + 1) // 2, (sw + 1) // 2
    out = np.zeros_like(sub)
    eh, ew = sh & ~1, sw & ~1
    x0, x1 = sub[0:eh:2, 0:ew:2], sub[0:eh:2, 1:ew:2]
    x2, x3 = sub[1:eh:2, 0:ew:2], sub[1:eh:2, 1:ew:2]
    out[:eh // 2, :ew // 2] = x0 + x1 + x2 + x3
    out[:eh // 2, cw:cw + ew // 2] = x0 - x1 + x2 - x3
    out[ch:ch + eh // 2, :ew // 2] = x0 + x1 - x2 - x3
    out[ch:ch + eh // 2, cw:cw + ew // 2] = x0 - x1 - x2 + x3
    if sw & 1:
        a, b = sub[0:eh:2, sw - 1], sub[1:eh:2, sw - 1]
        out[:eh // 2, cw - 1] = 2 * (a + b)
        out[ch:ch + eh // 2, cw - 1] = 2 * (a - b)
    if sh & 1:
        a, b = sub[sh - 1, 0:ew:2], sub[sh - 1, 1:ew:2]
        out[ch - 1, :ew // 2] = 2 * (a + b)
        out[ch - 1, cw:cw + ew // 2] = 2 * (a - b)
        if sw & 1:
            out[ch - 1, cw - 1] = sub[sh - 1, sw - 1] * 4
    return out


def haar_inv(p):
    """inv_simple() without overflow guard"""
    sh, sw = p.shape
    ch, cw = (sh + 1) // 2, (sw + 1) // 2
    eh, ew = sh & ~1, sw & ~1
    out = np.zeros_like(p)
    LL, LH = p[:eh // 2, :ew // 2], p[:eh // 2, cw:cw + ew // 2]
    HL, HH = p[ch:ch + eh // 2, :ew // 2], p[ch:ch + eh // 2, cw:cw + ew // 2]
    out[0:eh:2, 0:ew:2] = _tdiv(LL + LH + HL + HH, 4)
    out[0:eh:2, 1:ew:2] = _tdiv(LL - LH + HL - HH, 4)
    out[1:eh:2, 0:ew:2] = _tdiv(LL + LH - HL - HH, 4)
    out[1:eh:2, 1:ew:2] = _tdiv(LL - LH - HL + HH, 4)
    if sw & 1:
        LL, HL = p[:eh // 2, cw - 1], p[ch:ch + eh // 2, cw - 1]
        out[0:eh:2, sw - 1] = _tdiv(LL + HL, 4)
        out[1:eh:2, sw - 1] = _tdiv(LL - HL, 4)
    if sh & 1:
        LL, LH = p[ch - 1, :ew // 2], p[ch - 1, cw:cw + ew // 2]
        out[sh - 1, 0:ew:2] = _tdiv(LL + LH, 4)
        out[sh - 1, 1:ew:2] = _tdiv(LL - LH, 4)
        if sw & 1:
            out[sh - 1, sw - 1] = _tdiv(p[ch - 1, cw - 1], 4)
    return out


def fwd_sbt_lossless(plane_u8, cw=None, ch=None):
    """dsv_fwd_sbt with params->lossless on one u8 plane (h x w).  cw/ch: the
    coefficient plane size (chroma planes are rounded up to even dimensions,
    frame.c:41-42; the extra row/column stays zero like the reference's calloc)."""
    h, w = plane_u8.shape
    cw, ch = cw or w, ch or h
    c = np.zeros((ch, cw), np.int64)
    c[:h, :w] = plane_u8.astype(np.int64) - 128
    lvls = nlevels(cw, ch)
    for l in range(1, lvls + 1):
        sw, sh = _rshift_up(cw, l - 1), _rshift_up(ch, l - 1)
        sub = c[:sh, :sw]
        if 1 <= l <= lvls - 2:
            t = lift_fwd_1d(sub)                       # rows
            sub = lift_fwd_1d(t.T.copy()).T            # columns
        else:
            sub = haar_fwd(sub)
        c[:sh, :sw] = sub
    return c.astype(np.int32)


def inv_sbt_lossless(coefs, w=None, h=None):
    """dsv_inv_sbt with params->lossless; returns the u8 plane (h x w)"""
    ch, cw = coefs.shape
    w, h = w or cw, h or ch
    c = coefs.astype(np.int64).copy()
    lvls = nlevels(cw, ch)
    for l in range(lvls, 0, -1):
        sw, sh = _rshift_up(cw, l - 1), _rshift_up(ch, l - 1)
        sub = c[:sh, :sw]
        if 1 <= l <= lvls - 2:
            t = lift_inv_1d(sub.T.copy()).T            # columns first (inv_2d)
            sub = lift_inv_1d(t)                       # then rows
        else:
            sub = haar_inv(sub)
        c[:sh, :sw] = sub
    return np.clip(c[:h, :w] + 128, 0, 255).astype(np.uint8)

"""The numpy restatement in oracle/restate_sbt.py (lossless transform chain):
pinned against the unmodified reference, then used as a reference-independent
checker of the CUDA transform at full size (perfect reconstruction)."""
import os
import sys

import numpy as np
import pytest

import ops
import util

sys.path.insert(0, os.path.join(util.ROOT, "oracle"))
import restate_sbt as R  # noqa: E402

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def _planes(w, h, fmt, seed):
    rng = np.random.default_rng(seed)
    cw, ch = (w, h) if fmt == "444" else ((w + 1) // 2, (h + 1) // 2)
    y = rng.integers(0, 256, (h, w)).astype(np.uint8)
    # smooth-ish content exercises large and small coefficients alike
    y[: h // 2] = (np.add.outer(np.arange(h // 2), np.arange(w)) % 256).astype(np.uint8)
    u = rng.integers(0, 256, (ch, cw)).astype(np.uint8)
    v = rng.integers(0, 256, (ch, cw)).astype(np.uint8)
    return [y, u, v]


@need_ref
@pytest.mark.parametrize("w,h,fmt", [(352, 288, "420"), (200, 136, "420"), (176, 144, "444"), (270, 70, "444")])
def test_restatement_matches_reference(w, h, fmt):
    cfg = ops.Cfg(w, h, 0x5 if fmt == "420" else 0x0, isP=0, lossless=1)
    ref = ops.Ref()
    pl = _planes(w, h, fmt, 3)
    yuv = b"".join(p.tobytes() for p in pl)
    bd = np.zeros(cfg.nblk, np.uint8)
    for p in range(3):
        cw, ch = ref.coef_dims(cfg, p)
        want = ref.fwd_sbt(cfg, p, yuv, bd)
        got = R.fwd_sbt_lossless(pl[p], cw, ch)
        assert np.array_equal(got, want), "forward, plane %d" % p
        back = R.inv_sbt_lossless(got, pl[p].shape[1], pl[p].shape[0])
        assert np.array_equal(back, pl[p]), "restatement is not perfectly reconstructing, plane %d" % p
        rb = np.frombuffer(ref.inv_sbt(cfg, p, 1, want, bd), np.uint8).reshape(pl[p].shape)
        assert np.array_equal(rb, pl[p])


def _device_roundtrip(w, h, fmt, emu):
    cfg = ops.Cfg(w, h, 0x5 if fmt == "420" else 0x0, isP=0, lossless=1)
    pl = _planes(w, h, fmt, 11)
    yuv = b"".join(p.tobytes() for p in pl)
    bd = np.zeros(cfg.nblk, np.uint8)
    D = ops.Dev(cfg, emu)
    try:
        for p in range(3):
            got = D.fwd_sbt(p, yuv, bd)
            want = R.fwd_sbt_lossless(pl[p], got.shape[1], got.shape[0])
            assert np.array_equal(got, want), "plane %d forward transform differs from the restatement" % p
    finally:
        D.close()


def test_device_forward_lossless_vs_restatement_emulated():
    util.ensure_emu()
    _device_roundtrip(352, 288, "420", True)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,fmt", [(1920, 1080, "444"), (1920, 1080, "420"), (1280, 720, "420")])
def test_device_forward_lossless_vs_restatement_gpu(w, h, fmt):
    _device_roundtrip(w, h, fmt, False)

#!/usr/bin/env python
"""Regenerates tests/golden/golden.json from the UNMODIFIED reference
(oracle/_ref/dsv2, built from /root/reference by oracle/Makefile).

The reference ships no golden vectors, so the fixtures are
outputs of the reference itself on the deterministic synthetic clips of
tools/synth_y4m.py (kind="tri": integer arithmetic only, so the input bytes are
the same on every machine and the test can insist on them): md5 of the input y4m, of the .dsv the reference encoder
writes, and of the y4m the reference decoder writes from it.  Run in a container
that has /root/reference:   python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402

CASES = [
    # name, w, h, frames, fmt, fps, reference CLI args
    ("cif", 352, 288, 12, "420", 30, ["-qp=60", "-gop=48"]),
    ("cif_lowq", 352, 288, 8, "420", 30, ["-qp=20", "-gop=4"]),
    ("odd", 200, 136, 6, "420", 30, ["-qp=70", "-gop=3"]),
    ("cif444", 352, 288, 6, "444", 30, ["-qp=50", "-gop=5"]),
    ("cif444ll", 352, 288, 4, "444", 30, ["-qp=100"]),
    ("cif_cqp", 352, 288, 6, "420", 30, ["-qp=45", "-rc_mode=2", "-effort=5"]),
    ("oddchroma", 354, 290, 5, "420", 30, ["-qp=60", "-gop=48"]),
]


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def main():
    assert util.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, w, h, n, fmt, fps, args in CASES:
        y4m = util.clip("golden_" + name, w, h, n, fmt, fps=fps, kind="tri")
        dsv = util.ref_encode(y4m, args, "golden")
        dec = util.ref_decode(dsv)
        out[name] = {"w": w, "h": h, "frames": n, "fmt": fmt, "fps": fps, "args": args,
                     "y4m_md5": md5(y4m), "dsv_md5": md5(dsv), "dsv_bytes": os.path.getsize(dsv),
                     "decoded_y4m_md5": md5(dec), "decoded_frames_md5": util.frames_md5(util.read_y4m(dec)[2])}
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "fixtures")


if __name__ == "__main__":
    main()

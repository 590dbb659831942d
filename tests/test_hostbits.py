"""Host bit I/O: the fast plane writer / reader against the general one."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "digital-subband-video-2_b200", "host")


def test_plane_pairs_fast_path_equals_general_path(tmp_path):
    exe = str(tmp_path / "hzcc_fuzz")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-o", exe,
                    os.path.join(ROOT, "tests", "hzcc_fuzz.c"), os.path.join(HOST, "dsv_hzcc.c"),
                    os.path.join(HOST, "dsv_bits.c"), os.path.join(HOST, "dsv_core.c")], check=True)
    r = subprocess.run([exe, "240"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600)
    assert r.returncode == 0 and "all equal" in r.stdout, r.stdout[-400:]

"""CPU: the tile loader's eight-nucleotides-per-lane code (rnamotif_b200/csrc/gm_tilebits.h: expanded bytes,
reverse-complement bytes and base-bitset bits from a word of eight packed codes) against the per-nucleotide
definitions (expand_code / complement_byte = rm_b2bc and mk_rcmp, src/rnamot.c:200-208), compiled for the host:
every 16-bit half in four arrangements plus two million random words."""
import os
import subprocess

import helpers


def test_expand8_matches_per_nucleotide_definitions(tmp_path):
    exe = str(tmp_path / "tilebits_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(helpers.ROOT, "rnamotif_b200", "csrc"),
                    os.path.join(helpers.HERE, "csrc", "tilebits_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "ok", out


def test_lane_window_build_matches_per_nucleotide_loop(tmp_path):
    exe = str(tmp_path / "winbuild_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(helpers.ROOT, "rnamotif_b200", "csrc"),
                    os.path.join(helpers.HERE, "csrc", "winbuild_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "ok", out

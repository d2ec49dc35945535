"""CPU: the C-ABI library loads, exports every symbol include/gpumotif.h
declares, validates plans on the host, and refuses to run without a GPU (no
CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from rnamotif_b200 import gpumotif
import helpers


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(helpers.ROOT, "include", "gpumotif.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(gpumotif.EXPORTS), declared ^ set(gpumotif.EXPORTS)
    L = C.CDLL(gpumotif.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym


def test_plan_size_matches_header():
    hdr = os.path.join(helpers.ROOT, "include")
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "gpumotif_plan.h"\n'
                             'int main(){printf("%zu %zu %zu\\n", sizeof(gm_plan_t), sizeof(gm_hit_hdr_t), sizeof(gm_hit_el_t));return 0;}\n')
        exe = os.path.join(tmp, "s")
        subprocess.run(["gcc", "-I", hdr, src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    assert int(out[0]) == gpumotif.PLAN_BYTES
    assert int(out[1]) == 32 and int(out[2]) == 8


@pytest.mark.parametrize("name", helpers.golden_names())
def test_plan_check_accepts_golden_plans(name):
    plan = helpers.load_plan(name)
    why = gpumotif.plan_check(plan)
    # every shipped descriptor that the reference can run must be accepted
    assert why is None, why


def test_plan_check_rejects_garbage():
    assert gpumotif.plan_check(b"\0" * gpumotif.PLAN_BYTES) is not None
    plan = bytearray(helpers.load_plan("trna"))
    plan[8:12] = (0).to_bytes(4, "little")  # n_descr = 0
    assert gpumotif.plan_check(bytes(plan)) is not None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gpumotif.GpuMotifError, match="no CUDA device"):
        gpumotif.MotifSearch(helpers.load_plan("trna"))


def test_public_headers_are_plain_c():
    """include/*.h is the drop-in boundary: C99, no C++ or torch types."""
    import subprocess, tempfile
    inc = os.path.join(helpers.ROOT, "include")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "h.c")
        open(src, "w").write('#include "gpumotif.h"\n#include "gpumotif_plan.h"\nint main(void){return 0;}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", src], check=True)


def test_level0_organisation_chosen_per_plan():
    """gm_plan_describe (host side): the two-stage sieve for plans whose look-ahead bitsets (trna), literal (ire,
    pk1) or chain (descr.quad, qu+tr) alone leave few starts; the single-stage sieve where the first helix itself is
    the filter (score.1)."""
    def level0(name):
        return gpumotif.plan_describe(helpers.load_plan(name)).splitlines()[1]
    for name in ("trna", "ire", "pk1", "descr.quad", "qu+tr"):
        assert "sieve 1 two-stage 1" in level0(name), (name, level0(name))
    assert "sieve 1 two-stage 0" in level0("score.1"), level0("score.1")

"""CPU: libtomoenc.so builds for sm_100a, loads, and exports every symbol include/tomoenc.h declares
(no compute calls -- there is no GPU here)."""
import ctypes

import pytest
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tomoenc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(te_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_the_header(built_lib):
    declared = _declared_symbols()
    assert len(declared) >= 20
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, "declared in tomoenc.h but not exported: %s" % missing
    assert sorted(built_lib.EXPORTED_SYMBOLS) == declared, "ctypes signature table out of sync with tomoenc.h"
    L = built_lib.lib()
    assert L.te_version() >= 100
    assert L.te_launch_count() == 0


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    L = built_lib.lib()
    rc = L.te_gather(None, 3, built_lib.i64x3((1, 1, 1)), None, None, 1, built_lib.i32x3((1, 1, 1)), None, 0, None)
    assert rc == -1 and b"elem_size" in L.te_last_error()
    rc = L.te_gather(None, 1, built_lib.i64x3((1, 1, 1)), None, None, 0, built_lib.i32x3((1, 1, 1)), None, 0, None)
    assert rc == 0        # empty input is a no-op


@pytest.mark.gpu
def test_sass_contains_blackwell_instructions_on_the_gpu_box(built_lib):
    """The same SASS check on the library the GPU box actually loads (the driver runs -m gpu there)."""
    _check_sass(built_lib)


def test_sass_contains_blackwell_instructions(built_lib):
    """The conv kernels must be tcgen05 / TMA code, not recompiled legacy MMA."""
    _check_sass(built_lib)


def _check_sass(built_lib):
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "no tcgen05.mma in the library"
    assert "UTMALDG" in sass, "no TMA tensor loads in the library"
    # legacy mma.sync is tolerated in exactly one place: wgrad_mma_kernel, the 1-tap product of the transposed-conv
    # backward and the fallback for odd channel counts; every forward / data-gradient conv and
    # the 3x3x3 weight gradient (wgrad_umma_kernel) are tcgen05
    legacy = set()
    for block in sass.split("Function : ")[1:]:
        name = block.split("\n", 1)[0]
        if "HMMA." in block.replace("UTCHMMA", ""):
            legacy.add(name)
    assert all("wgrad_mma_kernel" in n for n in legacy), "legacy mma.sync found in %s" % sorted(legacy)
    assert not any("conv_umma_kernel" in n or "wgrad_umma_kernel" in n for n in legacy)
    umma = [b.split("\n", 1)[0] for b in sass.split("Function : ")[1:] if "UTCHMMA" in b]
    assert any("wgrad_umma_kernel" in n for n in umma), "the weight-gradient kernel must be tcgen05 code"

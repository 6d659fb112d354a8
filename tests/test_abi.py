"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/iqo_cuda.h declares, reports errors without a GPU (no CPU fallback), and its host
planner produces the reference's tables and index maps.  No kernel is launched here."""
import os
import re
import subprocess

import numpy as np
import pytest

import libiqo_b200 as iqo
from oracle_lib import AREA, LANCZOS, LINEAR, oracle_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "iqo_cuda.h")).read()
    return sorted(set(re.findall(r"IQO_CUDA_API\s+[\w\s\*]+?\b(iqo_cuda_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == iqo.exported_symbols()


def test_library_exports_every_declared_symbol():
    lib = iqo.lib()
    out = subprocess.check_output(["nm", "-D", "--defined-only", iqo.LIB_PATH]).decode()
    exported = set(re.findall(r"\bT (iqo_cuda_\w+)", out))
    for name in _declared_symbols():
        assert name in exported, name
        assert getattr(lib, name) is not None
    # and the C++ classes with the reference's mangled signatures (iqo::LanczosResizer etc.)
    cxx = subprocess.check_output(["nm", "-DC", "--defined-only", iqo.LIB_PATH]).decode()
    for sig in ("iqo::LanczosResizer::LanczosResizer(unsigned int, unsigned long, unsigned long, unsigned long, unsigned long, unsigned long)",
                "iqo::LanczosResizer::resize(unsigned long, unsigned char const*, unsigned long, unsigned char*)",
                "iqo::AreaResizer::AreaResizer(unsigned long, unsigned long, unsigned long, unsigned long)",
                "iqo::AreaResizer::resize(unsigned long, unsigned char const*, unsigned long, unsigned char*)",
                "iqo::LinearResizer::LinearResizer(unsigned long, unsigned long, unsigned long, unsigned long)",
                "iqo::LinearResizer::resize(unsigned long, unsigned char const*, unsigned long, unsigned char*)",
                "iqo::LanczosResizer::~LanczosResizer()"):
        assert sig in cxx, sig


def test_version_and_error_strings():
    lib = iqo.lib()
    assert b"sm_100a" in lib.iqo_cuda_version()
    assert isinstance(lib.iqo_cuda_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    if iqo.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(iqo.IqoCudaError) as e:
        iqo.LanczosResizer(3, 64, 48, 32, 24)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_argument_errors_are_reported_before_touching_cuda():
    for args, code in (((LANCZOS, 3, 0, 10, 5, 5, 1), -1), ((LANCZOS, 0, 10, 10, 5, 5, 1), -1),
                       ((LANCZOS, 3, 8, 8, 5, 5, 1), -2),      # source shorter than the kernel
                       ((LANCZOS, 1, 100, 100, 99, 99, 1), -3),  # border denominator 0
                       ((AREA, 0, 1 << 31, 10, 5, 5, 1), -4)):
        with pytest.raises(iqo.IqoCudaError) as e:
            iqo.plan_query(*args, 0)
        assert e.value.code == code, args


CASES = [
    (LANCZOS, 3, 1920, 1080, 1280, 720, 1), (LANCZOS, 3, 1920, 1080, 960, 540, 1),
    (LANCZOS, 2, 3840, 2160, 1920, 1080, 1), (LANCZOS, 2, 1920, 1080, 960, 540, 2),
    (LANCZOS, 4, 32768, 32768, 12000, 12000, 1), (LANCZOS, 3, 40, 30, 64, 48, 1),
    (LANCZOS, 5, 641, 479, 333, 211, 1), (LANCZOS, 2, 333, 211, 641, 479, 1),
    (LANCZOS, 9, 1000, 900, 77, 50, 1), (LANCZOS, 4, 100, 80, 50, 40, 2),
    (AREA, 0, 3840, 2160, 1920, 1080, 1), (AREA, 0, 1920, 1080, 1280, 720, 1),
    (AREA, 0, 641, 479, 333, 211, 1), (AREA, 0, 40, 30, 64, 48, 1),
    (LINEAR, 0, 1280, 720, 3840, 2160, 1), (LINEAR, 0, 40, 30, 100, 75, 1), (LINEAR, 0, 333, 211, 641, 479, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_planner_tables_match_oracle(case):
    kind, deg, sw, sh, dw, dh, px = case
    for axis, S, D in ((0, sw, dw), (1, sh, dh)):
        q = iqo.plan_query(kind, deg, sw, sh, dw, dh, px, axis)
        ref = oracle_table(kind, axis, S, D, deg, px)
        assert q["numCoefs"] == ref.shape[1] and q["numTables"] == ref.shape[0]
        assert np.array_equal(q["coefs"][: ref.shape[0]], ref)


def test_planner_index_maps():
    # Lanczos3 1080 -> 540: first = 2d - 5, border rows 3 + 3
    q = iqo.plan_query(LANCZOS, 3, 1920, 1080, 960, 540, 1, 1)
    assert (q["mainBegin"], q["mainEnd"]) == (3, 537)
    assert np.array_equal(q["first"], 2 * np.arange(540) - 5)
    assert (q["row"][3:537] == 0).all() and (q["row"][:3] > 0).all() and (q["row"][537:] > 0).all()
    assert q["numRows"] == 1 + 6
    # border rows: out-of-range taps removed
    assert q["coefs"][q["row"][0]].tolist() == [0, 0, 0, 0, 0, 28, 28, 9, -4, -2, 1, 0]
    # Area: first = floor(d*S/D); Linear 3x: first = floor(((2d+1)S - D)/(2D)) with replicated ends
    q = iqo.plan_query(AREA, 0, 1920, 1080, 1280, 720, 1, 0)
    assert np.array_equal(q["first"], np.arange(1280) * 1920 // 1280)
    q = iqo.plan_query(LINEAR, 0, 1280, 720, 3840, 2160, 1, 0)
    d = np.arange(3840)
    f = ((2 * d + 1) * 1280 - 3840) // (2 * 3840)
    f[0], f[-1] = 0, 1279
    assert np.array_equal(q["first"], f)
    assert q["coefs"][q["row"][0]].tolist() == [32768, 0] and q["coefs"][q["row"][-1]].tolist() == [32768, 0]
    # identity axis is a single tap of weight one
    q = iqo.plan_query(LANCZOS, 3, 64, 48, 64, 30, 1, 0)
    assert q["numCoefs"] == 1 and q["coefs"].tolist() == [[16384]] and np.array_equal(q["first"], np.arange(64))


def test_kernel_family_of_the_baseline_configs():
    """Host-only planner view: the kernel family each BASELINE config maps to (large, aligned launches)."""
    L, A, Li = iqo.LANCZOS, iqo.AREA, iqo.LINEAR
    expect = [
        ((L, 3, 1920, 1080, 1280, 720, 1), "ratio_stream"),      # cfg1
        ((A, 0, 3840, 2160, 1920, 1080, 1), "area2"),            # cfg2a
        ((Li, 0, 1280, 720, 3840, 2160, 1), "linear_up3"),       # cfg2b
        ((L, 2, 3840, 2160, 1920, 1080, 1), "half_sym"),         # cfg3 luma
        ((L, 2, 1920, 1080, 960, 540, 2), "half_small"),         # cfg3 chroma
        ((L, 3, 1920, 1080, 960, 540, 1), "half_sym"),           # cfg4
        ((L, 4, 32768, 32768, 12000, 12000, 1), "lanczos_stream"),  # cfg5
        ((L, 3, 1000, 700, 333, 500, 1), "lanczos_stream"),      # arbitrary Lanczos ratio
        ((A, 0, 1920, 1080, 1280, 720, 1), "area_down"),         # Area at 3:2
        ((A, 0, 1000, 700, 700, 400, 1), "packed"),              # Area at 10:7
        ((L, 3, 1920, 1080, 960, 720, 1), "ratio_stream"),       # 2:1 on X only
        ((A, 0, 3840, 2160, 1280, 720, 1), "area_down"),         # Area at 3:1
        ((A, 0, 1920, 1080, 1280, 540, 1), "area_down"),         # Area at 3:2 on X, 2:1 on Y
        ((A, 0, 1910, 1080, 955, 720, 1), "packed"),             # Area 2:1 on X, width not a multiple of 8
        ((Li, 0, 1280, 720, 1920, 1080, 1), "linear_up_2_3"),    # Linear at 2:3
        ((Li, 0, 1440, 810, 1920, 1080, 1), "linear_up_3_4"),
        ((Li, 0, 1920, 1080, 1280, 720, 1), "linear_3_2"),       # mild reductions share the Linear kernel
        ((Li, 0, 1920, 1080, 1440, 810, 1), "linear_4_3"),
        ((Li, 0, 1920, 1080, 960, 540, 1), "packed"),            # from 2:1 on the reference's iterator starts elsewhere
        ((Li, 0, 1284, 720, 1926, 1080, 1), "packed"),           # 2:3, width not a multiple of 8
        ((L, 5, 614, 411, 401, 342, 1), "generic"),              # a border row may wrap int16 after its division
    ]
    for args, name in expect:
        assert iqo.plan_kernel(*args)[0] == name, args


def test_cxx98_user_program_builds_against_the_public_headers(tmp_path):
    """A program written the way libiqo's users write theirs (reference: sample/resize_yuv420p.cpp:125-162,
    `#include <libiqo/iqo.hpp>`, C++98, default pxScale) compiles and links against include/libiqo + the library.
    When the reference tree is present (the build container) its own sample is compiled unchanged as well."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = tmp_path / "user.cpp"
    src.write_text(
        "#include <vector>\n#include <libiqo/iqo.hpp>\n"
        "int main(int argc, char **) {\n"
        "    if (argc < 100) return 0;  // link check only: never constructs a resizer on a box without a GPU\n"
        "    std::vector<unsigned char> s(64 * 48), d(32 * 24);\n"
        "    iqo::LanczosResizer l(3, 64, 48, 32, 24), c(2, 64, 48, 32, 24, 2);\n"
        "    iqo::AreaResizer a(64, 48, 32, 24);\n"
        "    iqo::LinearResizer u(32, 24, 64, 48);\n"
        "    l.resize(64, &s[0], 32, &d[0]); c.resize(64, &s[0], 32, &d[0]);\n"
        "    a.resize(64, &s[0], 32, &d[0]); u.resize(32, &d[0], 64, &s[0]);\n"
        "    return 0;\n}\n")
    exe = tmp_path / "user"
    libdir = os.path.dirname(iqo.LIB_PATH)
    subprocess.check_call([cxx, "-std=c++98", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-liqo_cuda", "-Wl,-rpath," + libdir])
    assert subprocess.call([str(exe)]) == 0
    sample = "/root/reference/sample/resize_yuv420p.cpp"
    if os.path.exists(sample):
        subprocess.check_call([cxx, "-std=c++98", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), sample])


def test_narrow_source_plans_are_accepted_on_x_only():
    """mainBegin > mainEnd: defined on X while mainBegin <= dstW (every column a border column), rejected on Y."""
    q = iqo.plan_query(0, 3, 10, 20, 3, 20, 1, 0)
    assert q["mainBegin"] == 0 and q["mainEnd"] == 0
    assert (q["row"] >= q["numTables"]).all()          # planner-made border rows only
    assert iqo.plan_kernel(0, 3, 10, 20, 3, 20, 1)[0] == "generic"
    with pytest.raises(iqo.IqoCudaError) as e:
        iqo.plan_query(0, 3, 7, 30, 2, 30, 1, 0)       # mainBegin (3) > dstW (2): the reference writes past the row
    assert e.value.code == -2
    with pytest.raises(iqo.IqoCudaError) as e:
        iqo.plan_query(0, 3, 30, 10, 30, 3, 1, 1)
    assert e.value.code == -2

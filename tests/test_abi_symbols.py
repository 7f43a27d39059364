"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/dcmrta.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "dcmrta.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcm_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from dcmrta_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dcmrta.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert b"sm_100a" in L.dcm_version()


def test_binary_targets_sm100a_only():
    import subprocess
    from dcmrta_b200 import library_path
    out = subprocess.run(["cuobjdump", "-lelf", str(library_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dcmrta_b200 import _lib
    h = C.c_void_p()
    rc = _lib.lib().dcm_create(C.byref(h), 0, 4, 20, 50, 5, 0)
    assert rc == -3 and not h.value                      # DCM_ERR_DEVICE
    assert b"no CPU fallback" in _lib.lib().dcm_last_error()
    with pytest.raises(_lib.DcmError):
        from dcmrta_b200 import BatchedTaskEnv
        BatchedTaskEnv(4, 20, 50)


def test_shape_limits_are_rejected_before_touching_the_device():
    from dcmrta_b200 import _lib
    h = C.c_void_p()
    for (B, A, T, M) in ((0, 20, 50, 5), (4, 65, 50, 5), (4, 20, 255, 5), (4, 20, 50, 17)):
        assert _lib.lib().dcm_create(C.byref(h), 0, B, A, T, M, 0) == -2


def test_layout_matches_design_numbers():
    """record sizes quoted in DESIGN.md for 20A/50T/M5"""
    import subprocess, sys
    src = r'''
#include <cstdio>
#include "dcmrta_b200/csrc/dcm_layout.h"
int main(){ DcmLayout L = dcm_make_layout(20,50,5); printf("%d %d %d %zu\n", L.dyn_bytes, L.sta_bytes, L.stage_bytes, sizeof(DcmHdr)); }
'''
    exe = ROOT / "build" / "layout_probe"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-x", "c++", "-", "-I", str(ROOT), "-o", str(exe)], input=src, text=True, check=True)
    dyn, sta, stage, hdr = map(int, subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    assert hdr == 48 and dyn % 16 == 0 and sta % 16 == 0
    assert (dyn, sta, stage) == (3520, 1280, 1552)


def test_sass_shows_the_hardware_paths_the_design_names():
    """DESIGN.md 4 / 9 by mnemonic (B200_PROFILING.md): the observation tile kernel moves its spans with TMA bulk copies against an mbarrier
    (UBLKCP.S.G in, UBLKCP.G.S out, SYNCS.*.TRANS64), k_step stages its gathers with cp.async (LDGSTS); the policy's attention runs on the
    tensor cores through the legacy path (HMMA.16816.F32.BF16: a 51 x 51 x 16 problem is below tcgen05's smallest tile).  No tcgen05 / UTCMMA
    anywhere: the step is not a contraction and the policy's dense GEMMs are cuBLASLt's."""
    import subprocess
    from dcmrta_b200 import build, library_path
    build.build_policy()
    env_sass = subprocess.run(["cuobjdump", "-sass", str(library_path())], capture_output=True, text=True).stdout
    for m in ("UBLKCP.S.G", "UBLKCP.G.S", "SYNCS.ARRIVE.TRANS64", "LDGSTS.E.64"):
        assert m in env_sass, m
    assert "HMMA" not in env_sass and "UTCMMA" not in env_sass and "UTCHMMA" not in env_sass
    pol_sass = subprocess.run(["cuobjdump", "-sass", str(build.POLICY_SO)], capture_output=True, text=True).stdout
    assert pol_sass.count("HMMA.16816.F32.BF16") >= 16 and "MUFU.EX2" in pol_sass
    assert "UTCMMA" not in pol_sass and "UTCHMMA" not in pol_sass

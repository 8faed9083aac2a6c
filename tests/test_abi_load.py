"""CPU suite: the C-ABI library loads, exports every symbol include/ssm_b200.h declares and rejects
bad arguments without a GPU; the Python mirror refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import ssm_b200
from ssm_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ssm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 12
    L = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libssm_b200.so does not export %s" % n
    assert sorted(_abi.EXPORTS) == names


def test_version_and_error_string():
    assert ssm_b200.abi_version() == 8
    assert isinstance(_abi.lib().ssm_last_error(), bytes)


def test_argument_errors_without_gpu():
    L = _abi.lib()
    t = _abi.SsmTensor(0, 0, 0, 0)
    # NULL data pointer
    rc = L.ssm_warp_fwd(ctypes.byref(t), ctypes.byref(t), ctypes.byref(t), 1, 3, 8, 8, 0, 0, None)
    assert rc == -1 and b"NULL" in L.ssm_last_error()
    # bad shape
    rc = L.ssm_warp_fwd(ctypes.byref(t), ctypes.byref(t), ctypes.byref(t), 0, 3, 8, 8, 0, 0, None)
    assert rc == -2
    # bad dtype / coord mode
    assert L.ssm_fuse_fwd(None, None, None, None, None, None, 1, 1, 8, 8, 7, 0, None) == -3
    assert L.ssm_fuse_fwd(None, None, None, None, None, None, 1, 1, 8, 8, 0, 5, None) == -3
    # misaligned pointer
    bad = _abi.SsmTensor(3, 64, 0, 64)
    assert L.ssm_warp_fwd(ctypes.byref(bad), ctypes.byref(bad), ctypes.byref(bad), 1, 1, 8, 8, 0, 0, None) == -4
    # host entry point: NULL host pointers
    assert L.ssm_synthesize_host(None, None, None, None, None, None, 1, 1, 8, 8, 0, None, 0) == -1
    # packed frames buffer must be 16-byte aligned
    ok = _abi.SsmTensor(64, 64, 0, 64)
    assert L.ssm_pack_frames(ctypes.byref(ok), ctypes.c_void_p(8), 1, 8, 8, 0, None) == -4


def test_workspace_sizes():
    L = _abi.lib()
    npx = 32 * 48
    assert L.ssm_warp_bwd_workspace_bytes(2, 3, 32, 48) == 256 + 8 * 2 * 3 * npx
    assert L.ssm_flow_pack_bwd_workspace_bytes(2, 7, 32, 48) == 256 + 12 * 2 * 6 * npx
    # round 2: no B x N x 6 staging buffer any more (the scatter pass recomputes d/d(warped frame))
    assert L.ssm_fuse_bwd_workspace_bytes(2, 7, 32, 48) == 256 + 8 * 2 * 6 * npx
    assert L.ssm_warp_bwd_workspace_bytes(0, 3, 32, 48) == 0
    assert L.ssm_packed_frames_bytes(2, 32, 48, 0) == 2 * 2 * npx * 16
    assert L.ssm_packed_frames_bytes(2, 32, 48, 1) == 2 * 2 * npx * 8
    assert L.ssm_synthesize_host_scratch_bytes(5, 7, 32, 48) >= 3 * 4 * npx * (6 + 4 + 8 + 24 * 7)


def test_cpu_tensors_are_refused():
    x = torch.zeros(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ssm_b200.warp(x, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ssm_b200.flow_pack(torch.zeros(1, 6, 8, 8), torch.zeros(1, 4, 8, 8), torch.tensor([0.5]))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ssm_b200.fuse(torch.zeros(1, 6, 8, 8), torch.zeros(1, 1, 16, 8, 8), torch.zeros(1, 1, 5, 8, 8), torch.tensor([0.5]))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", "/nonexistent/libssm_b200.so")
    with pytest.raises(RuntimeError, match="no\\s+CPU or eager fallback"):
        _abi.lib()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "superslomo-videointerpolation-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("SURVEY", ""), "%s mentions the oracle" % f


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI: include/ssm_b200.h compiles as C99 (no C++ or torch types) and a C program links
    against libssm_b200.so and reads the version without a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "ssm_b200.h"\n'
                   'int main(void) { printf("%d %d\\n", ssm_version(), SSM_ABI_VERSION);\n'
                   '  return ssm_warp_fwd(0, 0, 0, 1, 3, 8, 8, SSM_DTYPE_F32, SSM_COORD_DIV, 0) < 0 ? 0 : 1; }\n')
    libdir = os.path.dirname(ssm_b200._abi.LIB_PATH)
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-l:libssm_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr                    # NULL tensors are an argument error (< 0), not a crash
    a, b = out.stdout.split()
    assert a == b == str(ssm_b200.abi_version())

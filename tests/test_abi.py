"""The C-ABI library: builds for sm_100a without a GPU, loads, and exports every symbol that
include/nerfb200.h declares.  No compute calls (CPU only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nerfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nerfb200_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from nerf_experiments_b200 import _lib
    path = _lib.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 19
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_error_channel():
    from nerf_experiments_b200 import _lib
    L = _lib.lib()
    assert L.nerfb200_abi_version() == 6
    # argument validation happens before any CUDA call: callable without a GPU
    rc = L.nerfb200_composite_fwd(None, None, None, None, 4, 8, 0, None, None, None, None, None)
    assert rc == 1
    assert b"null pointer" in L.nerfb200_last_error()
    rc = L.nerfb200_resample_alloc(None, None, None, 4, 8, 4, 8.0, None, None, None, None, None)
    assert rc == 1 and b"Sf>Sc" in L.nerfb200_last_error()


def test_struct_sizes_match_the_header():
    """ctypes mirrors in _lib.py must have the C layout of include/nerfb200_mlp.h."""
    import subprocess
    import tempfile
    from nerf_experiments_b200 import _lib
    src = r'''
#include <stdio.h>
#include "nerfb200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(NbBlock), sizeof(NbOp), sizeof(NbProgram),
  sizeof(NbPeCfg), sizeof(NbMlpInputs), sizeof(NbPackChunk), sizeof(NbPackBias), sizeof(NbWgradItem),
  sizeof(NgStep), sizeof(NgBlock), sizeof(NgOp), sizeof(NgProgram), sizeof(NbAdamGroup), sizeof(NbGaussLayer)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [_lib.NbBlock, _lib.NbOp, _lib.NbProgram, _lib.NbPeCfg, _lib.NbMlpInputs, _lib.NbPackChunk,
               _lib.NbPackBias, _lib.NbWgradItem, _lib.NgStep, _lib.NgBlock, _lib.NgOp, _lib.NgProgram, _lib.NbAdamGroup, _lib.NbGaussLayer]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "nerf_experiments_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn

"""The C-ABI library loads, exports every symbol include/halotrace_b200.h declares, and the ctypes mirror
has the same struct sizes as the C compiler's. No compute calls (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import harness as H
from ice_halo_sim_b200 import _abi as A
from ice_halo_sim_b200 import lib as L

HEADER = os.path.join(H.ROOT, "include", "halotrace_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in L.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(L.PROTOTYPES) == names
    assert lib.hb_abi_version() == 2


def test_struct_sizes_match_the_c_compiler():
    names = [s.__name__ for s in A.ALL_STRUCTS]
    prog = '#include <stdio.h>\n#include "halotrace_b200.h"\nint main(void){\n' + \
        "".join(f'printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(H.ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    sizes = dict(line.split() for line in out.strip().splitlines())
    for s in A.ALL_STRUCTS:
        assert int(sizes[s.__name__]) == C.sizeof(s), s.__name__
    assert C.sizeof(A.HbExitRecord) == 96       # lumice::ExitRayRecord, exit_seam.hpp:52
    assert C.sizeof(A.HbWlEntry) == 20          # WlEntry, wl_pool.hpp:36
    assert C.sizeof(A.HbProjParams) == 76       # lm_proj::ProjParams


def test_no_cpu_fallback():
    """Without a usable sm_100 device creation fails loudly (BackendUnavailableError), it never falls back."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    from ice_halo_sim_b200 import B200TraceBackend, BackendUnavailableError
    with pytest.raises(BackendUnavailableError):
        B200TraceBackend(0)


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    pkg = os.path.join(H.ROOT, "ice_halo_sim_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for needle in ("halo_oracle", "libhalo_ref", "ref_driver", "oracle/_", "import harness",
                               "from oracle", "import oracle"):
                    assert needle not in txt, (f, needle)

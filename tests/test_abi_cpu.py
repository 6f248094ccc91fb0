"""The C-ABI boundary on a box without a GPU: libkamr.so loads, exports every symbol include/kamr.h declares,
the ctypes mirror has the C layout, and the product path fails loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "kamr.h")


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kamr_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(kamr_lib):
    from kitamr_jl_b200 import abi
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(kamr_lib, n), f"{n} declared in include/kamr.h but not exported by libkamr.so"
    assert sorted(abi.EXPORTS) == names, "abi.EXPORTS must list exactly the header's entry points"
    assert kamr_lib.kamr_version() == 1


def test_no_torch_or_cxx_types_in_header():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)   # comments may mention cudaStream_t
    for bad in ("torch", "at::", "std::", "cudaStream_t", "template"):
        assert bad not in src
    assert 'extern "C"' in src


def test_ctypes_layout_matches_c(tmp_path):
    """sizeof/offsetof of every struct that crosses the boundary, C compiler vs ctypes mirror."""
    from kitamr_jl_b200 import abi
    structs = {"kamr_config": abi.KamrConfig, "kamr_mesh": abi.KamrMesh, "kamr_ib": abi.KamrIB,
               "kamr_stats": abi.KamrStats, "kamr_kernel_time": abi.KamrKernelTime, "kamr_vs_adapt": abi.KamrVsAdapt}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HDR}"', 'int main(void){']
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in ct._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for f, _ in ct._fields_:
            assert int(out[f"{cname}.{f}"]) == getattr(ct, f).offset, f"{cname}.{f}"


def test_create_fails_loudly_without_gpu(kamr_lib):
    """No CPU fallback: without a CUDA device kamr_create must return an error and say why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kitamr_jl_b200 import abi, api
    cfg = abi.KamrConfig(2, 2, 0, 0, 1.0, 2 / 3, 5 / 3, 0.81, 0.1, 0, 0, 1, None)
    with pytest.raises(api.KamrError) as e:
        api.Context(cfg)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    # bad DIM/NDF is rejected before touching the device
    h = C.c_void_p()
    bad = abi.KamrConfig(2, 1, 0, 0, 1.0, 2 / 3, 5 / 3, 0.81, 0.1, 0, 0, 1, None)
    assert kamr_lib.kamr_create(C.byref(bad), C.byref(h)) != 0
    assert b"unsupported DIM/NDF" in kamr_lib.kamr_last_error(None)
    assert kamr_lib.kamr_create(None, C.byref(h)) != 0


def test_missing_library_raises(tmp_path):
    from kitamr_jl_b200 import abi
    with pytest.raises(RuntimeError) as e:
        abi.load(str(tmp_path / "nope.so"))
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    pkg = os.path.join(ROOT, "kitamr.jl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f)).read()
                for l in txt.splitlines():
                    if "oracle" in l:
                        assert "import" not in l and "#include" not in l and "CDLL" not in l, f"{f}: {l}"

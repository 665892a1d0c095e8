"""CPU tests of the boundary: the CUDA library loads here (no GPU) and exports exactly what
include/tribs_gpu.h declares; init fails loudly without a device (no silent CPU path)."""
import ctypes as C
import re

import numpy as np
import pytest

from refrun import ROOT, load_golden


def _declared():
    src = open(f"{ROOT}/include/tribs_gpu.h").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tribs_gpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from tribs_b200 import gpu
    L = gpu.load()
    names = _declared()
    assert len(names) >= 19
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/tribs_gpu.h but not exported"
    assert sorted(gpu.EXPORTS) == names
    import re as _re
    want = int(_re.search(r"#define\s+TRIBS_GPU_ABI_VERSION\s+(\d+)", open(f"{ROOT}/include/tribs_gpu.h").read()).group(1))
    assert L.tribs_gpu_abi_version() == want == 4


def test_field_table_matches_between_header_and_binding(built):
    from tribs_b200 import gpu
    assert gpu.FIELDS[:7] == ["NwtNew", "MuNew", "MiNew", "NtNew", "NfNew", "RuNew", "RiNew"]
    assert gpu.FIELDS[7:14] == ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld"]
    assert len(gpu.FIELDS) <= 64   # field masks are 64-bit


def test_init_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tribs_b200 import gpu
    basin, _, _ = load_golden("deep")
    with pytest.raises(gpu.TribsGpuError) as e:
        gpu.GpuBasin(basin)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    import os
    bad = []
    for base, _, files in os.walk(f"{ROOT}/tribs_b200"):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f)).read()
                if re.search(r"(from|import)\s+oracle|oracle/|tribs_oracle|pyoracle", txt):
                    bad.append(f)
    assert not bad, f"product sources mention the oracle: {bad}"


def test_ctypes_structs_have_the_layout_of_the_c_header(tmp_path):
    """Every struct that crosses the C ABI is declared twice: in include/tribs_gpu.h and as a ctypes
    Structure in tribs_b200/gpu.py. A tiny C program prints sizeof / offsetof for each field (a field
    missing on the C side fails to compile) and the numbers must equal ctypes' own."""
    import ctypes as C
    import subprocess
    from tribs_b200 import gpu
    pairs = [("tribs_mesh_t", gpu.Mesh), ("tribs_params_t", gpu.Params), ("tribs_state_t", gpu.StateT),
             ("tribs_part_t", gpu.Part), ("tribs_step_t", gpu.Step), ("tribs_et_params_t", gpu.EtParams),
             ("tribs_met_t", gpu.MetT), ("tribs_et_step_t", gpu.EtStep), ("tribs_snow_params_t", gpu.SnowParams),
             ("tribs_channel_t", gpu.Channel)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/tribs_gpu.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls), (cname, got[cname], C.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)

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
    assert L.tribs_gpu_abi_version() == 1


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

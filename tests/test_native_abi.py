"""CPU: libysi.so builds, loads, and exports exactly what include/ysi.h declares (no compute calls)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ysi.h")).read()
    return sorted(set(re.findall(r"YSI_API\s+[\w\s\*]+?\b(ysi_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from yolo_sam_inference_b200 import _native as nat
    assert _header_symbols() == sorted(nat.EXPORTS.keys())


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_library_loads_and_exports_every_symbol(precision):
    """Both operand-encoding builds (libysi_fp16.so, libysi.so) load side by side and export the same C ABI."""
    from yolo_sam_inference_b200 import _native as nat
    lib = nat.load(precision=precision)
    assert lib.ysi_version() == 1
    assert lib.ysi_operand_dtype().decode() == precision
    out = subprocess.run(["nm", "-D", "--defined-only", nat.lib_path(precision)], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (ysi_\w+)", out)))
    assert exported == _header_symbols()
    # nothing but the C ABI leaks out of the library
    assert not re.findall(r" T _ZN3ysi", out)


def test_metrics_struct_layout():
    from yolo_sam_inference_b200 import _native as nat
    d = nat.METRICS_DTYPE
    assert d.itemsize == 1192
    assert [d.fields[k][1] for k in ("area", "min_r", "perim_hist", "hull_area", "hull_perim_hist", "disk_n",
                                     "flags", "mask_hist")] == [0, 24, 40, 80, 88, 128, 152, 168]


def test_sass_contains_blackwell_tensor_and_tma_ops():
    """UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (B200_PROFILING.md table)."""
    from yolo_sam_inference_b200 import _native as nat
    nat.load()
    sass = subprocess.run(["cuobjdump", "-sass", nat.lib_path(nat.default_precision())], capture_output=True, text=True).stdout
    if not sass:
        return
    # + STTM = tcgen05.st (P written to tensor memory), USETMAXREG = setmaxnreg, UTMASTG / UTMAREDG = TMA store / reduce
    for op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "USETMAXREG"):
        assert op in sass, op
    assert "HMMA." not in sass.replace("UTCHMMA", "")       # no legacy mma.sync path


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from yolo_sam_inference_b200.sam_stage import SamStage
    with pytest.raises(RuntimeError):
        SamStage("vit_t", device="cuda:0")
    with pytest.raises(ValueError):
        SamStage("vit_t", device="cpu")

"""CPU: libmmf.so loads, exports every symbol include/mmf.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

import mmf
from mmf import _native as N
from oracle import mmf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_text():
    with open(os.path.join(ROOT, "include", "mmf.h")) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    declared = set(re.findall(r"\b(mmf_[a-z0-9_]+)\s*\(", header_text()))
    assert declared == set(N.EXPORTS)
    lib = ctypes.CDLL(mmf.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_constants_agree_between_header_python_and_oracle():
    h = header_text()
    assert int(re.search(r"#define MMF_P (\d+)", h).group(1)) == N.MMF_P == O.P == mmf.design.P
    assert float(re.search(r"#define MMF_PIVOT_TOL ([0-9.e+-]+)f", h).group(1)) == O.PIVOT_TOL
    assert float(re.search(r"#define MMF_CAL_TOL ([0-9.e+-]+)", h).group(1)) == O.CAL_TOL
    assert int(re.search(r"#define MMF_STATUS_EMPTY (\d+)", h).group(1)) == N.STATUS_EMPTY
    assert int(re.search(r"#define MMF_STATUS_RANKDEF (\d+)", h).group(1)) == N.STATUS_RANKDEF
    assert mmf.load_library().mmf_version() == int(re.search(r"#define MMF_VERSION (\d+)", h).group(1))


def test_struct_layouts_match_header():
    assert ctypes.sizeof(N.MmfConfig) == 48          # device, kernel, assume_finite, tc_variant, chunk_series, stream,
                                                     # host_narrow, host_threads, stream_solve, reserved1 (include/mmf.h)
    assert ctypes.sizeof(N.MmfStats) == 48


def test_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert mmf.device_count() == 0
    with pytest.raises(mmf.MmfError, match="no CUDA device"):
        mmf.ForecastEngine()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dss-ml-at-scale_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert "import oracle" not in src and "from oracle" not in src, f

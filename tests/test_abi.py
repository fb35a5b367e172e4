"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly the entry points
include/sgam_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "sgam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgam_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from sgam_neurips22_b200 import _lib, build
    build.build()
    lib = _lib.load()
    declared = header_functions()
    assert declared, "no prototypes found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sgam_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert lib.sgam_version() >= 100


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected on the host before any CUDA call and report through sgam_last_error()."""
    from sgam_neurips22_b200 import _lib
    lib = _lib.load()
    assert lib.sgam_vq_nearest(None, None, 0, 0, 0, None, None, None, None, None) == -1
    assert b"vq_nearest" in lib.sgam_last_error()
    assert lib.sgam_conv2d(None, None, None, None, None, 1, 8, 8, 4, 4, 3, 1, 0, 0, 0, None) == -1
    assert lib.sgam_splat_workspace_bytes(2, 16, 16) == 2 * 16 * 16 * 8
    assert lib.sgam_gn_splits(65536) == 128 and lib.sgam_gn_splits(16) == 8 and lib.sgam_gn_splits(4) == 2 and lib.sgam_gn_splits(1) == 1


def test_ops_refuse_cpu_tensors():
    import torch
    from sgam_neurips22_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.median_blur3(torch.zeros(1, 1, 4, 4))
    with pytest.raises(NotImplementedError):
        ops.dataset_id("kitti360")

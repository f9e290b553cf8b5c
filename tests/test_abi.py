"""The C-ABI shared library loads without a GPU and exports every entry point that
include/rslf_b200.h declares; compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rslf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(rslf_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("rslf_cuda_create", "rslf_cuda_upload_epis", "rslf_cuda_depth1d_pile", "rslf_cuda_depth2d",
                 "rslf_cuda_fine_to_coarse", "rslf_cuda_downsample_epis", "rslf_cuda_fuse_disp_maps",
                 "rslf_cuda_selective_median", "rslf_cuda_edge_confidence", "rslf_cuda_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from remotesensingproject_b200 import api
    lib = api.lib()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, "declared in include/rslf_b200.h but not exported: %s" % missing
    assert lib.rslf_cuda_abi_version() >= 2


def test_params_default_matches_reference_defaults():
    from remotesensingproject_b200 import api
    p = api.default_params()
    # rslf_depth_computation_core.hpp:16-31
    assert np.float32(p.edge_score_threshold) == np.float32(0.02)
    assert p.raw_score_threshold == 0.0
    assert p.mean_shift_max_iter == 10
    assert p.edge_confidence_filter_size == 9
    assert p.edge_confidence_opening_size == 1
    assert p.median_filter_size == 5
    assert np.float32(p.median_filter_epsilon) == np.float32(0.1)
    assert np.float32(p.propagation_epsilon) == np.float32(0.1)
    assert p.slope_factor == 1.0 and p.cut_shadows == 1
    assert np.float32(p.shadow_level) == np.float32(0.05 * 1.73205080757)
    assert np.float32(p.kernel_h) == np.float32(0.2)
    assert ctypes.sizeof(api.Params) == 15 * 4


def test_params_struct_matches_oracle_struct():
    import oracle
    from remotesensingproject_b200 import api
    assert [f[0] for f in api.Params._fields_] == [f[0] for f in oracle.Params._fields_]
    a, b = api.default_params(), oracle.default_params()
    for name, _ in api.Params._fields_:
        assert getattr(a, name) == getattr(b, name), name


def test_no_cpu_fallback():
    """Without a CUDA device creating a context raises instead of computing on the host."""
    import torch
    from remotesensingproject_b200 import api
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.RslfError):
        api.Context(0)
    with pytest.raises(api.RslfError):
        api.Depth2DComputer(np.zeros((4, 3, 16, 3), np.float32), -1.0, 1.0, 8)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "remotesensingproject_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f

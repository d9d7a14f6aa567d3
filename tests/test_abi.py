"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/*.h declares
(no compute calls -- there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "spearmint_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = _declared()
    for need in ("smk_cov_build_f32", "smk_potrf_lower_batched_f32", "smk_chol_solve_f32", "smk_predict_f32",
                 "smk_ei_sweep_f32", "smk_topk_f32", "smk_ei_over_hypers_host_f32"):
        assert need in names


def test_library_exports_every_declared_symbol():
    from spearmint_b200 import _lib
    L = _lib.lib()
    for name in _declared():
        assert hasattr(L, name), "libspearmint_b200.so does not export %s" % name
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == _declared()


def test_info_calls_work_without_a_gpu():
    from spearmint_b200 import _lib
    L = _lib.lib()
    assert L.smk_version() >= 100
    assert L.smk_npad(1) == 128 and L.smk_npad(128) == 128 and L.smk_npad(129) == 256
    assert L.smk_block(4) == 128 and L.smk_block(8) == 64
    assert L.smk_topk_workspace_bytes(100000, 20) > 0


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from spearmint_b200 import _lib
    from spearmint_b200.engine import GPEIEngine
    with pytest.raises(_lib.SmkError):
        GPEIEngine()

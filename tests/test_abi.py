"""The C-ABI shared library loads without a GPU and exports every symbol include/agd_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import HAS_GPU, ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "agd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(agd):
    lib = agd._native.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/agd_b200.h but not exported"
    # and the Python binding covers exactly the declared ABI
    assert sorted(agd.exported_symbols()) == names


def test_struct_layouts_match_header(agd):
    N = agd._native
    p = N.Params()
    N.lib().agd_default_params(ctypes.byref(p))
    assert (p.convergence_tol, p.num_iterations, p.reg_param, p.L0, p.beta, p.alpha, p.may_restart) == \
        (1e-4, 100, 0.0, 1.0, 0.5, 0.9, 1)          # AGD.scala:44-51
    assert p.Lexact == float("inf") and p.flags == 0
    assert ctypes.sizeof(N.Params) == N.lib().agd_sizeof_params() == 72
    assert ctypes.sizeof(N.Stats) == N.lib().agd_sizeof_stats() == 112


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(agd):
    with pytest.raises(agd.NativeError, match="no CPU fallback"):
        agd.Context(devices=[0]).parallelize([0.0, 1.0], [[1.0, 2.0], [3.0, 4.0]])


def test_unsupported_plugins_rejected(agd):
    class MyGradient(agd.Gradient):
        pass

    with pytest.raises(TypeError):
        agd.optimization._grad_kind(MyGradient())
    with pytest.raises(TypeError):
        agd.optimization._upd_kind(object())
    with pytest.raises(NotImplementedError):
        agd.LogisticGradient(numClasses=3)


def test_context_transport_arguments(agd):
    """Multi-process worlds name how the exchange is set up: NCCL ships the handles (needs the id exchange) or the host does
    (needs the handle exchange); checked before any native call."""
    with pytest.raises(ValueError, match="id_exchange"):
        agd.Context(devices=[0], world_size=2, first_rank=0)
    with pytest.raises(ValueError, match="handle_exchange"):
        agd.Context(devices=[0], world_size=2, first_rank=0, transport="ipc")
    with pytest.raises(ValueError, match="transport"):
        agd.Context(devices=[0], transport="mpi")
    c = agd.Context(devices=[0], world_size=2, first_rank=1, handle_exchange=lambda b: b + b)
    assert c.transport == "ipc" and c.world_size == 2 and c.first_rank == 1
    assert agd.Context(devices=[0, 1]).world_size == 2          # a single process owning several GPUs is a complete world
    assert agd._native.ABI_VERSION == 2 and agd._native.XCHG_HANDLE_BYTES == 192


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package or include/ may reference it."""
    pkg = os.path.join(ROOT, "spark-agd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text.replace("CPU oracle", "") and "import oracle" not in text and \
                    "liboracle" not in text, f"{f} references the oracle"

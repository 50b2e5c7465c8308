"""Host-side behaviour of the API shim that needs no GPU: module loading, device registry, error convention
(reference: ospray/api/API.cpp:37-66,171-187,321-329; ospray/common/OSPCommon.cpp:749-838)."""
import ctypes as C

import pytest
import torch

from ospray_b200 import capi


def test_load_module_and_unknown_module():
    L = capi.lib()
    assert L.ospLoadModule(b"cuda") == capi.OSP_NO_ERROR
    assert L.ospLoadModule(b"cuda") == capi.OSP_NO_ERROR  # idempotent
    assert L.ospLoadModule(b"does_not_exist") != capi.OSP_NO_ERROR


def test_unknown_device_type_is_an_error():
    L = capi.lib()
    L.ospLoadModule(b"cuda")
    assert not L.ospNewDevice(b"no_such_device")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cuda_device_fails_loudly_without_gpu():
    # the product has no CPU fallback: committing the device without a CUDA device must raise, not render
    with pytest.raises(capi.OSPRayError) as e:
        capi.init()
    assert "no CPU fallback" in str(e.value)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_api_calls_without_current_device_report_errors():
    # ASSERT_DEVICE (API.cpp:37-43): calls without a device return neutral values, they do not crash
    L = capi.lib()
    assert not L.ospNewWorld()
    assert not L.ospNewRenderer(b"pathtracer")


def test_python_package_refuses_missing_library(tmp_path, monkeypatch):
    import ospray_b200
    monkeypatch.setattr(ospray_b200, "LIB_DIR", str(tmp_path))
    with pytest.raises(RuntimeError, match="no Python/CPU fallback"):
        ospray_b200.lib_path("libospcuda.so")

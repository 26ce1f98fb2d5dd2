"""librtb200.so: it loads, exports every symbol include/rtb200.h declares, and (in a
GPU-less container) refuses to create a context instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    return sorted(set(re.findall(r"RTB_API[^;(]*?\b(rtb_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_documented_entry_points(binding):
    assert declared_symbols() == sorted(binding.EXPORTS)


def test_library_exports_every_declared_symbol(binding):
    lib = binding.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in rtb200.h but not exported"


def test_version_names_the_arch(binding):
    assert b"sm_100a" in binding.load().rtb_version()


def test_no_device_means_no_context(binding):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the refusal path is exercised in the GPU-less container")
    with pytest.raises(binding.RtbError) as e:
        binding.Context(0)
    assert e.value.status == -2  # RTB_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_null_arguments_are_rejected_without_a_device(binding):
    lib = binding.load()
    assert lib.rtb_context_create(0, None) == -1
    assert lib.rtb_scene_upload(None, None, 0) == -1
    assert lib.rtb_cancel(None) == -1
    assert lib.rtb_last_error(None) is not None


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ray_tracing-rendering_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "refbind" not in txt and "libref_oracle" not in txt and "oracle_port" not in txt, f

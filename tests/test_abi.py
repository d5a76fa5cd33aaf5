"""The C-ABI library loads on a CPU box and exports every symbol include/tcr_b200.h declares.
No compute calls here (there is no GPU): only symbol presence, argument validation and error reporting."""
import ctypes as C
import os
import re

import pytest

import tcresnet_b200  # noqa: F401
from tcresnet_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return L.load()


def test_header_and_binding_agree(lib):
    header = open(os.path.join(ROOT, "include", "tcr_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(tcr_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"libtcr_b200.so does not export {name}"


def test_abi_version_is_consistent_everywhere(lib):
    """Header, binding, library and the driver's build() check must agree (a mismatch fails the round's build gate)."""
    header = open(os.path.join(ROOT, "include", "tcr_b200.h")).read()
    declared = int(re.search(r"#define TCR_ABI_VERSION (\d+)", header).group(1))
    assert declared == L.ABI_VERSION == lib.tcr_abi_version()
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "_lib.ABI_VERSION" in entry


def test_every_entry_point_cites_the_reference():
    header = open(os.path.join(ROOT, "include", "tcr_b200.h")).read()
    for needle in ("datasets/preprocessors.py", "audio_nets/tc_resnet.py", "factory/audio_nets.py",
                   "helper/trainer.py", "helper/base.py", "common/model_loader.py"):
        assert needle in header


def test_default_config_is_the_baseline_shape(lib):
    c = L.TcrConfig()
    assert lib.tcr_config_default(C.byref(c)) == 0
    assert (c.model, c.num_classes, c.clip_samples, c.window_size_samples, c.window_stride_samples) == (8, 12, 16000, 640, 320)
    assert (c.num_mel_bins, c.num_mfccs) == (64, 40)
    assert abs(c.bn_decay - 0.997) < 1e-7 and abs(c.bn_epsilon - 1e-3) < 1e-9


def test_invalid_arguments_report_errors_without_a_gpu(lib):
    c = L.TcrConfig()
    lib.tcr_config_default(C.byref(c))
    h = C.c_void_p()
    assert lib.tcr_create(None, C.byref(h)) == 1
    c.max_batch = 0
    assert lib.tcr_create(C.byref(c), C.byref(h)) == 1 and b"max_batch" in lib.tcr_last_error()
    lib.tcr_config_default(C.byref(c))
    c.window_size_samples = 642
    assert lib.tcr_create(C.byref(c), C.byref(h)) == 3 and b"multiples of 4" in lib.tcr_last_error()
    lib.tcr_config_default(C.byref(c))
    c.dropout_keep_prob = 0.0
    assert lib.tcr_create(C.byref(c), C.byref(h)) == 1
    assert lib.tcr_destroy(None) == 0
    assert lib.tcr_get_info(None, None) == 1
    n = C.c_uint64(123)
    assert lib.tcr_launch_count(C.byref(n)) == 0 and n.value == 0      # nothing was launched on this CPU box


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(L.TcrError, match="no CPU fallback"):
        L.load(str(tmp_path / "libtcr_b200.so"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tc-resnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("tcr_oracle_free", ""), f"{f} mentions the oracle"
                assert "cuda_emu" not in src or f == "tcr_device.cuh", f"{f} references the test emulator"


def test_generated_register_dft_header_is_current():
    """csrc/tcr_fft_reg.cuh is generated (tools/gen_fft_reg.py): the committed file is what the generator emits today."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fft_reg.py")], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(ROOT, "tc-resnet_b200", "csrc", "tcr_fft_reg.cuh")).read()

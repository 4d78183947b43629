"""The C-ABI library loads and exports every symbol include/hpc_b200.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (REPO / "include" / "hpc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("hpc_attention_decode_fp8_async", "hpc_assign_attention_decode_task_sync",
                 "hpc_assign_attention_decode_task_async", "hpc_last_error"):
        assert must in syms


def test_library_exports_all_declared(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"symbols declared in include/hpc_b200.h but not exported: {missing}"


def test_version_and_error_strings(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    lib.hpc_version.restype = ctypes.c_char_p
    lib.hpc_built_json.restype = ctypes.c_char_p
    lib.hpc_last_error.restype = ctypes.c_char_p
    assert b"b200" in lib.hpc_version()
    assert b"sm_100a" in lib.hpc_built_json()
    assert lib.hpc_last_error() is not None


def test_python_package_imports_and_mirrors_reference_names(hpc):
    for name in ("attention_decode_fp8", "get_attention_decode_task_workspace",
                 "assign_attention_decode_task", "print_attention_decode_task", "QuantType"):
        assert hasattr(hpc, name), name
    assert hpc.QuantType.QPERTOKEN_PERHEAD_KPERTENSOR_VPERTENSOR.value == 1
    import torch

    assert hasattr(torch.ops.hpc, "attention_decode_fp8")
    assert hasattr(torch.ops.hpc, "assign_attention_decode_task")


def test_launcher_rejects_bad_arguments_without_gpu(lib_path):
    """Argument validation happens before any CUDA call: unsupported shapes fail loudly."""
    lib = ctypes.CDLL(str(lib_path))
    lib.hpc_last_error.restype = ctypes.c_char_p
    lib.hpc_assign_attention_decode_task_async.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] * 7 + [ctypes.c_void_p]
    rc = lib.hpc_assign_attention_decode_task_async(None, None, 148, 5000, 8, 1, 128, 1, 64, None)
    assert rc == 1 and b"batch" in lib.hpc_last_error()
    rc = lib.hpc_assign_attention_decode_task_async(None, None, 148, 4, 8, 1, 96, 1, 64, None)
    assert rc == 1 and b"tilen" in lib.hpc_last_error()

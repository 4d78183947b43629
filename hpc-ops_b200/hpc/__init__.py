"""hpc — B200 (sm_100a) build of the HPC-Ops quantized-inference hot path.

Same Python-over-torch API as the reference package (reference hpc/__init__.py:1-58): every
public function of the in-scope modules is re-exported at package level, and the same operators
are reachable as `torch.ops.hpc.*`. Underneath is a torch-free C-ABI library (`_C.so`, see
include/hpc_b200.h) of hand-written sm_100a kernels, bound with ctypes.
"""
import importlib
import sys
from pathlib import Path
from types import ModuleType
from typing import Dict

from . import _ffi

_pkg_dir = Path(__file__).parent

__all__ = []


def _discover_modules() -> Dict[str, ModuleType]:
    modules = {}
    for file in sorted(_pkg_dir.iterdir()):
        if file.suffix != ".py" or file.name.startswith("_"):
            continue
        name = file.stem
        try:
            modules[name] = importlib.import_module(f".{name}", package=__package__)
        except ImportError as e:  # same behaviour as the reference: warn, keep going
            print(f"WARNING: Failed to import {name}: {e}", file=sys.stderr)
    return modules


def _export_functions(modules: Dict[str, ModuleType]):
    for module in modules.values():
        funcs = {
            name: obj
            for name, obj in vars(module).items()
            if callable(obj) and not name.startswith("_")
        }
        globals().update(funcs)
        __all__.extend(funcs.keys())


_export_functions(_discover_modules())

__version__ = _ffi.lib.hpc_version().decode()
__built_json__ = _ffi.lib.hpc_built_json().decode()

# torch.ops.hpc.version() / built_json(), as registered by reference src/C/version.cc:14
from . import _ops as _ops_mod  # noqa: E402

_ops_mod.define("version() -> str")
_ops_mod.impl("version", lambda: __version__, "CompositeExplicitAutograd")
_ops_mod.define("built_json() -> str")
_ops_mod.impl("built_json", lambda: __built_json__, "CompositeExplicitAutograd")

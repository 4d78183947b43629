"""Registration of the `torch.ops.hpc.*` operators (schemas verbatim from the reference's
TORCH_LIBRARY_FRAGMENT blocks) with Python implementations that call the C-ABI library."""
import torch

_lib = torch.library.Library("hpc", "DEF")
_defined = set()


def define(schema: str):
    name = schema.split("(", 1)[0].strip()
    if name not in _defined:
        _lib.define(schema)
        _defined.add(name)
    return name


def impl(name: str, fn, key: str):
    _lib.impl(name, fn, key)

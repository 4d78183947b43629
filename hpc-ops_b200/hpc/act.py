"""Stand-alone activation / FP8 quantisation operators (API of reference hpc/act.py; the masked
DeepEP-layout variants are out of scope)."""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _ops
from ._ffi import check as _check_rc, lib as _lib, ptr as _ptr, stream_of as _stream_of


def _require(cond: bool, msg: str):
    if not cond:
        raise RuntimeError(msg)


def _act_mul_and_quant_impl(input, scale, use_bf16_mul, output):
    # reference src/activation/entry.cc:17-50
    _require(input.is_cuda and scale.is_cuda, "input and scale must be cuda tensors")
    _require(input.dtype == torch.bfloat16, "input dtype must be bfloat16")
    _require(input.is_contiguous(), "input tensor must be contiguous")
    _require(scale.dtype == torch.float32 and scale.numel() >= 1, "scale must be a float32 tensor")
    num_col = input.shape[-1]
    num_row = input.numel() // num_col if num_col else 0
    shape = list(input.shape)
    shape[-1] //= 2
    if output is None:
        output = torch.empty(shape, dtype=torch.float8_e4m3fn, device=input.device)
    else:
        _require(output.is_cuda and output.is_contiguous() and output.dtype == torch.float8_e4m3fn
                 and output.numel() == num_row * (num_col // 2),
                 "output must be a contiguous cuda fp8_e4m3 tensor of shape [..., C]")
    _check_rc(_lib.hpc_act_mul_and_quant_async(_ptr(output), _ptr(input), _ptr(scale), num_row,
                                               num_col, int(bool(use_bf16_mul)), _stream_of(input)),
              "act_mul_and_quant")
    return output


_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _scaled_fp8_quant_impl(input, scale, output):
    # reference src/activation/entry.cc:158-202
    _require(input.is_cuda, "input must be a CUDA tensor")
    _require(input.is_contiguous(), "input must be contiguous")
    _require(input.numel() > 0, "input must be non-empty")
    _require(input.dtype in _DTYPE_CODE, "input dtype must be float32, float16, or bfloat16")
    if output is None:
        output = torch.empty_like(input, dtype=torch.float8_e4m3fn)
    _require(output.is_cuda, "output must be a CUDA tensor")
    _require(output.is_contiguous(), "output must be contiguous")
    _require(output.shape == input.shape, "output shape must match input shape")
    _require(output.dtype == torch.float8_e4m3fn, "output dtype must be float8_e4m3fn")
    _require(scale is not None, "scale is required for scaled_fp8_quant")
    _require(scale.is_cuda, "scale must be a CUDA tensor")
    _require(scale.dtype == torch.float32, "scale dtype must be float32")
    _require(scale.numel() == 1, "scale must contain one element")
    _check_rc(_lib.hpc_scaled_fp8_quant_async(_ptr(output), _ptr(input), _ptr(scale), input.numel(),
                                              _DTYPE_CODE[input.dtype], _stream_of(input)),
              "scaled_fp8_quant")
    return output, scale


_ops.define("act_mul_and_quant(Tensor input, Tensor scale, bool use_bf16_mul, Tensor? output) -> "
            "(Tensor)")
_ops.impl("act_mul_and_quant", _act_mul_and_quant_impl, "CUDA")
_ops.define("scaled_fp8_quant(Tensor input, Tensor? scale, Tensor? output) -> (Tensor, Tensor)")
_ops.impl("scaled_fp8_quant", _scaled_fp8_quant_impl, "CUDA")


def act_mul_and_quant(gate_up: Tensor, scale: Tensor, use_bf16_mul: bool = True,
                      output: Tensor = None) -> Tensor:
    """silu(gate_up[:, :C]) * gate_up[:, C:] * scale[0] -> fp8_e4m3 [N, C]; gate_up bf16 [N, 2C]
    (reference hpc/act.py:7-35). With `use_bf16_mul` the product is formed in bf16 as the
    reference kernel does."""
    return torch.ops.hpc.act_mul_and_quant(gate_up, scale, use_bf16_mul, output)


def scaled_fp8_quant(input: Tensor, scale: Optional[Tensor] = None,
                     output: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Static per-tensor FP8 quantisation: (input / scale[0]).to(fp8_e4m3); returns (output, scale)
    (reference hpc/act.py scaled_fp8_quant; scale is required, src/activation/entry.cc:179)."""
    return torch.ops.hpc.scaled_fp8_quant(input, scale, output)


@torch.library.register_fake("hpc::act_mul_and_quant")
def _act_mul_and_quant_fake(input, scale, use_bf16_mul, output):
    shape = list(input.shape)
    shape[-1] //= 2
    return torch.empty(shape, dtype=torch.float8_e4m3fn, device=input.device)


@torch.library.register_fake("hpc::scaled_fp8_quant")
def _scaled_fp8_quant_fake(input, scale, output):
    return torch.empty_like(input, dtype=torch.float8_e4m3fn), scale

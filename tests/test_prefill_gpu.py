"""GPU parity tests of the FP8 block-sparse / dense prefill (both quant schemes) vs the CPU oracle,
at the reference's tolerance atol=0.1 (tests/test_attention_blocksparse_*_fp8.py:218/228) plus a
tighter relative-error bound."""
from pathlib import Path

import pytest
import torch

from oracle import prefill as op

pytestmark = pytest.mark.gpu


def _run(hpc, d, kpt):
    c = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    mask = c["block_mask"].to(torch.uint8).contiguous() if c["block_mask"] is not None else None
    qt = (hpc.QuantType.QPERTOKEN_PERHEAD_KPERTOKEN_PERHEAD_VPERHEAD if kpt
          else hpc.QuantType.QPERTOKEN_PERHEAD_KPERTENSOR_VPERTENSOR)
    return hpc.attention_with_kvcache_blocksparse_prefill_fp8(
        c["q"], c["kcache"], c["vcache"], c["qscale"], c["kscale"], c["vscale"], c["cu_seqlens_q"],
        c["block_ids"], c["seqlens_kv"], d["max_q"], quant_type=qt, block_mask=mask)


def _check(my, gt, tag, allow_nan_rows=False):
    my, gt = my.float().cpu(), gt.float()
    if allow_nan_rows:
        # a Q tile without any active KV tile is NaN in the reference too (hpc/attention.py:274-277):
        # the NaN rows must coincide, everything else is compared
        assert torch.equal(torch.isnan(my), torch.isnan(gt)), f"{tag}: NaN pattern differs"
        keep = ~torch.isnan(gt)
        my, gt = my[keep], gt[keep]
    assert torch.isfinite(my).all(), f"{tag}: non-finite"
    err = (my - gt).abs()
    rel = err.norm() / gt.norm().clamp_min(1e-6)
    assert torch.allclose(my, gt, atol=0.1), f"{tag}: max abs err {err.max():.4f}"
    assert rel < 0.03, f"{tag}: rel err {rel:.4f}"


def _oracle(d, kpt):
    return op.blocksparse_prefill(d["q"], d["kcache"], d["vcache"], d["qscale"], d["kscale"],
                                  d["vscale"], d["cu_seqlens_q"], d["seqlens_kv"], d["block_ids"],
                                  d["block_mask"], kpt)


@pytest.mark.parametrize("kpt", [False, True])
@pytest.mark.parametrize("seq", [1024, 2048])
@pytest.mark.parametrize("skip", [None, 0.0, 0.5])
@pytest.mark.parametrize("layout", ["nhd", "hnd"])
def test_blocksparse_prefill_reference_grid(hpc, kpt, seq, skip, layout):
    """reference grid: B=2, (Hq, Hkv)=(4, 1), seq in {1024, 2048}, skip in {0, 0.5}, nhd/hnd;
    skip=None is the dense (no mask) path."""
    d = op.make_inputs([seq, seq], [seq, seq], 4, 1, skip, kpt, layout=layout)
    _check(_run(hpc, d, kpt), _oracle(d, kpt), f"kpt={kpt} seq={seq} skip={skip} {layout}")


@pytest.mark.parametrize("kpt", [False, True])
def test_blocksparse_prefill_ragged_and_chunked(hpc, kpt):
    """Lengths that are not tile multiples, q shorter than kv (chunked prefill over an existing
    cache), GQA 8/2, a request of one token."""
    q_lens = [1, 130, 257, 64]
    kv_lens = [1, 130, 900, 1000]
    d = op.make_inputs(q_lens, kv_lens, 8, 2, 0.4, kpt, seed=3)
    _check(_run(hpc, d, kpt), _oracle(d, kpt), f"ragged kpt={kpt}", allow_nan_rows=True)
    d = op.make_inputs(q_lens, kv_lens, 8, 2, None, kpt, seed=4, layout="hnd")
    _check(_run(hpc, d, kpt), _oracle(d, kpt), f"ragged dense kpt={kpt}")


@pytest.mark.parametrize("kpt", [False, True])
@pytest.mark.parametrize("num_seq_q", [100, 500, 1500, 3904])
@pytest.mark.parametrize("layout", ["nhd", "hnd"])
def test_dense_kvcache_prefill_fp8(hpc, kpt, num_seq_q, layout):
    """hpc.attention_with_kvcache_prefill_fp8: grid of reference
    tests/test_attention_with_kvcache_q*_prefill_fp8.py:92-101 (B=4 -> 2 here, kv 3904, q chunk of
    100..3904 tokens at the end of the cache), tolerance atol=0.05 (:233)."""
    B, S = 2, 3904
    d = op.make_inputs([num_seq_q] * B, [S] * B, 4, 1, None, kpt, layout=layout, seed=10086)
    c = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    qt = (hpc.QuantType.QPERTOKEN_PERHEAD_KPERTOKEN_PERHEAD_VPERHEAD if kpt
          else hpc.QuantType.QPERTOKEN_PERHEAD_KPERTENSOR_VPERTENSOR)
    my = hpc.attention_with_kvcache_prefill_fp8(
        c["q"], c["kcache"], c["vcache"], c["qscale"], c["kscale"], c["vscale"], c["cu_seqlens_q"],
        c["block_ids"], c["seqlens_kv"], d["max_q"], quant_type=qt)
    gt = _oracle(d, kpt)
    my, gt = my.float().cpu(), gt.float()
    assert torch.isfinite(my).all()
    assert torch.allclose(my, gt, atol=0.05), (my - gt).abs().max()


def test_blocksparse_prefill_short_mask_width(hpc):
    """Kb shorter than the causal extent: exactly one extra tile (index Kb) is visited
    (reference kernels.cuh:2195-2216)."""
    d = op.make_inputs([1024], [1024], 4, 1, 0.3, False, seed=5)
    d["block_mask"] = d["block_mask"][:, :, :, :5].contiguous()  # keep only 5 of the 8 KV columns
    _check(_run(hpc, d, False), _oracle(d, False), "short mask")


def test_blocksparse_prefill_golden(hpc):
    from test_oracle_prefill import load

    for name in ("prefill_kvpt.npz", "prefill_kpertoken.npz"):
        z, d, kpt, layout = load(name)
        if layout == 1:
            d["kcache"] = d["kcache"].transpose(1, 2).contiguous().transpose(1, 2)
            d["vcache"] = d["vcache"].transpose(1, 2).contiguous().transpose(1, 2)
        _check(_run(hpc, d, kpt), torch.from_numpy(z["out"]), name)


def test_umma_b_operand_mn_major_sw128(hpc):
    """Descriptor convention used by the prefill PV GEMM: A K-major SW128, B = V[keys, d] MN-major
    SW128 (d contiguous)."""
    import numpy as np
    from test_decode_gpu import _idesc, _rand_fp8, _run_umma, _sw128_image

    gen = torch.Generator().manual_seed(5)
    Pf, Pu = _rand_fp8((128, 128), gen)
    Vf, Vu = _rand_fp8((128, 128), gen)
    want = Pf @ Vf
    got = _run_umma(hpc, _sw128_image(Pu), _sw128_image(Vu), 128, _idesc(128, 128, 0, 1), 4,
                    (16, 1024, 2, 32), (16, 1024, 2, 4096))
    assert torch.equal(got, want), (got - want).abs().max()

"""GPU parity tests of hpc.rope_norm_store_kv / rope_norm_store_kv_fp8 vs the CPU oracle: grid and
tolerances of reference tests/test_rope.py:228-367 (atol 8e-2 for bf16 q / caches, atol 0.5 for the
dequantised fp8 q), plus exact checks the reference does not make: page-tail zeroing, untouched
pages, fp8 cache contents within one e4m3 step, padded decode batches."""
import pytest
import torch

from oracle import rope as orp
from synth import rope as sr

pytestmark = pytest.mark.gpu


def _cuda(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _unpadded(d):
    n = d["num_req"]
    if d["real_rows"] is None:
        return d["qkv"], d["num_seqlen"], d["q_index"], d["kv_indices"]
    return d["qkv"][:d["real_rows"]], d["num_seqlen"][:n], d["q_index"][:n + 1], d["kv_indices"][:n]


@pytest.mark.parametrize("hq,hkv,dim", [(8, 1, 128), (64, 8, 128), (4, 2, 64)])
@pytest.mark.parametrize("policy", [0, 1, 2])
@pytest.mark.parametrize("num_req", [7, 16])
@pytest.mark.parametrize("is_prefill,mtp", [(True, None), (False, 0), (False, 1)])
def test_rope_norm_store_kv(hpc, hq, hkv, dim, policy, num_req, is_prefill, mtp):
    d = sr.make_inputs(num_req, is_prefill, mtp, hq, hkv, dim, seed=num_req + policy)
    c = _cuda(d)
    out_q = hpc.rope_norm_store_kv(
        c["kcache"], c["vcache"], c["qkv"], c["cos_sin"], c["num_seqlen"], c["q_index"],
        c["kv_indices"], is_prefill, q_norm_weight=c["q_norm_w"] if policy else None,
        k_norm_weight=c["k_norm_w"] if policy else None, qk_norm_policy=policy)
    qkv, ns, qi, ki = _unpadded(d)
    kc, vc = d["kcache"].clone(), d["vcache"].clone()
    ref_q = orp.rope_norm_store_kv(kc, vc, qkv, d["cos_sin"], ns, qi, ki, d["q_norm_w"], d["k_norm_w"], policy)
    rows = ref_q.shape[0]
    assert torch.allclose(out_q[:rows].float().cpu(), ref_q.float(), atol=8e-2)
    assert torch.allclose(c["kcache"].float().cpu(), kc.float(), atol=8e-2)
    assert torch.allclose(c["vcache"].float().cpu(), vc.float(), atol=8e-2)
    assert torch.equal(c["vcache"].cpu(), vc)  # V is a copy: bit-exact, incl. zeroed tails, untouched pages


@pytest.mark.parametrize("hq,hkv,dim", [(8, 1, 128), (64, 8, 128)])
@pytest.mark.parametrize("policy", [0, 1, 2])
@pytest.mark.parametrize("quant_policy", [1, 2])
@pytest.mark.parametrize("num_req", [7, 16])
@pytest.mark.parametrize("is_prefill,mtp", [(True, None), (False, 0), (False, 1)])
def test_rope_norm_store_kv_fp8(hpc, hq, hkv, dim, policy, quant_policy, num_req, is_prefill, mtp):
    d = sr.make_inputs(num_req, is_prefill, mtp, hq, hkv, dim, seed=3 * num_req + policy)
    c = _cuda(d)
    ks = torch.tensor([0.1], device="cuda")
    vs = torch.tensor([0.1], device="cuda")
    q_scale_val = 2.0
    qsi = torch.tensor([1.0 / q_scale_val], device="cuda")
    k8, v8 = c["kcache"].to(torch.float8_e4m3fn), c["vcache"].to(torch.float8_e4m3fn)
    qlens = (d["q_index"][1:] - d["q_index"][:-1])
    max_seqlens = int(qlens.max()) if is_prefill else mtp + 1
    q8, qs, flag = hpc.rope_norm_store_kv_fp8(
        k8, v8, c["qkv"], c["cos_sin"], c["num_seqlen"], c["q_index"], c["kv_indices"], is_prefill,
        ks, vs, quant_policy, max_seqlens=max_seqlens, q_scale_inv=qsi if quant_policy == 2 else None,
        q_norm_weight=c["q_norm_w"] if policy else None, k_norm_weight=c["k_norm_w"] if policy else None,
        qk_norm_policy=policy)
    assert flag.shape == (d["num_seqlen"].shape[0], hkv) and flag.dtype == torch.int32
    assert int(flag.abs().sum()) == 0
    qkv, ns, qi, ki = _unpadded(d)
    kc8, vc8 = d["kcache"].to(torch.float8_e4m3fn), d["vcache"].to(torch.float8_e4m3fn)
    ref8, ref_scale = orp.rope_norm_store_kv_fp8(
        kc8, vc8, qkv, d["cos_sin"], ns, qi, ki, ks.cpu(), vs.cpu(), quant_policy, qsi.cpu(), 448.0,
        d["q_norm_w"], d["k_norm_w"], policy)
    kc, vc = d["kcache"].clone(), d["vcache"].clone()
    ref_q = orp.rope_norm_store_kv(kc, vc, qkv, d["cos_sin"], ns, qi, ki, d["q_norm_w"], d["k_norm_w"], policy)
    rows = ref_q.shape[0]
    if quant_policy == 1:
        if is_prefill:
            pad128 = (max_seqlens + 127) // 128 * 128
            assert qs.shape == (d["num_seqlen"].shape[0], hq, pad128)
            mask = torch.arange(pad128).expand(qlens.shape[0], pad128) < qlens.unsqueeze(1)
            scale = qs.cpu().permute(0, 2, 1)[mask]
        else:
            assert qs.shape == (d["qkv"].shape[0], hq)
            scale = qs.cpu()[:rows]
        assert torch.allclose(scale, ref_scale, rtol=2e-2, atol=1e-6)
        deq = q8[:rows].float().cpu() * scale[:, :, None]
    else:
        assert qs is None
        deq = q8[:rows].float().cpu() * q_scale_val
    assert torch.allclose(deq, ref_q.float(), atol=0.5)  # reference tolerance
    # tighter: the same e4m3 codes as the oracle's quantisation, up to one code where bf16 / fp32
    # rounding of the rotated value straddles a rounding boundary
    dq = (q8[:rows].float().cpu() - ref8.float()).abs()
    assert float((dq > 0.13 * ref8.float().abs().clamp_min(2 ** -6)).float().mean()) < 2e-3
    for mine, ref in ((k8, kc8), (v8, vc8)):
        a, b = mine.float().cpu(), ref.float()
        assert float(((a - b).abs() > 0.13 * b.abs().clamp_min(2 ** -6)).float().mean()) < 2e-3
    assert torch.equal(v8.cpu().view(torch.uint8), vc8.view(torch.uint8))  # V: exact (same product)


def test_rope_golden_fixtures(hpc):
    from test_oracle_rope import load_rope

    for name in ("rope_prefill_p2.npz", "rope_decode_p1.npz"):
        d, out, (num_req, is_prefill, mtp, hq, hkv, policy) = load_rope(name)
        c = _cuda(d)
        q = hpc.rope_norm_store_kv(c["kcache"], c["vcache"], c["qkv"], c["cos_sin"], c["num_seqlen"],
                                   c["q_index"], c["kv_indices"], is_prefill, c["q_norm_w"], c["k_norm_w"],
                                   qk_norm_policy=policy)
        assert torch.allclose(q.float().cpu(), out["q"].float(), atol=8e-2), name
        assert torch.allclose(c["kcache"].float().cpu(), out["kcache"].float(), atol=8e-2), name
        assert torch.equal(c["vcache"].cpu(), out["vcache"]), name


def test_rope_feeds_decode_attention(hpc):
    """The fp8 store produces exactly what attention_decode_fp8 consumes: rope -> decode equals
    decode on a cache quantised by the oracle."""
    from oracle import attention as oa

    hq, hkv, B = 32, 8, 4
    d = sr.make_inputs(B, False, 0, hq, hkv, 128, seed=5, max_num_kv_blocks=64, pad_decode=False)
    c = _cuda(d)
    ks = torch.tensor([0.05], device="cuda")
    vs = torch.tensor([0.05], device="cuda")
    k8 = (c["kcache"].float() / 0.05).to(torch.float8_e4m3fn)
    v8 = (c["vcache"].float() / 0.05).to(torch.float8_e4m3fn)
    q8, qs, _ = hpc.rope_norm_store_kv_fp8(k8, v8, c["qkv"], c["cos_sin"], c["num_seqlen"], c["q_index"],
                                           c["kv_indices"], False, ks, vs, 1, max_seqlens=1)
    lens = c["num_seqlen"]
    tm = hpc.get_attention_decode_task_workspace(B, int(lens.max()), hkv, 64)
    hpc.assign_attention_decode_task(lens, tm, hkv, 1, True, 64)
    blocks = c["kv_indices"].clamp_min(0).contiguous()
    y = hpc.attention_decode_fp8(q8, k8, v8, blocks, lens, qs, ks, vs, mtp=0, new_kv_included=True,
                                 task_map=tm)
    gt = oa.decode_fp8_kvpertensor(q8.cpu(), k8.cpu(), v8.cpu(), blocks.cpu(), lens.cpu(), qs.cpu(),
                                   ks.cpu(), vs.cpu(), 1)
    assert torch.allclose(y.float().cpu(), gt.float(), atol=0.2)

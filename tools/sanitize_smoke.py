"""One small invocation of every hot-path op (decode both quant types, prefill both, blockwise +
per-tensor MoE, route GEMM, W=1 allreduce, rope), meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

(the reference's replay harness does the same per call: conftest.py:74-159). Results are checked
against the oracle so that a clean sanitizer run is also a correct one."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "hpc-ops_b200"))
import torch  # noqa: E402

import hpc  # noqa: E402
from oracle import attention as oa, moe as om, prefill as op, rope as orp  # noqa: E402
from synth import rope as sr  # noqa: E402


def cuda(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def main():
    which = set(sys.argv[1:]) or {"decode", "prefill", "moe", "gemm", "allreduce", "rope"}
    if "decode" in which:
        B, hkv, hq = 3, 2, 8
        lens = [300, 129, 64]
        d = oa.make_decode_fp8_inputs(B, 1, lens, hkv, hq, seed=1, device="cuda")
        tm = hpc.get_attention_decode_task_workspace(B, max(lens), hkv, 64)
        hpc.assign_attention_decode_task(d["kv_lens_total"], tm, hkv, 1, True, 64)
        y = hpc.attention_decode_fp8(d["q"], d["kvcache"][:, 0], d["kvcache"][:, 1], d["block_ids"],
                                     d["kv_lens_total"], d["q_scale"], d["k_scale"], d["v_scale"],
                                     mtp=0, new_kv_included=True, task_map=tm)
        c = {k: v.cpu() for k, v in d.items()}
        gt = oa.decode_fp8_kvpertensor(c["q"], c["kvcache"][:, 0], c["kvcache"][:, 1], c["block_ids"],
                                       c["kv_lens_total"], c["q_scale"], c["k_scale"], c["v_scale"], 1)
        assert torch.allclose(y.float().cpu(), gt.float(), atol=0.2)
        d = oa.make_decode_fp8_kpt_inputs(B, 1, lens, hkv, hq, seed=2, device="cuda")
        y = hpc.attention_decode_fp8(d["q"], d["kcache"], d["vcache"], d["block_ids"], d["kv_lens_total"],
                                     d["q_scale"], d["k_scale"], d["v_scale"], mtp=0, new_kv_included=True,
                                     quant_type=hpc.QuantType.QPERTOKEN_PERHEAD_KPERTOKEN_PERHEAD_VPERHEAD,
                                     task_map=tm)
        c = {k: v.cpu() for k, v in d.items()}
        gt = oa.decode_fp8_kpertoken(c["q"], c["kcache"], c["vcache"], c["block_ids"], c["kv_lens_total"],
                                     c["q_scale"], c["k_scale"], c["v_scale"], 1)
        assert torch.allclose(y.float().cpu(), gt.float(), atol=0.1)
        print("decode ok")
    if "prefill" in which:
        for kpt in (False, True):
            d = op.make_inputs([300], [300], 4, 1, 0.5, kpt, seed=3)
            c = cuda(d)
            qt = (hpc.QuantType.QPERTOKEN_PERHEAD_KPERTOKEN_PERHEAD_VPERHEAD if kpt
                  else hpc.QuantType.QPERTOKEN_PERHEAD_KPERTENSOR_VPERTENSOR)
            y = hpc.attention_with_kvcache_blocksparse_prefill_fp8(
                c["q"], c["kcache"], c["vcache"], c["qscale"], c["kscale"], c["vscale"], c["cu_seqlens_q"],
                c["block_ids"], c["seqlens_kv"], d["max_q"], quant_type=qt,
                block_mask=c["block_mask"].to(torch.uint8).contiguous())
            gt = op.blocksparse_prefill(d["q"], d["kcache"], d["vcache"], d["qscale"], d["kscale"],
                                        d["vscale"], d["cu_seqlens_q"], d["seqlens_kv"], d["block_ids"],
                                        d["block_mask"], kpt)
            assert torch.allclose(y.float().cpu(), gt.float(), atol=0.1)
        print("prefill ok")
    if "moe" in which:
        d = om.make_moe_blockwise_inputs(96, 4, 256, 256, 8, 1, False, seed=5)
        c = cuda(d)
        y = hpc.fuse_moe_blockwise_fp8(c["x"], c["x_scale"], c["gate_up_weight"], c["gate_up_weight_scale"],
                                       c["down_weight"], c["down_weight_scale"], c["topk_ids"],
                                       c["topk_scale"], 0, 8)
        gt = om.fuse_moe_blockwise(d["x"], d["x_scale"], d["gate_up_weight"], d["gate_up_weight_scale"],
                                   d["down_weight"], d["down_weight_scale"], d["topk_ids"],
                                   d["topk_scale"], 0, None)
        err = (y.float().cpu() - gt.float()).abs()
        assert float((err > 0.01 + 0.01 * gt.float().abs()).float().mean()) < 1e-3
        print("moe ok")
    if "gemm" in which:
        from oracle import gemm as og
        x, w, wh, wl, scale = og.make_inputs(48, 192, 1024, device="cuda")
        ws = hpc.get_gemm_bf16xfp32_workspace(192, 4096)
        y = hpc.gemm_bf16xfp32(x, wh, wl, scale, True, True, ws)
        ref = og.gemm_split_exact(x, wh, wl, scale)
        assert float(((y.double() - ref).abs().max() / ref.abs().max())) < 1e-4
        print("gemm ok")
    if "rope" in which:
        d = sr.make_inputs(3, True, None, 8, 2, 128, seed=7, max_num_kv_blocks=32)
        c = cuda(d)
        q = hpc.rope_norm_store_kv(c["kcache"], c["vcache"], c["qkv"], c["cos_sin"], c["num_seqlen"],
                                   c["q_index"], c["kv_indices"], True, c["q_norm_w"], c["k_norm_w"],
                                   qk_norm_policy=2)
        kc, vc = d["kcache"].clone(), d["vcache"].clone()
        ref = orp.rope_norm_store_kv(kc, vc, d["qkv"], d["cos_sin"], d["num_seqlen"], d["q_index"],
                                     d["kv_indices"], d["q_norm_w"], d["k_norm_w"], 2)
        assert torch.allclose(q.float().cpu(), ref.float(), atol=8e-2) and torch.equal(c["vcache"].cpu(), vc)
        print("rope ok")
    torch.cuda.synchronize()
    print("smoke ok")


if __name__ == "__main__":
    main()

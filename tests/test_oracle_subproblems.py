"""CPU: the sampled-parity helpers of tools/bench_extras.py (used at BASELINE shapes, where the full
oracle is too large) must reproduce the full oracle on small shapes: MoE rows of chosen tokens via
the sub-problem {tokens} x {their experts}; prefill (head, Q-tile) items via chunked sub-requests;
allreduce rows via regenerated per-rank host inputs."""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "tools"))


def test_moe_token_subproblem_equals_full_oracle():
    import bench_extras as bx
    from oracle import moe as om
    from synth.moe import make_moe_blockwise_inputs

    d = make_moe_blockwise_inputs(64, 4, 256, 256, 16, 1, False, seed=3)
    full = om.fuse_moe_blockwise(d["x"], d["x_scale"], d["gate_up_weight"], d["gate_up_weight_scale"],
                                 d["down_weight"], d["down_weight_scale"], d["topk_ids"],
                                 d["topk_scale"], 0, None)
    toks = [0, 17, 63]
    sub, _, nexp = bx._oracle_moe_tokens(d, toks)
    assert nexp <= 12
    assert torch.equal(sub, full[toks])


def test_prefill_item_subproblem_equals_full_oracle():
    import bench_extras as bx
    from oracle import prefill as op

    for kpt in (False, True):
        d = op.make_inputs([1024], [1024], 4, 2, 0.5, kpt, seed=11)
        full = op.blocksparse_prefill(d["q"], d["kcache"], d["vcache"], d["qscale"], d["kscale"],
                                      d["vscale"], d["cu_seqlens_q"], d["seqlens_kv"], d["block_ids"],
                                      d["block_mask"], kpt)
        items = [(0, 0), (1, 3), (3, 7), (2, 7)]
        subs, _ = bx._oracle_prefill_items(d, kpt, items)
        for (h, t), o in zip(items, subs):
            ref = full[128 * t:128 * (t + 1), h]
            assert torch.allclose(o.float(), ref.float(), atol=1e-2, rtol=1e-2), (kpt, h, t)


def test_allreduce_rows_regeneration():
    import bench_extras as bx
    from oracle import allreduce as oa

    world, T, H = 3, 24, 256
    xs = [bx._cpu_inputs(r, T, H) for r in range(world)]
    residual = torch.randn((T, H), generator=torch.Generator().manual_seed(10000)).to(torch.bfloat16)
    weight = torch.randn((H,), generator=torch.Generator().manual_seed(9999)).to(torch.bfloat16)
    res, out = oa.allreduce_rmsnorm(xs, residual, weight, 1e-6)
    rows = torch.tensor([0, 5, 23])
    r2, o2 = bx._oracle_allreduce_rows(world, T, H, rows)
    assert torch.equal(r2, res[rows]) and torch.equal(o2, out[rows])

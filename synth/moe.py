"""Synthetic inputs for the blockwise FusedMoE (distributions of reference
tests/test_fuse_moe_blockwise.py:285-319)."""
import torch


def make_moe_blockwise_inputs(num_tokens, num_topk, hidden, inter, num_expert_total, size_ep=1,
                              shared=False, seed=41, device="cpu", wscale_abs=False):
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    E = num_expert_total // size_ep
    topk_ids = torch.multinomial(torch.ones((num_tokens, num_expert_total), device=dev), num_topk,
                                 replacement=False, generator=g).to(torch.int32)
    topk_ids, _ = torch.sort(topk_ids, dim=1)
    topk_scale = torch.rand((num_tokens, num_topk), generator=g, device=dev)
    topk_scale = topk_scale / topk_scale.sum(dim=1, keepdim=True)
    x = (torch.randn((num_tokens, hidden), generator=g, device=dev) / 100).to(torch.float8_e4m3fn)
    x_scale = torch.randn((num_tokens, hidden // 128), generator=g, device=dev)

    def weights(rows, cols):
        w = torch.empty((E, rows, cols), dtype=torch.float8_e4m3fn, device=dev)
        for e in range(E):
            w[e] = torch.randn((rows, cols), generator=g, device=dev).to(torch.float8_e4m3fn)
        return w

    guw = weights(inter * 2, hidden)
    guws = torch.randn((E, inter * 2 // 128, (hidden // 128 + 3) // 4 * 4), generator=g, device=dev)
    dw = weights(hidden, inter)
    dws = torch.randn((E, hidden // 128, (inter // 128 + 3) // 4 * 4), generator=g, device=dev)
    sh = torch.randn((num_tokens, hidden), generator=g, device=dev).to(torch.bfloat16) if shared else None
    return dict(x=x, x_scale=x_scale, gate_up_weight=guw, gate_up_weight_scale=guws, down_weight=dw,
                down_weight_scale=dws, topk_ids=topk_ids, topk_scale=topk_scale, shared_output=sh)

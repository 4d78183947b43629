"""Host-side logic of the sharded path on CPU: world_size-2 gloo processes reproduce the oracle
(reduce-scatter by owner slice -> residual + RMSNorm -> all-gather), i.e. the exact dataflow the
high-throughput kernel implements with multimem.ld_reduce / multimem.st."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, n, hidden, q):
    sys.path.insert(0, str(REPO))
    from oracle import allreduce as oa

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xs, residual, weight, n_pad = oa.make_inputs(world, n, hidden)
    ref_res, ref_out = oa.allreduce_rmsnorm(xs, residual, weight, 1e-6)
    s, e = oa.token_slice(n_pad, world, rank)
    # reduce-scatter: every rank sums its own token slice over all ranks (fp32 accumulate)
    mine = xs[rank].float().clone()
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    part_res = (mine[s:e] + residual[s:e].float()).to(torch.bfloat16)
    part_out = oa.rmsnorm(part_res, weight, 1e-6)
    # all-gather of the normalised slices
    outs = [torch.empty_like(part_out) for _ in range(world)]
    dist.all_gather(outs, part_out)
    full = torch.cat(outs, 0)
    ok = torch.allclose(full.float(), ref_out.float(), atol=0.1, rtol=0.1) and torch.allclose(
        part_res.float(), ref_res[s:e].float(), atol=0.1, rtol=0.1)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,hidden", [(77, 512), (128, 1024)])
def test_two_rank_gloo_dataflow_matches_oracle(n, hidden):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, hidden, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]


def test_token_slices_partition_the_padded_batch():
    from oracle import allreduce as oa

    for w in (1, 2, 4, 8):
        for n in (1, 77, 128, 4096):
            n_pad = (n + w - 1) // w * w
            cover = []
            for r in range(w):
                s, e = oa.token_slice(n_pad, w, r)
                cover += list(range(s, e))
            assert cover == list(range(n_pad))

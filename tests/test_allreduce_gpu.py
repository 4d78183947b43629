"""GPU parity tests of the fused AllReduce + residual + RMSNorm (the only sharded path).
World sizes that exceed the visible GPU count are skipped; W=1 runs on any box.
Processes are spawned like the reference's tests (tests/test_fuse_allreduce_rmsnorm_*.py)."""
import math
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _setup(rank):
    for p in (str(REPO), str(REPO / "hpc-ops_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.cuda.set_device(rank)


def _run_ht(rank, world, n, hidden, blocks, port, q):
    try:
        _setup(rank)
        os.environ["HPC_B200_COMM_PORT"] = str(port)
        import hpc
        from oracle import allreduce as oa

        dev = torch.device("cuda", rank)
        xs, residual, weight, n_pad = oa.make_inputs(world, n, hidden)
        ref_res, ref_out = oa.allreduce_rmsnorm([x[:n] for x in xs], residual[:n], weight, 1e-6)
        comm = hpc.MulticastCommunicator(rank, world, rank, f"ht_{world}_{n}_{hidden}_{blocks}")
        in_x, in_hdl = hpc.empty_multimem(comm, [n_pad, hidden], dtype=torch.bfloat16, device=dev)
        out_x, out_hdl = hpc.empty_multimem(comm, [n_pad, hidden], dtype=torch.bfloat16, device=dev)
        in_x.zero_()
        in_x[:n] = xs[rank][:n].to(dev)
        residual = residual.to(dev)
        weight = weight.to(dev)
        out_res = torch.empty_like(residual)
        s, e = oa.token_slice(n_pad, world, rank)
        off = s * hidden * 2
        comm.Barrier()
        for _ in range(3):  # repeated calls exercise the barrier slot reuse
            hpc.fuse_allreduce_rmsnorm_high_throughput(
                in_x[s:e], in_hdl.get_multimem_buff((e - s, hidden), torch.bfloat16, off),
                residual[s:e], weight, 1e-6, in_hdl.signal_buffer_ptrs_dev, rank, world, blocks,
                out_x[s:e], out_hdl.get_multimem_buff((e - s, hidden), torch.bfloat16, off),
                out_res[s:e])
        torch.cuda.synchronize()
        comm.Barrier()
        ok1 = torch.allclose(out_res[s:min(e, n)].float().cpu(), ref_res[s:min(e, n)].float(),
                             atol=0.1, rtol=0.1)
        ok2 = torch.allclose(out_x[:n].float().cpu(), ref_out.float(), atol=0.1, rtol=0.1)
        q.put((rank, bool(ok1 and ok2), f"mc={in_hdl.has_multicast}"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, False, repr(ex)[:500]))


def _run_ll(rank, world, seq, hidden, big_ws, two_shot, port, q):
    """`seq`: token counts of consecutive calls on ONE workspace (varying batch sizes exercise the
    dirty-buffer bookkeeping); big_ws: workspace large enough for the one-shot layout."""
    try:
        _setup(rank)
        os.environ["HPC_B200_COMM_PORT"] = str(port)
        import hpc
        from oracle import allreduce as oa

        dev = torch.device("cuda", rank)
        nmax = max(seq)
        xs, residual, weight, _ = oa.make_inputs(world, nmax, hidden)
        comm = hpc.MulticastCommunicator(rank, world, rank, f"ll_{world}_{nmax}_{hidden}_{big_ws}")
        m_pad = 2 * math.ceil(nmax / world) * world * 3  # reference sizing (test :54-60)
        if big_ws:
            m_pad = max(m_pad, nmax * world * 3)
        ws, hdl = hpc.empty_multimem(comm, [m_pad, hidden], dtype=torch.bfloat16, device=dev)
        ws.view(torch.int32).fill_(-2147483648)  # 0x80000000 words
        mc = hdl.get_multimem_buff([m_pad, hidden], dtype=torch.bfloat16)
        buf_bytes = (m_pad * hidden * 2 // 3) // 16 * 16
        flags = torch.tensor([0, 2, buf_bytes, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
        weight = weight.to(dev)
        torch.cuda.synchronize()
        comm.Barrier()
        ok = True
        for it, n in enumerate(seq):
            # different data every call: a stale row from an earlier call cannot pass
            xs_i = [torch.roll(x[:nmax], it + 1, 0)[:n].contiguous() for x in xs]
            res_i = torch.roll(residual[:nmax], it + 2, 0)[:n].contiguous()
            ref_res, ref_out = oa.allreduce_rmsnorm(xs_i, res_i, weight.cpu(), 1e-6)
            x = xs_i[rank].to(dev).contiguous()
            r = res_i.to(dev).contiguous()
            out = torch.empty_like(x)
            out_res = torch.empty_like(r)
            hpc.fuse_allreduce_rmsnorm_low_latency(x, mc, hdl.data_buffer_ptrs_dev, ws,
                                                   flags.view(torch.uint32), world, rank, r,
                                                   weight, 1e-6, 0, out, out_res, True,
                                                   use_two_shot=two_shot)
            torch.cuda.synchronize()
            ok = ok and torch.allclose(out_res.float().cpu(), ref_res.float(), atol=0.1, rtol=0.1)
            ok = ok and torch.allclose(out.float().cpu(), ref_out.float(), atol=0.1, rtol=0.1)
        comm.Barrier()
        q.put((rank, bool(ok), f"mc={hdl.has_multicast}"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, False, repr(ex)[:500]))


def _spawn(target, world, args):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    only = os.environ.get("HPC_B200_TEST_WORLDS")  # e.g. "4,8" on a multi-GPU lease: skip the rest
    if only and str(world) not in only.split(","):
        pytest.skip(f"world size {world} not selected by HPC_B200_TEST_WORLDS")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 21000 + (os.getpid() * 7 + hash(args) % 997) % 15000
    procs = [ctx.Process(target=target, args=(r, world, *args, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    res = []
    while not q.empty():
        res.append(q.get())
    for p in procs:
        if p.is_alive():
            p.kill()
    assert len(res) == world, f"only {len(res)} of {world} ranks reported: {res}"
    assert all(r[1] for r in res), res


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("n,hidden,blocks", [(128, 5120, 16), (77, 7168, 78), (256, 8192, 32)])
def test_allreduce_rmsnorm_high_throughput(world, n, hidden, blocks):
    _spawn(_run_ht, world, (n, hidden, blocks))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("n,hidden", [(128, 5120), (77, 7168), (8, 8192)])
def test_allreduce_rmsnorm_low_latency(world, n, hidden):
    """reference grid (tests/test_fuse_allreduce_rmsnorm_low_latency.py:121-125), two-shot protocol,
    5 calls so that every Lamport buffer is reused."""
    _spawn(_run_ll, world, ((n,) * 5, hidden, False, True))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("two_shot", [True, False])
def test_allreduce_rmsnorm_low_latency_varying_batch(world, two_shot):
    """Decode batches change size every step: a call must clear what the PREVIOUS call dirtied,
    not what its own size suggests (64, 8, 8, 64 left stale rows behind in round 1)."""
    _spawn(_run_ll, world, ((64, 8, 8, 64, 1, 33, 64, 2), 8192, not two_shot, two_shot))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n", [1, 8, 32])
def test_allreduce_rmsnorm_low_latency_one_shot(world, n):
    """one-shot protocol (workspace sized [tokens][world][hidden] per buffer)."""
    _spawn(_run_ll, world, ((n,) * 5, 8192, True, False))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_allreduce_rmsnorm_c5_shape(world):
    """BASELINE config C5: 4096 tokens x hidden 8192 (the W=2 shape once showed max_abs_err 20)."""
    _spawn(_run_ht, world, (4096, 8192, 148))

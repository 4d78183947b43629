"""Single-node symmetric / multicast memory bootstrap for the fused allreduce (B200 build).

Replaces the reference's hand-rolled substrate (src/communicator/*: unix-socket fd passing +
cuMemCreate / cuMulticast*) with NCCL-initialised torch.distributed plus
`torch.distributed._symmetric_memory` (CUDA VMM symmetric allocations with an NVLS multicast
mapping when the fabric supports it). Same Python surface as the reference class
`torch.classes.hpc.MulticastCommunicator` (src/communicator/entry.cc:79-90):
    MulticastCommunicator(rank, world_size, device_id, comm_name)
    .CreateTensorSync(nbytes) -> {rank: uint8 tensor (peer mapped), ..., -1: multicast tensor | None}
    .Barrier() .GetRank() .GetWorldSize() .GetDeviceId()
The reference's tests spawn bare processes, so the communicator bootstraps torch.distributed
itself (TCPStore on 127.0.0.1, port derived from comm_name) when no process group exists yet.
"""
import os as _os
import zlib as _zlib

import torch
import torch.distributed as _dist


class _PtrHolder:
    """Expose a raw device address range to torch through the CUDA array interface."""

    def __init__(self, ptr: int, nbytes: int, keepalive):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
        }
        self._keepalive = keepalive


def _tensor_from_ptr(ptr: int, nbytes: int, device, keepalive=None):
    t = torch.as_tensor(_PtrHolder(ptr, nbytes, keepalive), device=device)
    t._hpc_keepalive = keepalive
    return t


class MulticastCommunicator:
    def __init__(self, rank: int, world_size: int, device_id: int = -1, comm_name: str = "hpc"):
        self._rank = int(rank)
        self._world = int(world_size)
        self._device_id = int(device_id) if device_id >= 0 else torch.cuda.current_device()
        self._name = comm_name
        torch.cuda.set_device(self._device_id)
        self._owns_pg = False
        if not _dist.is_initialized():
            port = int(_os.environ.get("HPC_B200_COMM_PORT", 0)) or (
                20000 + (_zlib.crc32(comm_name.encode()) % 20000))
            store = _dist.TCPStore("127.0.0.1", port, self._world, is_master=(self._rank == 0),
                                   wait_for_workers=False)
            _dist.init_process_group("nccl", store=store, rank=self._rank, world_size=self._world,
                                     device_id=torch.device("cuda", self._device_id))
            self._owns_pg = True
        assert _dist.get_world_size() == self._world, "communicator world size mismatch"
        assert _dist.get_rank() == self._rank, "communicator rank mismatch"
        self._group = _dist.group.WORLD
        self._allocs = []

    # --- reference method names -----------------------------------------------------------
    def GetRank(self) -> int:
        return self._rank

    def GetWorldSize(self) -> int:
        return self._world

    def GetDeviceId(self) -> int:
        return self._device_id

    def Barrier(self) -> None:
        _dist.barrier(device_ids=[self._device_id])
        torch.cuda.synchronize(self._device_id)

    def has_multicast(self) -> bool:
        return bool(self._allocs) and self._allocs[-1][1].multicast_ptr != 0

    def CreateTensorSync(self, nbytes: int):
        """Collective: symmetric allocation of `nbytes` on every rank.
        Returns {r: uint8 tensor of rank r's buffer, -1: multicast tensor or None}."""
        import torch.distributed._symmetric_memory as symm_mem

        dev = torch.device("cuda", self._device_id)
        local = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(local, self._group)
        self._allocs.append((local, hdl))
        out = {}
        for r in range(self._world):
            out[r] = local if r == self._rank else hdl.get_buffer(r, (int(nbytes),), torch.uint8)
        mc_ptr = int(hdl.multicast_ptr) if self._world > 1 else 0
        out[-1] = _tensor_from_ptr(mc_ptr, int(nbytes), dev, keepalive=hdl) if mc_ptr else None
        return out

    def __del__(self):
        try:
            if self._owns_pg and _dist.is_initialized():
                _dist.destroy_process_group()
        except Exception:
            pass

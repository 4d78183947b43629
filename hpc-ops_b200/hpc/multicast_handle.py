"""Symmetric buffer + multicast view + signal pad (API of reference hpc/multicast_handle.py:7-200)."""
from itertools import accumulate as _accumulate
from operator import mul as _mul

import torch

_REGISTRY = []  # live handles, used to find peer pointers of a tensor in the P2P fallback


class MulticastHandle:
    def __init__(self, multicomm, size, dtype: torch.dtype = None):
        self.rank_ = multicomm.GetRank()
        self.world_size_ = multicomm.GetWorldSize()
        numel = list(_accumulate(size, func=_mul))[-1]
        self.buffer_size_ = numel * dtype.itemsize
        # data || 16-B aligned signal pad of 72 * SMs * 4 bytes (reference multicast_handle.py:25-38)
        signal_offset = (self.buffer_size_ + 15) // 16 * 16
        sms = torch.cuda.get_device_properties(multicomm.GetDeviceId()).multi_processor_count
        self.signal_size_ = 72 * sms * 4
        total = signal_offset + self.signal_size_
        self.org_buffer_dict_ = multicomm.CreateTensorSync(total)
        self.org_buffer_dict_[self.rank_][:] = 0
        self.data_buffer_list_ = [self.org_buffer_dict_[i][: self.buffer_size_]
                                  for i in range(self.world_size_)]
        mc = self.org_buffer_dict_[-1]
        self.multimem_data_buffer_ = mc[: self.buffer_size_] if mc is not None else None
        self.signal_buffer_list_ = [self.org_buffer_dict_[i][signal_offset:]
                                    for i in range(self.world_size_)]
        self.multimem_signal_buffer_ = mc[signal_offset:] if mc is not None else None
        self.data_buffer_ptrs_ = torch.tensor([t.data_ptr() for t in self.data_buffer_list_],
                                              dtype=torch.int64)
        self.signal_buffer_ptrs_ = torch.tensor([t.data_ptr() for t in self.signal_buffer_list_],
                                                dtype=torch.int64)
        device = self.org_buffer_dict_[self.rank_].device
        self.data_buffer_ptrs_dev_ = self.data_buffer_ptrs_.to(device=device)
        self.signal_buffer_ptrs_dev_ = self.signal_buffer_ptrs_.to(device=device)
        multicomm.Barrier()  # every rank's pad is zeroed before anyone signals
        _REGISTRY.append(self)

    @property
    def rank(self) -> int:
        return self.rank_

    @property
    def world_size(self) -> int:
        return self.world_size_

    @property
    def buffer_size(self) -> int:
        return self.buffer_size_

    @property
    def has_multicast(self) -> bool:
        return self.multimem_data_buffer_ is not None

    @property
    def data_buffer_ptrs(self):
        return self.data_buffer_ptrs_

    @property
    def data_buffer_ptrs_dev(self):
        return self.data_buffer_ptrs_dev_

    @property
    def signal_buffer_ptrs(self):
        return self.signal_buffer_ptrs_

    @property
    def signal_buffer_ptrs_dev(self):
        return self.signal_buffer_ptrs_dev_

    def _view(self, raw, size, dtype, storage_offset):
        numel = list(_accumulate(size, func=_mul))[-1]
        nbytes = numel * dtype.itemsize
        return raw[storage_offset: storage_offset + nbytes].view(dtype).reshape(tuple(size))

    def get_buffer(self, rank, size, dtype=torch.uint8, storage_offset: int = 0):
        """Tensor view of rank `rank`'s data buffer (storage_offset in bytes)."""
        return self._view(self.data_buffer_list_[rank], size, dtype, storage_offset)

    def get_multimem_buff(self, size, dtype=torch.uint8, storage_offset: int = 0):
        """Tensor view at the multicast (NVLS) address of the data buffer; when the fabric has no
        multicast support this returns the local view and the kernels use the P2P path."""
        if self.multimem_data_buffer_ is None:
            return self.get_buffer(self.rank_, size, dtype, storage_offset)
        return self._view(self.multimem_data_buffer_, size, dtype, storage_offset)

    def get_signal(self, rank):
        return self.signal_buffer_list_[rank]

    def contains(self, ptr: int):
        base = self.data_buffer_list_[self.rank_].data_ptr()
        return base <= ptr < base + max(self.buffer_size_, 1)

    def barrier(self):
        return None


def _find_handle(ptr: int):
    for h in reversed(_REGISTRY):
        if h.contains(ptr):
            return h
    return None

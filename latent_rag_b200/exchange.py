"""PeerExchange: the Python handle on a native lk_comm (include/latentknn.h) -- candidate
exchange between the GPUs of a row-sharded index over NVLink peer memory, fused with the final
k-way merge (one kernel per rank and search call; no collective library on the data path).

Net-new relative to the reference, which is a single process.
"""
from __future__ import annotations

from ctypes import byref, c_void_p
from typing import List, Optional, Tuple

import torch

from . import _native as nat


class PeerExchange:
    def __init__(self, device: int, rank: int, world: int, max_b: int = 4096, max_k: int = nat.LK_MAX_K_FUSED):
        self._lib = nat.load()
        nat.require_device()
        self.device, self.rank, self.world = int(device), int(rank), int(world)
        self.max_b, self.max_k = int(max_b), int(max_k)
        self._h = c_void_p()
        nat.check(self._lib.lk_comm_create(byref(self._h), self.device, self.rank, self.world, self.max_b,
                                           self.max_k), "lk_comm_create")

    # -- wiring ------------------------------------------------------------------------
    def ipc_handle(self) -> torch.Tensor:
        h = torch.zeros(nat.LK_IPC_HANDLE_BYTES, dtype=torch.uint8)
        nat.check(self._lib.lk_comm_ipc_handle(self._h, c_void_p(h.data_ptr())), "lk_comm_ipc_handle")
        return h

    def open_peers(self, handles: torch.Tensor) -> None:
        """handles: uint8 [world, LK_IPC_HANDLE_BYTES] on the host, rank-major."""
        handles = handles.to("cpu", torch.uint8).contiguous()
        assert handles.numel() == self.world * nat.LK_IPC_HANDLE_BYTES
        nat.check(self._lib.lk_comm_open_peers(self._h, c_void_p(handles.data_ptr())), "lk_comm_open_peers")

    def connect(self, group=None) -> "PeerExchange":
        """Exchange the IPC handles over the process group and map every peer's buffer."""
        import torch.distributed as dist

        if self.world > 1:
            dev = torch.device(f"cuda:{self.device}") if dist.get_backend(group) == "nccl" else torch.device("cpu")
            mine = self.ipc_handle().to(dev)
            allh = torch.empty(self.world * nat.LK_IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            self.open_peers(allh.cpu().view(self.world, -1))
            dist.barrier(group=group)  # nobody publishes into a buffer its owner has not finished setting up
        return self

    def attach_local(self, peers: List["PeerExchange"]) -> None:
        """Several ranks driven from one process (tests): map the peers' buffers directly."""
        for p in peers:
            if p is not self:
                nat.check(self._lib.lk_comm_attach_local(self._h, p.rank, p._h), "lk_comm_attach_local")

    # -- data path ---------------------------------------------------------------------
    def _stream(self) -> c_void_p:
        return c_void_p(int(torch.cuda.current_stream(self.device).cuda_stream))

    @staticmethod
    def _prep(d: torch.Tensor, i: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return d.to(torch.float32).contiguous(), i.to(torch.int64).contiguous()

    def exchange_merge(self, d: torch.Tensor, i: torch.Tensor, k: Optional[int] = None):
        """[b, k] local candidates (global ids, CUDA) -> the global top-k on every rank."""
        d, i = self._prep(d, i)
        b, kk = d.shape
        k = kk if k is None else int(k)
        out_d = torch.empty((b, k), dtype=torch.float32, device=d.device)
        out_i = torch.empty((b, k), dtype=torch.int64, device=d.device)
        for lo in range(0, b, self.max_b):
            hi = min(b, lo + self.max_b)
            nat.check(self._lib.lk_comm_exchange_merge(self._h, c_void_p(d[lo:hi].data_ptr()),
                                                       c_void_p(i[lo:hi].data_ptr()), hi - lo, k,
                                                       c_void_p(out_d[lo:hi].data_ptr()),
                                                       c_void_p(out_i[lo:hi].data_ptr()), self._stream()),
                      "lk_comm_exchange_merge")
        return out_d, out_i

    def begin(self) -> None:
        nat.check(self._lib.lk_comm_begin(self._h), "lk_comm_begin")

    def publish(self, d: torch.Tensor, i: torch.Tensor) -> None:
        d, i = self._prep(d, i)
        nat.check(self._lib.lk_comm_publish(self._h, c_void_p(d.data_ptr()), c_void_p(i.data_ptr()), d.size(0),
                                            d.size(1), self._stream()), "lk_comm_publish")
        torch.cuda.current_stream(self.device).synchronize()  # d, i may be temporaries

    def collect(self, b: int, k: int):
        dev = torch.device(f"cuda:{self.device}")
        out_d = torch.empty((b, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((b, k), dtype=torch.int64, device=dev)
        nat.check(self._lib.lk_comm_collect(self._h, b, k, c_void_p(out_d.data_ptr()), c_void_p(out_i.data_ptr()),
                                            self._stream()), "lk_comm_collect")
        return out_d, out_i

    def check(self) -> None:
        nat.check(self._lib.lk_comm_check(self._h), "lk_comm_check")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lk_comm_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

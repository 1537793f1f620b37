"""Peer-memory exchange for the SyncBatchNorm statistics (csrc/peer.cu): symmetric buffers mapped into every rank of the
process group (torch.distributed._symmetric_memory: cuMemMap over NVLink / NVSwitch), slot bookkeeping per layer and direction.

Plumbing only: the exchange itself is done by the kernels (`fsnet_bn_finalize_sync`, `fsnet_peer_allreduce_f64`).  When the
buffers can not be set up (single process, gloo / CPU emulation, symmetric memory unsupported on the box, FSNET_PEER_SYNCBN=0)
``get()`` returns None and the executor uses the process group's all_reduce instead (engine.py)."""
import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

_state = {"tried": False, "inst": None, "why": None}


class PeerExchange:
    DATA_BYTES = 32 << 20          # 68 slots x world x 2C doubles: 8 ranks, ResNet-50 channel counts -> ~20 MB
    FLAG_SLOTS = 4096

    def __init__(self):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", torch.cuda.current_device())
        group = dist.group.WORLD
        self.data = symm.empty(self.DATA_BYTES // 8, dtype=torch.float64, device=dev)
        self.flags = symm.empty(self.FLAG_SLOTS * self.world, dtype=torch.int32, device=dev)
        self.data.zero_()
        self.flags.zero_()
        self._hd = symm.rendezvous(self.data, group)
        self._hf = symm.rendezvous(self.flags, group)
        self.bufs_dev = int(self._hd.buffer_ptrs_dev)
        self.flags_dev = int(self._hf.buffer_ptrs_dev)
        self.seq = torch.zeros(self.FLAG_SLOTS, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier()               # nobody stores into a peer's buffers before every rank has zeroed its own
        self.slots: Dict[Tuple, _lib.Peer] = {}
        self._next_off = 0

    def slot(self, key, n: int) -> "_lib.Peer":
        """The exchange slot of `key` (layer, direction) for n doubles per rank; allocated on first use, in call order -- the
        same on every rank because every rank runs the same network."""
        pe = self.slots.get(key)
        if pe is None:
            i = len(self.slots)
            if i >= self.FLAG_SLOTS or (self._next_off + self.world * n) * 8 > self.DATA_BYTES:
                raise _lib.FsnetError("peer exchange buffers exhausted (raise PeerExchange.DATA_BYTES / FLAG_SLOTS)")
            pe = _lib.Peer(self.bufs_dev, self.flags_dev, self.rank, self.world, self._next_off, i * self.world,
                           self.seq.data_ptr() + 4 * i)
            pe.n = n
            self._next_off += self.world * n
            self.slots[key] = pe
        assert pe.n == n, "one exchange slot, two sizes"
        return pe


def get() -> Optional[PeerExchange]:
    """The process-wide exchange, or None when peer memory is not available (callers fall back to dist.all_reduce)."""
    if _state["tried"]:
        return _state["inst"]
    _state["tried"] = True
    if os.environ.get("FSNET_PEER_SYNCBN", "1") == "0":
        _state["why"] = "disabled by FSNET_PEER_SYNCBN=0"
        return None
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available()
            and dist.get_backend() == "nccl"):
        _state["why"] = "no NCCL process group"
        return None
    ok = torch.zeros(1, device="cuda")
    try:
        inst = PeerExchange()
        ok += 1
    except Exception as e:  # noqa: BLE001 - any failure of the optional fast path selects the collective fallback
        inst = None
        _state["why"] = f"{type(e).__name__}: {e}"
    dist.all_reduce(ok)              # all ranks or none
    if int(ok.item()) != dist.get_world_size():
        inst = None
    _state["inst"] = inst
    if dist.get_rank() == 0:
        print(f"[fsnet_b200] SyncBN statistics: {'NVLink peer memory (one-shot, fused into bn_finalize)' if inst else 'NCCL all_reduce (' + str(_state['why']) + ')'}",
              flush=True)
    return inst


def reset():
    _state.update(tried=False, inst=None, why=None)

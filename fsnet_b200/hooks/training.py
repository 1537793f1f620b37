"""Training / validation hooks (reference: vision_base/pipeline_hooks/train_val_hooks/
base_training_hooks.py:9-49 and base_validation_hooks.py:5-28)."""
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


class _LazyBatch(dict):
    """The graph's static batch.  Big tensors carry an *external* CUDA event recorded by the copy stream after their
    host->device copy; the FIRST access of such a key while the step is being captured inserts an event-wait node at
    exactly that point of the graph, so at replay the encoder starts as soon as ('image', 0) has landed while the
    loss-only tensors (original images, masks) are still in flight.  Tensors the model never touches are never waited for."""

    def __init__(self, base, events):
        super().__init__(base)
        self._events, self._waited = events, set()

    def _wait(self, key):
        ev = self._events.get(key)
        if ev is not None and key not in self._waited:
            torch.cuda.current_stream().wait_event(ev)
            self._waited.add(key)

    def __getitem__(self, key):
        self._wait(key)
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def _wait_all(self):
        for key in self._events:
            self._wait(key)

    def items(self):
        self._wait_all()
        return super().items()

    def values(self):
        self._wait_all()
        return super().values()

    def copy(self):
        self._wait_all()
        return dict(super().items())


class BaseTrainingHook(object):
    """One optimisation step: zero_grad, host->device, forward, ``loss.mean().backward()``,
    ``clip_grad_norm_``, ``optimizer.step()`` -- same order and semantics as the reference.

    ``cuda_graph=True`` (or env FSNET_CUDA_GRAPH=1) replays the whole step as ONE CUDA graph: the first
    ``graph_warmup`` calls run eagerly (each is a normal step), the next call captures
    zero_grad+forward+backward+clip+step and every call from then on copies its batch into the graph's
    static inputs and replays it.  One call is still exactly one optimiser step; the returned tensors
    are the graph's static outputs (valid until the next call).  ``overlap_h2d`` (opt-in, env FSNET_OVERLAP_H2D=1): the
    batch's big tensors are copied on a side stream and the captured step waits for each of them at its first use through
    external CUDA events (the graph contains event-wait nodes, tensors the model never reads are never waited for).
    Functionally validated on B200 (graph == eager trajectories), but the end-to-end step time did not change in the
    first measurement (9.55 ms with and without), so it stays off until the timeline is understood."""

    def __init__(self, tensor_keys: Optional[List[str]] = None, clip_gradients: Optional[float] = None, cuda_graph=None,
                 graph_warmup: int = 3, overlap_h2d=None, **kwargs):
        self.tensor_keys = tensor_keys
        self.clip_gradients = clip_gradients
        if cuda_graph is None:
            cuda_graph = os.environ.get("FSNET_CUDA_GRAPH", "0").lower() in ("1", "true")
        self.cuda_graph = bool(cuda_graph)
        self.graph_warmup = graph_warmup
        if overlap_h2d is None:
            overlap_h2d = os.environ.get("FSNET_OVERLAP_H2D", "0").lower() in ("1", "true")
        self.overlap_h2d = bool(overlap_h2d)     # graph mode: batch copies on a side stream, waited for inside the graph
        self._events, self._copy_stream, self._copy_order = {}, None, []
        self._calls = 0
        self._graph = None
        self._static_in = None
        self._static_out = None
        self._side_stream = None

    def _to_device(self, data):
        for key in data:
            if isinstance(data[key], torch.Tensor):
                if self.tensor_keys is None or key in self.tensor_keys:
                    data[key] = data[key].cuda(non_blocking=True).contiguous()
        return data

    @staticmethod
    def sync_gradients(meta_arch):
        """Data-parallel gradient averaging for models that are NOT wrapped in DistributedDataParallel (a DDP-wrapped model is left
        alone: its reducer does it).  Capturable into the step's CUDA graph, unlike DDP's reducer hooks.  The bulk of the bytes --
        the convolution weights -- was already averaged DURING backward, bucket by bucket, by the executor (engine.Tape); what is
        left here are the small BatchNorm / bias gradients (one concatenated all-reduce) and, as a fall-back, any flat gradient
        buffer the executor did not reduce (all-reduced in place: parameter gradients are views of it, no copies)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if os.environ.get("FSNET_DIAG_NO_GRADSYNC", "0") == "1":       # timing diagnostics only: ranks drift apart
            return
        if isinstance(meta_arch, nn.parallel.DistributedDataParallel):
            return
        from .. import engine
        world = dist.get_world_size()
        avg = dist.get_backend() == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        by_storage = {}
        for p in meta_arch.parameters():
            if p.grad is not None:
                by_storage.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p.grad)
        small = []
        for ptr, grads in by_storage.items():
            if ptr in engine.REDUCED_STORAGES:
                continue
            st = grads[0].untyped_storage()
            n = sum(g.numel() for g in grads)
            # the executor's flat gradient buffer: every gradient is a view of ONE storage -> one in-place all-reduce of the whole
            # storage (slots of parameters without a gradient ride along; they are never read)
            if len(grads) > 1 and n * 4 <= st.nbytes() and all(g.dtype == torch.float32 and g.is_contiguous() for g in grads):
                n = st.nbytes() // 4
                flat = torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, 0, (n,))     # the whole pool, in place
                dist.all_reduce(flat, op=op)
                if not avg:
                    flat.div_(world)
            else:
                small += grads
        if os.environ.get("FSNET_DEBUG_SYNC"):
            print(f"[fsnet_b200] sync_gradients: {len(by_storage)} storages, reduced earlier {len(engine.REDUCED_STORAGES)}, "
                  f"small tensors {len(small)} ({sum(g.numel() for g in small)} elements)", flush=True)
        engine.REDUCED_STORAGES.clear()
        if small and os.environ.get("FSNET_DIAG_NO_GRADSYNC", "0") != "small":       # ("small": timing diagnostics only)
            flat = torch.cat([g.reshape(-1) for g in small])
            dist.all_reduce(flat, op=op)
            if not avg:
                flat.div_(world)
            torch._foreach_copy_(small, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in small]), small)])

    def _step(self, data, meta_arch, optimizer, meta):
        from .. import engine
        engine.Tape.bucketed_allreduce = (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                                          and not isinstance(meta_arch, nn.parallel.DistributedDataParallel)
                                          and os.environ.get("FSNET_BUCKETED_ALLREDUCE", "0") == "1")
        optimizer.zero_grad()
        output: dict = meta_arch(data, meta)
        output["loss"].mean().backward()
        self.sync_gradients(meta_arch)
        if getattr(optimizer, "sync_hyperparams", None) is not None and len(optimizer.param_groups) == 1:
            optimizer.step(max_norm=self.clip_gradients)      # FusedAdam: clip_grad_norm_ folded into the update (2 launches)
        else:
            if self.clip_gradients is not None:
                torch.nn.utils.clip_grad_norm_(meta_arch.parameters(), self.clip_gradients)
            optimizer.step()
        return output

    def __call__(self, data: Dict, meta_arch: nn.Module, optimizer, writer=None, training_loss_logger=None,
                 global_step: int = 0, epoch_num: int = 0):
        meta = dict(epoch_num=epoch_num, global_step=global_step, is_training=True)
        if self.cuda_graph:
            output = self._graphed(data, meta_arch, optimizer, meta)
        else:
            output = self._step(self._to_device(data), meta_arch, optimizer, meta)
        if training_loss_logger is not None:
            training_loss_logger.update(output["loss_dict"])
            training_loss_logger.update_hm(output.get("hm", dict()))
        return output

    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def _tensorise(data):
        """Graph replay only refreshes TENSOR entries of the static batch.  The fisheye datasets deliver their per-sample MEI
        calibration as a list of dicts (`calib_meta`, fisheye_dataset.py:45-58,254; KITTI360FisheyeDataset picks image_02 or
        image_03 per sample): pack it into the `calib_mei` [B,3] fp64 tensor FishEyeDecoder reads, so that it is copied every step
        (replays used the capture batch's calibration for every later batch)."""
        if "calib_meta" in data and "calib_mei" not in data:
            from ..functional import MeiRayTable
            data["calib_mei"] = MeiRayTable.calib_host_tensor(data["calib_meta"])
        return data

    def _graphed(self, data, meta_arch, optimizer, meta):
        data = self._tensorise(data)
        if self._calls == 0:
            for group in optimizer.param_groups:          # Adam must keep its step counters on the device
                if "capturable" in group and not optimizer.state:
                    group["capturable"] = True
            self._side_stream = torch.cuda.Stream()
        self._calls += 1
        if self._graph is None and self._calls <= self.graph_warmup:
            # eager warm-up steps on a side stream (allocator, cuDNN/cuBLAS handles, lazily built kernel state)
            cur = torch.cuda.current_stream()
            self._side_stream.wait_stream(cur)
            with torch.cuda.stream(self._side_stream):
                out = self._step(self._to_device(data), meta_arch, optimizer, meta)
            cur.wait_stream(self._side_stream)
            return out
        if self._graph is None:
            dev = next(meta_arch.parameters()).device
            self._static_in = {}
            for k, v in data.items():
                if isinstance(v, torch.Tensor) and (self.tensor_keys is None or k in self.tensor_keys or k == "calib_mei"):
                    self._static_in[k] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                else:
                    self._static_in[k] = v
            self._events = {}
            if self.overlap_h2d:
                try:
                    big = [k for k, v in self._static_in.items() if isinstance(v, torch.Tensor) and v.numel() * v.element_size() >= (1 << 20)]
                    self._events = {k: torch.cuda.Event(external=True) for k in big}
                    self._copy_stream = torch.cuda.Stream()
                    # network inputs first (target frame first), loss-only tensors after them
                    rank = lambda k: (0 if (isinstance(k, tuple) and k[0] == "image") else 1, 0 if (isinstance(k, tuple) and k[-1] == 0) else 1)
                    self._copy_order = sorted(big, key=rank)
                except (TypeError, RuntimeError):            # torch without external events: plain in-order copies
                    self._events = {}
            self._copy_in(data)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            optimizer.zero_grad(set_to_none=True)
            lazy = _LazyBatch(self._static_in, self._events)
            # FSNET_GRAPH_PRIORITY=1 captures the step from a HIGH-priority stream, so that its kernels win the block scheduler against
            # the weight gradients on the executor's default-priority side stream; measured slower (6.04 vs 5.90 ms per step): off
            capture_stream = torch.cuda.Stream(priority=-1) if os.environ.get("FSNET_GRAPH_PRIORITY", "0") == "1" else None
            with torch.cuda.graph(self._graph, stream=capture_stream):
                self._static_out = self._step(lazy, meta_arch, optimizer, meta)
            if os.environ.get("FSNET_DEBUG_H2D"):
                print(f"[fsnet_b200] graph capture: {len(self._events)} external copy events, waited in-graph for {sorted(map(str, lazy._waited))}, "
                      f"copy order {list(map(str, self._copy_order))}", flush=True)
        else:
            self._copy_in(data)
        if getattr(optimizer, "sync_hyperparams", None) is not None:
            optimizer.sync_hyperparams()                  # scheduler changes of lr reach the captured step through device memory
        self._graph.replay()
        return self._static_out

    def _copy_in(self, data):
        main = torch.cuda.current_stream()
        if self._events:
            side = self._copy_stream
            side.wait_stream(main)                   # the previous replay has finished reading the static batch
            with torch.cuda.stream(side):
                for k in self._copy_order:
                    v = data.get(k)
                    if isinstance(v, torch.Tensor):
                        self._static_in[k].copy_(v, non_blocking=True)
                    self._events[k].record(side)
        for k, v in data.items():
            if k in self._events:
                continue
            dst = self._static_in.get(k)
            if isinstance(dst, torch.Tensor) and isinstance(v, torch.Tensor):
                dst.copy_(v, non_blocking=True)


class BaseValidationHook(object):
    """Inference call used by the evaluation hooks (base_validation_hooks.py:5-28)."""

    def __init__(self, tensor_keys: Optional[List[str]] = None, **kwargs):
        self.tensor_keys = tensor_keys

    def __call__(self, data: Dict, meta_arch: nn.Module, global_step: int = 0, epoch_num: int = 0) -> Dict:
        for key in data:
            if isinstance(data[key], torch.Tensor):
                if self.tensor_keys is None or key in self.tensor_keys:
                    data[key] = data[key].cuda().contiguous()
        return meta_arch(data, dict(epoch_num=epoch_num, global_step=global_step, is_training=False))

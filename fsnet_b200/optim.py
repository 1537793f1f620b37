"""Optimiser / scheduler factories and checkpoint helpers with the reference's names and behaviour
(vision_base/networks/optimizers/{optimizers,schedulers}.py, vision_base/networks/utils/utils.py:3-19)."""
import torch
import torch.nn as nn
import torch.optim as optim


class FusedAdam(optim.Adam):
    """torch.optim.Adam (same hyper-parameters, same ``state_dict`` layout: ``step`` / ``exp_avg`` / ``exp_avg_sq`` per
    parameter) whose ``step`` is two CUDA launches over a device-resident tensor table: ``fsnet_grad_sumsq`` (global gradient
    norm) and ``fsnet_adam_step`` (clip coefficient applied on the fly + the Adam update).  ``step(max_norm=...)`` fuses the
    ``clip_grad_norm_`` of base_training_hooks.py:46-47 into the update; hyper-parameters and the step counter live on the
    device, so the call is CUDA-graph capturable and follows ``param_group['lr']`` changes made by a scheduler.
    Parameters on the CPU (or amsgrad / maximize / several distinct groups' options) use the stock implementation."""

    CHUNK = 4096

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **kwargs):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kwargs)
        self._fused = {}          # group index -> dict(hyper, sumsq, table cache)

    def _fusable(self, group):
        ps = [p for p in group["params"] if p.grad is not None]
        return (bool(ps) and all(p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and not p.grad.is_sparse
                                 and p.is_contiguous() and p.grad.is_contiguous() for p in ps)
                and not group.get("amsgrad", False) and not group.get("maximize", False))

    def _group_state(self, gi, group, dev):
        st = self._fused.get(gi)
        if st is None:
            st = dict(hyper=torch.zeros(8, device=dev, dtype=torch.float64), sumsq=torch.zeros(1, device=dev, dtype=torch.float64),
                      host=torch.zeros(8, dtype=torch.float64).pin_memory(), key=None, table=None, step=None)
            self._fused[gi] = st
        return st

    def sync_hyperparams(self):
        """Copy lr / betas / eps / weight_decay of every group to the device (call outside a captured graph after a
        scheduler step; ``step`` calls it itself when not capturing)."""
        for gi, group in enumerate(self.param_groups):
            st = self._fused.get(gi)
            if st is None:
                continue
            h = st["host"]
            self._wait_staging(st, "host_ev")              # the previous copy out of this pinned buffer has executed
            h[0], h[1], h[2], h[3], h[4] = group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["weight_decay"]
            st["hyper"][:5].copy_(h[:5], non_blocking=True)
            self._mark_staging(st, "host_ev")

    @staticmethod
    def _wait_staging(st, key):
        ev = st.get(key)
        if ev is not None and not torch.cuda.is_current_stream_capturing():      # (a capture is preceded by a device synchronise)
            ev.synchronize()

    @staticmethod
    def _mark_staging(st, key):
        """A pinned staging buffer is rewritten by the host on a later step; the asynchronous copy out of it must have run by
        then (the CPU may run more than one step ahead of the GPU)."""
        if torch.cuda.is_available() and not torch.cuda.is_current_stream_capturing():
            ev = st.get(key)
            if ev is None:
                ev = st[key] = torch.cuda.Event()
            ev.record()

    @torch.no_grad()
    def step(self, closure=None, max_norm=None):
        from . import _lib
        import ctypes
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            if not self._fusable(group):
                if max_norm is not None:
                    torch.nn.utils.clip_grad_norm_([p for p in group["params"] if p.grad is not None], max_norm)
                saved, self.param_groups = self.param_groups, [group]
                try:
                    super().step()
                finally:
                    self.param_groups = saved
                continue
            ps = [p for p in group["params"] if p.grad is not None]
            dev = ps[0].device
            st = self._group_state(gi, group, dev)
            for p in ps:                                   # torch.optim.Adam's state layout
                s = self.state[p]
                if len(s) == 0:
                    s["step"] = torch.zeros((), dtype=torch.float32, device=dev)
                    s["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p in ps)
            if key != st["key"]:
                if st["table"] is None or st["table"].shape[0] != len(ps):
                    # allocated on the first (eager) step; a later capture only rewrites the pinned rows and re-issues the copy
                    st["table_host"] = torch.zeros(len(ps), 6, dtype=torch.int64).pin_memory()
                    st["table"] = torch.empty(len(ps), 6, dtype=torch.int64, device=dev)
                self._wait_staging(st, "table_ev")
                rows, chunk = st["table_host"], 0
                for i, p in enumerate(ps):
                    s = self.state[p]
                    rows[i, 0], rows[i, 1], rows[i, 2], rows[i, 3] = p.data_ptr(), p.grad.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr()
                    rows[i, 4], rows[i, 5] = p.numel(), chunk
                    chunk += (p.numel() + self.CHUNK - 1) // self.CHUNK
                st["table"].copy_(rows, non_blocking=True)
                self._mark_staging(st, "table_ev")
                st["key"], st["n_chunks"] = key, chunk
                if st["step"] is None:                      # resume: the device counter starts from the checkpointed step
                    st["hyper"][6:7].fill_(float(self.state[ps[0]]["step"]))
                    st["step"] = True
            if not capturing:
                self.sync_hyperparams()
            st["hyper"][5:6].fill_(float(max_norm) if max_norm is not None else 0.0)
            st["sumsq"].zero_()
            n = len(ps)
            _lib.call("fsnet_grad_sumsq", st["table"], n, ctypes.c_longlong(st["n_chunks"]), st["sumsq"])
            _lib.call("fsnet_adam_step", st["table"], n, ctypes.c_longlong(st["n_chunks"]), st["sumsq"], st["hyper"])
            st["last_params"] = ps
        return loss

    def state_dict(self):
        # materialise the shared device step counter into the per-parameter ``step`` entries torch expects
        for gi, st in self._fused.items():
            if st.get("last_params"):
                step = st["hyper"][6].float()
                for p in st["last_params"]:
                    self.state[p]["step"] = step.clone()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._fused = {}                                   # device counters / tables are rebuilt from the loaded state

    def total_norm(self, gi=0):
        """Gradient norm of the last step (what clip_grad_norm_ returns), as a 0-d device tensor."""
        return self._fused[gi]["sumsq"].sqrt().float().reshape(())


def build_optimizer(model: nn.Module, name, **kwargs):
    table = {"sgd": optim.SGD, "adam": FusedAdam, "adamw": optim.AdamW}
    if name.lower() not in table:
        raise NotImplementedError(name)
    return table[name.lower()](model.parameters(), **kwargs)


class PolyLR(optim.lr_scheduler._LRScheduler):
    def __init__(self, optimizer, gamma=0.9, n_iteration=-1):
        self.step_size, self.gamma = n_iteration, gamma
        super().__init__(optimizer)

    def get_lr(self):
        decay = max(0.0, 1 - self._step_count / float(self.step_size)) ** self.gamma
        return [base_lr * decay for base_lr in self.base_lrs]


def build_scheduler(optimizer, name=None, **kwargs):
    if name is None:
        return optim.lr_scheduler.ExponentialLR(optimizer, 1.0)
    table = {"steplr": optim.lr_scheduler.StepLR, "multisteplr": optim.lr_scheduler.MultiStepLR,
             "exponentiallr": optim.lr_scheduler.ExponentialLR, "cosineannealinglr": optim.lr_scheduler.CosineAnnealingLR,
             "polylr": PolyLR}
    if name.lower() not in table:
        raise NotImplementedError(name)
    return table[name.lower()](optimizer, **kwargs)


def _unwrap(model):
    return model.module if (torch.distributed.is_available() and torch.distributed.is_initialized() and hasattr(model, "module")) else model


def save_models(path, model, optimizer=None):
    """{'model_state_dict', 'optimizer_state_dict'} with the DDP wrapper stripped."""
    state = {"model_state_dict": _unwrap(model).state_dict()}
    if optimizer is not None:
        state["optimizer_state_dict"] = optimizer.state_dict()
    torch.save(state, path)


def load_models(path, model, optimizer=None, map_location=None, strict=False):
    state = torch.load(path, map_location=map_location)
    _unwrap(model).load_state_dict(state["model_state_dict"], strict=strict)
    if optimizer is not None and "optimizer_state_dict" in state:
        optimizer.load_state_dict(state["optimizer_state_dict"])

"""Whole-network executor of the tcgen05 path.

The reference runs its depth net / PoseNet as ~hundreds of autograd nodes (cuDNN conv, BN, ReLU, add,
interpolate, cat ...).  Here one ``torch.autograd.Function`` per network walks the module tree (which
only holds parameters, in the reference's layout), launches the hand-written kernels of
csrc/conv_tc.cu + csrc/act_tc.cu in forward order, and records a tape of backward closures that are
replayed in reverse.  Activations never leave the device-resident bf16 "planes" / fp32 "raw" buffers.

Per convolution: forward   conv (tcgen05, bf16x3, BN statistics in the epilogue) -> bn_finalize -> act_planes
                 backward  bn_bwd_reduce -> bn_bwd_apply (-> dy plane) -> conv_wgrad + conv (data gradient)
"""
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, peer, tc
from .tc import Fp32, Planes, View, pad16, pad_in

MOMENTUM_DEFAULT = 0.1


class Act:
    """An activation tensor of the graph: a channel slice of a planes buffer plus its gradient buffer.
    Slices of a concat buffer share the gradient buffer and its bookkeeping through ``root``."""

    def __init__(self, planes: Planes, c_off=0, c=None, relu=True, grad_ring=0, root: Optional["Act"] = None):
        self.planes, self.c_off, self.c = planes, c_off, planes.c if c is None else c
        self.relu = relu
        self.root = root if root is not None else self
        self._grad: Optional[Fp32] = None
        self._grad_ring = grad_ring
        self._written = False
        self._dirty = False
        self.zero_ring = False          # the ring holds materialised ZERO padding (network input for the 7x7 stem)

    n = property(lambda self: self.planes.n)
    h = property(lambda self: self.planes.h)
    w = property(lambda self: self.planes.w)
    grad = property(lambda self: self.root._grad)

    @property
    def grad_written(self):
        return self.root._written

    @grad_written.setter
    def grad_written(self, v):
        self.root._written = v

    @property
    def ring_dirty(self):
        return self.root._dirty

    @ring_dirty.setter
    def ring_dirty(self, v):
        self.root._dirty = v

    def pview(self) -> View:
        return self.planes.view(self.c_off, self.c)

    def ensure_grad(self):
        r = self.root
        if r._grad is None:
            r._grad = Fp32(r.n, r.h, r.w, r.planes.c, ring=r._grad_ring, device=r.planes.t.device)
        return r._grad

    def gview(self) -> View:
        return self.grad.view(self.c_off, self.c)

    def gview_full_ring(self) -> View:
        """The ringed gradient buffer seen as a plain [N, H+2, W+2, C] tensor (target of a pad-(k-1) data gradient)."""
        g = self.grad
        assert g.ring == 1 and self.c_off == 0 and self.c == g.c
        return View(g.t.data_ptr(), g.n, g.h + 2, g.w + 2, g.c, 0, g.c, 0)


class LayerState:
    """Per-convolution device state that persists across steps (operand planes, statistics, scale/shift)."""

    def __init__(self, conv: nn.Conv2d, bn: Optional[nn.Module], need_dgrad=True):
        self.conv, self.bn = conv, bn
        w = conv.weight
        # the 7x7 / stride-2 stem reads image planes padded to 8 channels (tc.pad_in), every other layer multiples of 16
        stem = tuple(conv.kernel_size) == (7, 7) and tuple(conv.stride) == (2, 2) and tuple(conv.padding) == (3, 3) and w.shape[1] <= 8
        self.w = tc.ConvWeights(w, need_dgrad, ci_pad=pad_in(w.shape[1]) if stem else None)
        C = self.w.co_pad
        dev = w.device
        self.C, self.c_real = C, w.shape[0]
        self.stats = torch.zeros(2 * C, device=dev, dtype=torch.float64)
        self.kh, self.kw = conv.kernel_size
        self.stride = conv.stride[0]
        self.pad = conv.padding[0]
        self.replicate = getattr(conv, "padding_mode", "zeros") == "replicate"
        assert conv.dilation[0] == 1 and conv.groups == 1, "tcgen05 path: dilation/groups unsupported"

    def padded(self, t: Optional[torch.Tensor], fill=0.0):
        if t is None:
            return None
        if t.shape[0] == self.C:
            return t.detach()
        out = torch.full((self.C,), fill, device=t.device, dtype=torch.float32)
        out[: t.shape[0]] = t.detach()
        return out


class Inv:
    """Per-invocation results of one layer (a network may run several times per step, e.g. the PoseNet):
    BN scale/shift, mean/invstd and the element count; the backward of THAT invocation reads them."""

    def __init__(self, st: LayerState):
        dev = st.stats.device
        self.st = st
        self.ss = torch.empty(2 * st.C, device=dev, dtype=torch.float32)
        self.mi = torch.empty(2 * st.C, device=dev, dtype=torch.float32)
        self.count = 1.0


_DIAG_NO_SYNCBN = os.environ.get("FSNET_DIAG_NO_SYNCBN", "0") == "1"       # timing diagnostics only: local statistics (wrong numerics at N>1)


def _sync_world(bn) -> int:
    if _DIAG_NO_SYNCBN:
        return 1
    if isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


REDUCED_STORAGES = set()      # storages whose gradients were already averaged over the ranks during backward (see Tape._reduce_bucket)


class Tape:
    """Forward executor + backward tape for one network invocation."""
    _wgrad_tables: Dict = {}
    # Data-parallel gradient exchange overlapped with the backward pass (switched on by BaseTrainingHook for models that are not
    # wrapped in DistributedDataParallel): the pooled weight-gradient accumulators are averaged over the ranks in a few buckets,
    # each launched (NCCL, asynchronous) as soon as the backward walk has finished the bucket's layers -- the reference gets the
    # same overlap from DDP's reducer (scripts/train.py:102).  The batched re-layout into parameter gradients then reads reduced
    # accumulators, so the bulk of the gradient bytes never goes through the hook's post-backward exchange.
    bucketed_allreduce = False
    N_BUCKETS = 4
    # Weight gradients on a second stream.  In the backward walk of a layer, dy feeds two independent launches: the data gradient
    # (which the rest of the walk waits for) and the weight gradient (needed only by the batched re-layout at the very end).  The
    # 34 weight-gradient launches of a cfg2a step (1.1 ms, most of them small grids with K-split tails) run on a side stream
    # that forks after dy is written and joins before the re-layout (and before a gradient bucket is exchanged), so they fill the
    # SMs the main chain leaves idle; inside the step graph the fork/join become parallel branches.  FSNET_WGRAD_STREAM=0: off.
    side_wgrad = os.environ.get("FSNET_WGRAD_STREAM", "1") != "0"
    _side_streams: Dict = {}

    def __init__(self, states: Dict[int, LayerState], training: bool, need_grad: bool, weights_fresh=False):
        self.states, self.training, self.need_grad = states, training, need_grad
        self.backward_ops: List = []
        self.param_grads: Dict[int, torch.Tensor] = {}      # id(parameter) -> gradient tensor
        self.weights_fresh = weights_fresh                    # operand planes already refreshed by one batched launch
        self._pools = None
        self._sync_scaled: List[torch.Tensor] = []            # SyncBN affine gradients to divide by the world size (see _bn_bwd)
        self._sync_world = 1
        self._side = None                                      # side stream with weight gradients in flight (None: joined)
        self._side_keep: List = []                             # operands of those launches: not handed back to the allocator before the join

    N_SIDE = int(os.environ.get("FSNET_WGRAD_STREAMS", "1"))     # side streams used round-robin

    def _wgrad_side_stream(self, dev):
        pool = Tape._side_streams.get(dev)
        if pool is None:
            pool = Tape._side_streams[dev] = [torch.cuda.Stream(device=dev) for _ in range(max(Tape.N_SIDE, 1))]
        self._side_turn = (getattr(self, "_side_turn", -1) + 1) % len(pool)
        return pool[self._side_turn]

    def _join_side(self):
        if self._side is not None:
            for st in Tape._side_streams.get(self._side.device, []):
                torch.cuda.current_stream().wait_stream(st)
            self._side = None
            self._side_keep.clear()

    def _make_pools(self):
        """One zeroed workspace per backward pass instead of one fill kernel per layer: BN-backward sums (fp64),
        wgrad accumulators (fp32) and the all-zero gradients of BN-cancelled conv biases."""
        dev = next(iter(self.states.values())).stats.device
        n_sums = sum(2 * st.C for st in self.states.values())
        n_acc = sum(st.w.acc_numel for st in self.states.values())
        n_zero = sum(st.c_real for st in self.states.values())
        n_w = sum(st.conv.weight.numel() for st in self.states.values())
        # BatchNorm affine and convolution bias gradients live behind the weight gradients in the SAME flat buffer: with every
        # parameter gradient a view of one storage, the data-parallel exchange is ONE in-place all-reduce (the ~100 small tensors
        # used to go through cat + all-reduce + a copy per tensor: 0.45 ms of a 6.5 ms step at 2 GPUs)
        n_small = sum(3 * st.c_real for st in self.states.values())
        self._pools = dict(sums=torch.zeros(n_sums, device=dev, dtype=torch.float64), acc=torch.zeros(n_acc, device=dev, dtype=torch.float32),
                           zero=torch.zeros(n_zero, device=dev, dtype=torch.float32), off={},
                           gw=torch.empty(n_w + n_small, device=dev, dtype=torch.float32), gw_off={}, gw_used=False, small_off={})
        o_s = o_a = o_z = o_w = 0
        o_small = n_w
        descs = []
        for key, st in self.states.items():
            self._pools["small_off"][key] = o_small          # [dgamma | dbeta | conv bias], c_real each
            o_small += 3 * st.c_real
            self._pools["off"][key] = (o_s, o_a, o_z)
            self._pools["gw_off"][key] = o_w
            d = _lib.WgradDesc()
            d.acc_off, d.grad_off = o_a, o_w
            d.cout, d.cin, d.kh, d.kw, d.cout_pad, d.cin_pad = st.w.co, st.w.ci, st.kh, st.kw, st.w.co_pad, st.w.ci_pad
            descs.append(d)
            o_s += 2 * st.C; o_a += st.w.acc_numel; o_z += st.c_real; o_w += st.conv.weight.numel()
        # the offsets only depend on the layer list: one device table per network, built on the first (eager) backward
        sig = (id(self.states), len(self.states), n_acc, n_w)
        if Tape._wgrad_tables.get("sig:%d" % id(self.states)) != sig:
            raw = bytes((_lib.WgradDesc * len(descs))(*descs))
            Tape._wgrad_tables[id(self.states)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            Tape._wgrad_tables["sig:%d" % id(self.states)] = sig
        self._pools["gw_table"] = Tape._wgrad_tables[id(self.states)]
        # gradient buckets: contiguous layer ranges of the accumulator pool, keyed by their FIRST layer (the last one the backward
        # walk reaches); boundaries only depend on the layer list, so every rank cuts the same buckets
        self._buckets, self._works = {}, []
        if Tape.bucketed_allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            keys = list(self.states.keys())
            sizes = [self.states[k].w.acc_numel for k in keys]
            target, start, run, lo = max(n_acc // Tape.N_BUCKETS, 1), 0, 0, 0
            for i, k in enumerate(keys):
                run += sizes[i]
                if run >= target or i == len(keys) - 1:
                    self._buckets[keys[start]] = (lo, lo + run)
                    lo, run, start = lo + run, 0, i + 1

    def _small_grad(self, st: LayerState, which: int) -> torch.Tensor:
        """Slice of the flat gradient buffer for a layer's small parameter gradients: 0 BatchNorm weight, 1 BatchNorm bias, 2 conv bias."""
        if self._pools is None:
            self._make_pools()
        o = self._pools["small_off"][id(st.conv)] + which * st.c_real
        self._pools["gw_used"] = True
        return self._pools["gw"][o:o + st.c_real]

    def _pool(self, st: LayerState, which: str):
        if self._pools is None:
            self._make_pools()
        o_s, o_a, o_z = self._pools["off"][id(st.conv)]
        if which == "sums":
            return self._pools["sums"][o_s:o_s + 2 * st.C]
        if which == "acc":
            return self._pools["acc"][o_a:o_a + st.w.acc_numel]
        return self._pools["zero"][o_z:o_z + st.c_real]

    def state(self, conv, bn, need_dgrad=True) -> LayerState:
        st = self.states.get(id(conv))
        if st is None:
            st = LayerState(conv, bn, need_dgrad)
            self.states[id(conv)] = st
        return st

    # -------------------------------------------------------------------------------------------
    def conv_bn_act(self, x: Act, conv: nn.Conv2d, bn, relu=True, residual: Optional[Act] = None, down=None,
                    up=1, dst: Optional[Act] = None, need_dgrad=True, grad_ring=0) -> Act:
        """conv -> BatchNorm (batch statistics when training) -> (+ residual | + BN(down conv)) -> ReLU -> planes.
        ``down`` = (conv, bn) of the 1x1 down-sample branch applied to ``residual``'s source ``x_down``."""
        st = self.state(conv, bn, need_dgrad)
        if not self.weights_fresh:
            st.w.refresh(conv.weight)
        N = x.n
        Ho = (x.h + 2 * st.pad - st.kh) // st.stride + 1
        Wo = (x.w + 2 * st.pad - st.kw) // st.stride + 1
        raw = Fp32(N, Ho, Wo, st.C, device=x.planes.t.device)
        bn_train = bn is not None and (self.training and bn.training)
        use_ring = st.replicate or (x.zero_ring and x.planes.ring == st.pad)     # zero ring == zero padding, read as data
        tc.conv(x.planes, st.w, raw, st.stride, st.pad, use_ring=use_ring, stats=st.stats if bn_train else None, in_view=x.pview())
        count = float(N * Ho * Wo)
        inv = Inv(st)
        self._finalize(inv, bn, bn_train, count)
        res_mode, res_view, res_ss = 0, None, None
        down_state = None
        if down is not None:
            dconv, dbn, x_down = down
            down_state = self._down_forward(x_down, dconv, dbn)
            res_mode, res_view, res_ss = 2, down_state["raw"].view(), down_state["inv"].ss
        elif residual is not None:
            res_mode, res_view = 1, residual.pview()
        if dst is None:
            out_planes = Planes(N, Ho * up, Wo * up, st.C, ring=1, device=raw.t.device)
            out = Act(out_planes, relu=relu, grad_ring=grad_ring)
        else:
            out = dst
        _lib.call("fsnet_act_planes", raw.view(), inv.ss, res_mode, res_view, res_ss, int(relu), up, out.pview())
        if self.need_grad:
            if need_dgrad:
                x.ensure_grad()
            if residual is not None and down is None:
                residual.ensure_grad()
            if down_state is not None:
                down_state["x"].ensure_grad()
            self.backward_ops.append(lambda: self._conv_bn_act_bwd(x, inv, bn, bn_train, raw, out, relu, residual, down_state, up, need_dgrad))
        return out

    def _finalize(self, inv: Inv, bn, bn_train: bool, count: float):
        st = inv.st
        conv = st.conv
        inv.count = count
        if bn is None:
            # plain convolution with bias: scale = 1, shift = bias
            inv.ss[: st.C] = 1.0
            inv.ss[st.C:] = 0.0 if conv.bias is None else st.padded(conv.bias)
            return
        world = _sync_world(bn) if bn_train else 1
        px = None
        if world > 1:
            count = count * world
            px = peer.get()
            if px is None:
                dist.all_reduce(st.stats)
        inv.count = count
        if bn.momentum is None:
            raise NotImplementedError("BatchNorm(momentum=None) (cumulative moving average) is not implemented on the tcgen05 path; "
                                      "no reference configuration uses it")
        momentum = bn.momentum
        if px is not None:
            # SyncBatchNorm: the cross-rank sum of the statistics happens INSIDE the finalize kernel, over NVLink peer memory
            _lib.call("fsnet_bn_finalize_sync", st.stats, tc.c_double(count), st.padded(bn.weight, 1.0), st.padded(bn.bias),
                      st.padded(conv.bias), self._buf(bn.running_mean, st), self._buf(bn.running_var, st), bn.num_batches_tracked,
                      float(momentum), float(bn.eps), st.C, inv.ss, inv.mi, px.slot((id(st), "fwd"), 2 * st.C))
            return
        _lib.call("fsnet_bn_finalize", st.stats, tc.c_double(count), st.padded(bn.weight, 1.0), st.padded(bn.bias),
                  st.padded(conv.bias), self._buf(bn.running_mean, st), self._buf(bn.running_var, st),
                  bn.num_batches_tracked if bn_train else None, float(momentum), float(bn.eps), int(bn_train), st.C,
                  inv.ss, inv.mi)

    @staticmethod
    def _buf(t, st):
        # running statistics are updated in place; channel counts of BN layers are always multiples of 16
        assert t is None or t.shape[0] == st.C, "BatchNorm channels must be a multiple of 16 on the tcgen05 path"
        return t

    def _down_forward(self, x: Act, conv, bn):
        st = self.state(conv, bn)
        if not self.weights_fresh:
            st.w.refresh(conv.weight)
        Ho = (x.h - 1) // st.stride + 1
        Wo = (x.w - 1) // st.stride + 1
        raw = Fp32(x.n, Ho, Wo, st.C, device=x.planes.t.device)
        bn_train = self.training and bn.training
        tc.conv(x.planes, st.w, raw, st.stride, 0, stats=st.stats if bn_train else None, in_view=x.pview())
        inv = Inv(st)
        self._finalize(inv, bn, bn_train, float(x.n * Ho * Wo))
        return dict(st=st, inv=inv, raw=raw, x=x, bn=bn, bn_train=bn_train)

    # -------------------------------------------------------------------------------------------
    def _bn_bwd(self, inv: Inv, bn, bn_train, g_view: View, up, mask_view, mask_ss, raw: Fp32, res_mode=0, res_view=None):
        """(ReLU o BatchNorm) backward -> dy plane; fills the BN / bias parameter gradients."""
        st = inv.st
        sums = self._pool(st, "sums")                      # zeroed slice of the pass-wide workspace
        has_bn = bn is not None and bn_train
        # BatchNorm in eval mode inside a training step (ResNet(norm_eval=True) / frozen stages): the same two kernels with the
        # RUNNING statistics as (mean, invstd) and an infinite count, which removes the two batch-statistics terms of the input
        # gradient -- what is left is the fixed per-channel scale gamma * invstd
        eval_bn = bn is not None and not bn_train
        mi = inv.mi if bn is not None else None
        _lib.call("fsnet_bn_bwd_reduce", g_view, up, mask_view, mask_ss, raw.view(), mi, sums)
        world = _sync_world(bn) if has_bn else 1
        if world > 1:
            px = peer.get()
            if px is not None:
                _lib.call("fsnet_peer_allreduce_f64", sums, 2 * st.C, px.slot((id(st), "bwd"), 2 * st.C))
            else:
                dist.all_reduce(sums)
        fold_dgrad = st.replicate and st.kh == 3 and st.C < 64          # see fsnet_conv: folded x-taps need ring == pad
        dy = Planes(raw.n, raw.h, raw.w, st.C, ring=2 if fold_dgrad else 0, device=raw.t.device)     # fsnet_bn_bwd_apply zeroes the ring
        gamma = st.padded(bn.weight, 1.0) if bn is not None else None
        C = st.c_real
        dgamma = dbeta = None
        dev = raw.t.device
        if bn is not None:
            if bn.weight is not None and bn.weight.requires_grad:
                dgamma, dbeta = self._small_grad(st, 0), self._small_grad(st, 1)
                self.param_grads[id(bn.weight)] = dgamma
                self.param_grads[id(bn.bias)] = dbeta
            if st.conv.bias is not None and st.conv.bias.requires_grad:
                if eval_bn:      # not cancelled by running statistics: d bias = gamma * invstd * sum(g)
                    self.param_grads[id(st.conv.bias)] = (gamma[:C] * mi[st.C:st.C + C] * sums[:C]).float()
                else:
                    self.param_grads[id(st.conv.bias)] = self._pool(st, "zero")      # cancelled by the batch mean
        elif st.conv.bias is not None and st.conv.bias.requires_grad:
            dbeta = self._small_grad(st, 2)
            self.param_grads[id(st.conv.bias)] = dbeta
        _lib.call("fsnet_bn_bwd_apply", g_view, up, mask_view, mask_ss, raw.view(), mi, gamma, sums,
                  tc.c_double(float("inf") if eval_bn else inv.count), dy.view(), res_mode, res_view, dgamma, dbeta, C)
        if world > 1 and dgamma is not None:
            # SyncBN: the kernel wrote the affine gradients from the ALL-REDUCED sums, i.e. already summed over the ranks, and the
            # data-parallel gradient averaging that follows (hook / DDP) would count them `world` times.  torch's SyncBatchNorm
            # keeps these two local; (global sum) / world is the same thing once the ranks are averaged.
            self._sync_scaled += [dgamma, dbeta]
            self._sync_world = world
        return dy

    def _conv_bwd(self, x: Act, st: LayerState, dy: Planes, need_dgrad=True):
        """Weight gradient and (accumulated) data gradient of one convolution.  Both read dy; the data gradient is issued first on
        the main stream, the weight gradient on the side stream behind an event recorded BEFORE it: the block scheduler sees the
        critical-path launch first, the weight gradient fills SMs as they free up."""
        conv = st.conv
        do_w = conv.weight.requires_grad
        side = self._wgrad_side_stream(dy.t.device) if (do_w and Tape.side_wgrad and dy.t.device.type == "cuda") else None
        acc = self._pool(st, "acc") if do_w else None            # (the zeroed workspace is created on the main stream, before the fork)
        use_ring = st.replicate or (x.zero_ring and x.planes.ring == st.pad)
        forked = None
        if side is not None:
            forked = torch.cuda.Event()
            forked.record()                                      # dy is written
        elif do_w:
            tc.conv_wgrad(x.pview(), use_ring, dy.view(), st.w, st.stride, st.pad, acc=acc)
        if need_dgrad and x.grad is not None:
            self._conv_dgrad(x, st, dy)
        if side is not None:
            side.wait_event(forked)
            with torch.cuda.stream(side):
                tc.conv_wgrad(x.pview(), use_ring, dy.view(), st.w, st.stride, st.pad, acc=acc)
            self._side = side
            self._side_keep.append((dy, x.planes))
        if do_w:
            # the accumulator is re-laid out into this slice of the flat gradient buffer by ONE batched launch at the end
            # of the backward pass (run_backward)
            o_w = self._pools["gw_off"][id(st.conv)]
            self.param_grads[id(conv.weight)] = self._pools["gw"][o_w:o_w + conv.weight.numel()].view_as(conv.weight)
            self._pools["gw_used"] = True
        self._reduce_bucket(id(conv))

    def _conv_dgrad(self, x: Act, st: LayerState, dy: Planes):
        k = st.kh
        if st.replicate:
            # x_pad = replicate_pad(x): gradient of the padded tensor (pad k-1 correlation), then fold the ring
            assert x.grad.ring == 1, "replicate consumers need a ringed gradient buffer"
            tc.conv_dgrad(dy, st.w, x.gview_full_ring(), pad=k - 1, accumulate=x.grad_written, use_ring=(dy.ring == k - 1))
            x.grad_written = True
            x.ring_dirty = True
            return
        if x.grad.ring == 1 and not x.grad_written:
            x.grad.t.zero_()                       # interior-only writer on a ringed buffer: the ring must read as zero
            x.grad_written = True
        src = dy
        if st.stride == 2:
            src = Planes(x.n, x.h, x.w, st.C, ring=0, device=dy.t.device)
            _lib.call("fsnet_zero_insert", dy.view(), src.view())
        tc.conv_dgrad(src, st.w, x.gview(), pad=k - 1 - st.pad, accumulate=x.grad_written)
        x.grad_written = True

    def _reduce_bucket(self, key):
        """The backward walk has finished layer `key`: if it is the first layer of a gradient bucket, every accumulator of the
        bucket is final -- average it over the ranks now, concurrently with the rest of the backward pass."""
        rng = self._buckets.pop(key, None) if self._pools is not None else None
        if rng is not None:
            op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
            t = self._pools["acc"][rng[0]:rng[1]]
            if self._side is not None:
                # the bucket's accumulators are final once the side stream's weight-gradient launches have run: the collective is
                # ordered behind THEM (issued from the side stream), the main chain does not wait
                for st in Tape._side_streams.get(self._side.device, []):
                    if st is not self._side:
                        self._side.wait_stream(st)
                with torch.cuda.stream(self._side):
                    work = dist.all_reduce(t, op=op, async_op=True)
            else:
                work = dist.all_reduce(t, op=op, async_op=True)
            self._works.append((work, t, op))

    def _fold_if_needed(self, a: Act):
        if getattr(a, "ring_dirty", False):
            _lib.call("fsnet_fold_ring", a.grad.view())
            a.ring_dirty = False

    def _conv_bn_act_bwd(self, x: Act, inv: Inv, bn, bn_train, raw, out: Act, relu, residual, down_state, up, need_dgrad=True):
        st = inv.st
        if out.grad is None or not out.grad_written:
            return                                  # nothing downstream needs this activation's gradient
        self._fold_if_needed(out)
        g_view = out.gview()
        mask_view = out.pview() if (relu and up == 1) else None
        mask_ss = inv.ss if (relu and up == 2) else None
        if relu and up == 1 and bn is not None and residual is None and down_state is None:
            # no residual: the ReLU mask is sign(raw*scale+shift), recomputed from the raw output the kernels read anyway
            # (saves the 4 bytes/element read of the activation planes in both BatchNorm-backward passes)
            mask_view, mask_ss = None, inv.ss
        res_mode, res_view = 0, None
        if residual is not None and down_state is None and residual.grad is not None:
            if residual.grad.ring == 1 and not residual.grad_written:
                residual.grad.t.zero_()
            res_mode = 2 if residual.grad_written else 1
            res_view = residual.gview()
            residual.grad_written = True
        dy = self._bn_bwd(inv, bn, bn_train, g_view, up, mask_view, mask_ss, raw, res_mode, res_view)
        if down_state is not None:
            ds = down_state
            dyd = self._bn_bwd(ds["inv"], ds["bn"], ds["bn_train"], g_view, 1, out.pview() if relu else None, None, ds["raw"])
            self._conv_bwd(ds["x"], ds["st"], dyd)
        self._conv_bwd(x, st, dy, need_dgrad)

    # -------------------------------------------------------------------------------------------
    def maxpool(self, x: Act) -> Act:
        out = Act(Planes(x.n, (x.h + 1) // 2, (x.w + 1) // 2, x.c, ring=1, device=x.planes.t.device))
        argmax = torch.empty(out.n, out.h, out.w, x.c, device=x.planes.t.device, dtype=torch.uint8) if self.need_grad else None
        _lib.call("fsnet_maxpool_planes", x.pview(), out.pview(), argmax)
        if self.need_grad:
            x.ensure_grad()

            def bwd():
                if out.grad is None or not out.grad_written or x.grad is None:
                    return
                _lib.call("fsnet_maxpool_bwd", x.pview(), argmax, out.gview(), x.gview(), int(x.grad_written))
                x.grad_written = True
            self.backward_ops.append(bwd)
        return out

    def copy_skip(self, skip: Act, dst: Act):
        _lib.call("fsnet_copy_planes", skip.pview(), dst.pview())
        if self.need_grad:
            skip.ensure_grad()

            def bwd():
                if dst.grad is None or not dst.grad_written or skip.grad is None:
                    return
                self._fold_if_needed(dst)
                _lib.call("fsnet_add_slice", skip.gview(), dst.gview(), int(skip.grad_written))
                skip.grad_written = True
            self.backward_ops.append(bwd)

    def conv_out(self, x: Act, conv: nn.Conv2d, key) -> Fp32:
        """Plain convolution + bias whose fp32 result leaves the tcgen05 graph (dispconv logits, pose output).
        Its incoming gradient is looked up as ``self.out_grads[key]`` (fp32 NHWC, padded channels) at backward time."""
        st = self.state(conv, None)
        if not self.weights_fresh:
            st.w.refresh(conv.weight)
        Ho = (x.h + 2 * st.pad - st.kh) // st.stride + 1
        Wo = (x.w + 2 * st.pad - st.kw) // st.stride + 1
        out = Fp32(x.n, Ho, Wo, st.C, device=x.planes.t.device)
        tc.conv(x.planes, st.w, out, st.stride, st.pad, use_ring=st.replicate, bias=st.padded(conv.bias), in_view=x.pview())
        if self.need_grad:
            x.ensure_grad()

            def bwd():
                g_out = self.out_grads.get(key)
                if g_out is None:
                    return
                g = Fp32.wrap(g_out)
                dy = self._bn_bwd(Inv(st), None, False, g.view(), 1, None, None, out)
                self._conv_bwd(x, st, dy)
            self.backward_ops.append(bwd)
        return out

    def run_backward(self, out_grads: Dict):
        self.out_grads = out_grads
        for op in reversed(self.backward_ops):
            op()
        self._join_side()
        if self._pools is not None and self._pools["gw_used"]:
            if self._buckets or self._works:
                for key in list(self._buckets):              # buckets whose first layer has no weight gradient (frozen stages)
                    self._reduce_bucket(key)
                for work, t, op in self._works:
                    work.wait()
                    if op == dist.ReduceOp.SUM:
                        t.div_(dist.get_world_size())
                # (bucketed mode: the weights were averaged as accumulators; the small gradients behind them in the flat buffer still
                # need their exchange)
                n_w = sum(st.conv.weight.numel() for st in self.states.values())
                tail = self._pools["gw"][n_w:]
                if tail.numel():
                    op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
                    dist.all_reduce(tail, op=op)
                    if op == dist.ReduceOp.SUM:
                        tail.div_(dist.get_world_size())
                REDUCED_STORAGES.add(self._pools["gw"].untyped_storage().data_ptr())
            _lib.call("fsnet_wgrad_to_param_batched", self._pools["gw_table"], len(self.states), self._pools["acc"], self._pools["gw"])
        if self._pools is not None:
            REDUCED_STORAGES.add(self._pools["zero"].untyped_storage().data_ptr())      # exact zeros on every rank: nothing to exchange
        if self._sync_scaled:
            torch._foreach_div_(self._sync_scaled, float(self._sync_world))
            self._sync_scaled = []

    def take_param_grads(self, params):
        """The parameter gradients of this pass, handed to autograd WITHOUT keeping a reference: AccumulateGrad adopts a gradient
        (p.grad becomes the view of the flat buffer itself) only when nobody else holds the tensor; with the tape's dict still
        pointing at them every gradient was cloned -- 108 device-to-device copies (45 MB) per step that no kernel list shows, and
        p.grad tensors with private storages, which sent the data-parallel exchange through cat + all-reduce + 108 copies."""
        out = tuple(self.param_grads.pop(id(p), None) for p in params)
        self.param_grads.clear()
        return out


# ---------------------------------------------------------------------------------------------------
# network walkers
# ---------------------------------------------------------------------------------------------------
def _image_act(img: torch.Tensor) -> Act:
    n, c, h, w = img.shape
    # ring 3 = the stem's zero padding, materialised (the 7x7/2 stem then runs on the folded-tap convolution path)
    p = Planes(n, h, w, pad_in(c), ring=3, device=img.device)
    _lib.call("fsnet_image_to_planes_ring", img.detach().float().contiguous(), c, p.view(), 1)
    a = Act(p, relu=False)
    a.zero_ring = True
    return a


def resnet_forward(tape: Tape, net, img: torch.Tensor, want_grads: bool) -> List[Act]:
    """ResNet.forward (fsnet_b200/networks/resnet.py) on the tcgen05 path -> the five feature Acts."""
    from .networks.resnet import BasicBlock
    x = _image_act(img)
    feats = []
    a = tape.conv_bn_act(x, net.conv1, net.bn1, relu=True, need_dgrad=False)
    feats.append(a)
    a = tape.maxpool(a)
    for i in range(net.num_stages):
        for block in getattr(net, f"layer{i + 1}"):
            inp = a
            if isinstance(block, BasicBlock):
                o = tape.conv_bn_act(inp, block.conv1, block.bn1, relu=True)
                if block.downsample is not None:
                    a = tape.conv_bn_act(o, block.conv2, block.bn2, relu=True, down=(block.downsample[0], block.downsample[1], inp))
                else:
                    a = tape.conv_bn_act(o, block.conv2, block.bn2, relu=True, residual=inp)
            else:
                o = tape.conv_bn_act(inp, block.conv1, block.bn1, relu=True)
                o = tape.conv_bn_act(o, block.conv2, block.bn2, relu=True)
                if block.downsample is not None:
                    a = tape.conv_bn_act(o, block.conv3, block.bn3, relu=True, down=(block.downsample[0], block.downsample[1], inp))
                else:
                    a = tape.conv_bn_act(o, block.conv3, block.bn3, relu=True, residual=inp)
        feats.append(a)
    return feats


def decoder_forward(tape: Tape, dec, feats: List[Act]) -> Dict[int, Fp32]:
    """DepthDecoder._trunk on the tcgen05 path -> {scale: logits (fp32 NHWC buffer, padded channels)}."""
    logits = {}
    x = feats[-1]
    dev = x.planes.t.device
    for i in range(4, -1, -1):
        up0 = dec.convs[("upconv", i, 0)]
        up1 = dec.convs[("upconv", i, 1)]
        cd = int(dec.num_ch_dec[i])
        skip = feats[i - 1] if (dec.use_skips and i > 0) else None
        cs = skip.c if skip is not None else 0
        cat = Planes(x.n, x.h * 2, x.w * 2, cd + cs, ring=1, device=dev)
        cat_act = Act(cat, relu=False, grad_ring=1)
        up_part = Act(cat, 0, cd, relu=True, root=cat_act)
        tape.conv_bn_act(x, up0.sequence[0], up0.sequence[1], relu=True, up=2, dst=up_part)
        if skip is not None:
            tape.copy_skip(skip, Act(cat, cd, cs, relu=False, root=cat_act))
        x = tape.conv_bn_act(cat_act, up1.sequence[0], up1.sequence[1], relu=True, grad_ring=1)
        if i in dec.scales:
            logits[i] = tape.conv_out(x, dec.convs[("dispconv", i)], ("logits", i))
            if ("uncertain_logz", i) in dec.convs:          # MultiChannelDepthDecoderUncertain: second head on the same activation
                logits[("uncertain", i)] = tape.conv_out(x, dec.convs[("uncertain_logz", i)], ("uncertain", i))
    return logits


def pose_decoder_forward(tape: Tape, dec, last: Act) -> Fp32:
    """PoseDecoder convolutions (num_input_features = 1) -> fp32 NHWC [N, h, w, pad16(6*n_pred)]."""
    if dec.num_input_features != 1:
        raise NotImplementedError("tcgen05 PoseDecoder path handles num_input_features=1 (the only wiring the reference shows)")
    x = tape.conv_bn_act(last, dec.convs["squeeze"], None, relu=True)
    x = tape.conv_bn_act(x, dec.convs[("pose", 0)], None, relu=True)
    x = tape.conv_bn_act(x, dec.convs[("pose", 1)], None, relu=True)
    return tape.conv_out(x, dec.convs[("pose", 2)], "pose")


# ---------------------------------------------------------------------------------------------------
# autograd wrappers
# ---------------------------------------------------------------------------------------------------
def _nhwc_grad(g: torch.Tensor, c_pad: int) -> torch.Tensor:
    g = g.detach().float().permute(0, 2, 3, 1)
    if g.shape[-1] != c_pad:
        out = torch.zeros(*g.shape[:-1], c_pad, device=g.device, dtype=torch.float32)
        out[..., : g.shape[-1]] = g
        return out
    return g.contiguous()


def _check_supported(backbone):
    want = tuple([-1] + list(range(getattr(backbone, "num_stages", 4))))
    if tuple(getattr(backbone, "out_indices", want)) != want:
        raise NotImplementedError(f"tcgen05 path: out_indices must be {want} (stem + every stage, as in every reference config), "
                                  f"got {tuple(backbone.out_indices)}")
    if any(d != 1 for d in getattr(backbone, "dilations", (1,))):
        raise NotImplementedError("tcgen05 path: dilated convolutions are not implemented (no shipped config uses them)")


class _DepthNetFn(torch.autograd.Function):
    """ResNet encoder + U-Net decoder up to the per-scale logits, as ONE autograd node."""

    @staticmethod
    def forward(ctx, runner, img, *params):
        need_grad = runner.grad_enabled and any(p.requires_grad for p in params)
        tape = Tape(runner.states, runner.training, need_grad, runner.weights_fresh)
        feats = resnet_forward(tape, runner.backbone, img, need_grad)
        logits = decoder_forward(tape, runner.decoder, feats)
        ctx.tape, ctx.params, ctx.scales = tape, params, list(logits.keys())
        ctx.c_pad = {s: logits[s].c for s in logits}
        n = runner.decoder.num_output_channels
        outs = []
        for s in ctx.scales:
            t = logits[s].t.permute(0, 3, 1, 2)         # NCHW view of the NHWC buffer (channels_last strides)
            n_s = 1 if isinstance(s, tuple) else n      # ("uncertain", scale): one-channel uncertainty head
            if n_s != logits[s].c:
                t = t[:, :n_s].contiguous(memory_format=torch.channels_last)
            outs.append(t)
        runner.last_features = feats
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_logits):
        tape = ctx.tape
        grads = {(s if isinstance(s, tuple) else ("logits", s)): _nhwc_grad(g, ctx.c_pad[s])
                 for s, g in zip(ctx.scales, g_logits) if g is not None}
        tape.run_backward(grads)
        return (None, None) + tape.take_param_grads(ctx.params)


class _PoseNetFn(torch.autograd.Function):
    """PoseNet: 6-channel ResNet + PoseDecoder convolutions, as ONE autograd node -> [N, C, h, w] pose map."""

    @staticmethod
    def forward(ctx, runner, img, *params):
        need_grad = runner.grad_enabled and any(p.requires_grad for p in params)
        tape = Tape(runner.states, runner.training, need_grad, runner.weights_fresh)
        feats = resnet_forward(tape, runner.backbone, img, need_grad)
        out = pose_decoder_forward(tape, runner.decoder, feats[-1])
        ctx.tape, ctx.params, ctx.c_pad = tape, params, out.c
        n = 6 * runner.decoder.num_frames_to_predict_for
        return out.t.permute(0, 3, 1, 2)[:, :n].contiguous()

    @staticmethod
    def backward(ctx, g):
        tape = ctx.tape
        tape.run_backward({"pose": _nhwc_grad(g, ctx.c_pad)})
        return (None, None) + tape.take_param_grads(ctx.params)


class Runner:
    """Binds a backbone and a decoder module to the persistent per-layer device state."""

    def __init__(self, backbone, decoder):
        self.backbone, self.decoder = backbone, decoder
        self.states: Dict[int, LayerState] = {}
        self.last_features = None

    def params(self):
        return [p for p in list(self.backbone.parameters()) + list(self.decoder.parameters())]

    def _prep(self):
        _check_supported(self.backbone)
        self.training = self.backbone.training
        self.grad_enabled = torch.is_grad_enabled()
        self.weights_fresh = self._refresh_all()

    def _refresh_all(self) -> bool:
        """From the second invocation on (all LayerStates exist) every convolution's bf16 operand planes are rebuilt
        from the fp32 parameters by ONE launch over a device-resident descriptor table."""
        if not self.states:
            return False
        sig = tuple(st.conv.weight.data_ptr() for st in self.states.values())
        if getattr(self, "_table_sig", None) != sig:
            descs = (_lib.WeightDesc * len(self.states))(*[st.w.desc(st.conv.weight) for st in self.states.values()])
            raw = bytes(descs)
            dev = next(iter(self.states.values())).stats.device
            self._table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            self._table_sig = sig
        _lib.call("fsnet_weight_planes_batched", self._table, len(self.states))
        return True

    def depth_logits(self, img):
        self._prep()
        outs = _DepthNetFn.apply(self, img, *self.params())
        keys = []                                       # the insertion order of decoder_forward
        for s in range(4, -1, -1):
            if s in self.decoder.scales:
                keys.append(s)
                if ("uncertain_logz", s) in self.decoder.convs:
                    keys.append(("uncertain", s))
        return dict(zip(keys, outs))

    def pose_map(self, img):
        self._prep()
        return _PoseNetFn.apply(self, img, *self.params())

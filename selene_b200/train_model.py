"""Training step on the B200 kernels (reference: selene_sdk/train_model.py:36-686, of which the hot
part is the step of :469-508: sample -> forward (train mode) -> BCELoss -> zero_grad -> backward ->
optimizer.step with the SGD of models/deepsea.py:66-74).

``B200Trainer`` owns an ``sb_net`` in training mode and exposes ``step(inputs, targets)``;
``TrainModel`` keeps the reference constructor's shape for the arguments that matter to the step
(model, data_sampler, optimizer kwargs, batch_size, max_steps) and drives it.  Under
``torch.distributed`` (one process per GPU) gradients are averaged with one NCCL all-reduce of the flat
gradient vector before the update — the DDP scheme of SURVEY §8e.
"""
import ctypes as C
import logging
from time import time

import numpy as np

from . import _lib
from .nets import B200Net, layers_from_module

logger = logging.getLogger("selene")


class B200Trainer(object):
    def __init__(self, model_or_layers, sequence_length, lr=0.01, momentum=0.9, weight_decay=1e-6,
                 dropout=None, perm=(0, 1, 2, 3), device=None, seed=0, precision="fp32"):
        import torch
        layers = model_or_layers if isinstance(model_or_layers, list) else layers_from_module(model_or_layers)
        self.net = B200Net(layers, sequence_length, perm=perm, device=device)
        self._lib = self.net._lib
        self._h = self.net._h
        self.precision = precision
        if precision == "bf16":
            _lib.check(self._lib.sb_net_train_init_tc(self._h), "sb_net_train_init_tc")
        else:
            _lib.check(self._lib.sb_net_train_init(self._h), "sb_net_train_init")
        self.n_layers = int(self._lib.sb_net_n_layers(self._h))
        self.n_params = int(self._lib.sb_net_n_params(self._h))
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.device = torch.device("cuda", self.net.device)
        self.grads = torch.zeros((self.n_params,), dtype=torch.float32, device=self.device)
        self.loss = torch.zeros((1,), dtype=torch.float32, device=self.device)
        self._ws = None
        self._shapes = [(l["weight"].shape, l["bias"].shape) for l in layers]
        self.seed = int(seed)
        self.step_count = 0
        if dropout:
            for layer, p in dropout.items():
                _lib.check(self._lib.sb_net_set_dropout(self._h, int(layer), float(p)), "sb_net_set_dropout")
        # data parallel: the classifier's gradients (the tail of the flat vector, 89 % of DeepSEA's parameters) are
        # final early in the backward pass; an event recorded there lets their all-reduce overlap the conv backward
        self._tail_layer = next((i for i, l in enumerate(layers) if l["type"] == "fc"), None)
        self._tail_off = 0
        self._tail_event = None
        self._comm_stream = None

    def _enable_overlap(self):
        import torch
        if self._tail_event is None and self._tail_layer is not None and self._tail_layer > 0:
            self._tail_off = int(self._lib.sb_net_param_offset(self._h, self._tail_layer))
            self._tail_event = torch.cuda.Event()
            self._tail_event.record()                  # creates the underlying cudaEvent_t
            _lib.check(self._lib.sb_net_set_grad_event(self._h, self._tail_layer,
                                                       C.c_void_p(self._tail_event.cuda_event)), "sb_net_set_grad_event")
            self._comm_stream = torch.cuda.Stream(device=self.device)

    @classmethod
    def dropout_of_module(cls, model):
        """{layer index: p} from the module order conv/fc ... Dropout (the Dropout after a layer applies to
        that layer's output)."""
        import torch.nn as nn
        out, idx = {}, -1
        for m in model.modules():
            if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)):
                idx += 1
            elif isinstance(m, nn.Dropout) and idx >= 0 and m.p > 0:
                out[idx] = float(m.p)
        return out

    def _workspace(self, n):
        import torch
        fn = (self._lib.sb_net_train_workspace_bytes_tc if self.precision == "bf16"
              else self._lib.sb_net_train_workspace_bytes)
        need = int(fn(self._h, int(n)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        return self._ws

    def forward_backward(self, inputs, targets, masks=None):
        """inputs: CUDA fp32 (n, L, 4); targets: CUDA fp32 (n, F); masks: optional {layer: CUDA fp32 mask in
        the layer's channel-last output layout, already scaled}.  Fills ``self.grads`` / ``self.loss``."""
        inputs, targets = inputs.contiguous(), targets.contiguous()
        n = int(inputs.shape[0])
        ws = self._workspace(n)
        mask_arr = None
        if masks:
            mask_arr = (C.c_void_p * self.n_layers)()
            for k, t in masks.items():
                mask_arr[int(k)] = t.data_ptr()
        fb = self._lib.sb_net_forward_backward_tc if self.precision == "bf16" else self._lib.sb_net_forward_backward
        _lib.check(fb(
            self._h, _lib.ptr(inputs), _lib.ptr(targets), n, mask_arr, self.seed * 1000003 + self.step_count,
            _lib.ptr(ws), ws.numel(), _lib.ptr(self.grads), _lib.ptr(self.loss), _lib.stream_ptr()),
            "sb_net_forward_backward")

    def allreduce_grads(self, overlap=True):
        """Mean of the flat gradient vector over the ranks (NCCL).  With ``overlap`` the classifier tail is reduced
        on a side stream as soon as ``sb_net_forward_backward`` has produced it (the event of sb_net_set_grad_event),
        i.e. while the convolution gradients are still being computed; the head follows on the current stream."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        div = None if avg == dist.ReduceOp.AVG else dist.get_world_size()
        if overlap and self._tail_event is not None:
            cur = torch.cuda.current_stream()
            self._comm_stream.wait_event(self._tail_event)
            with torch.cuda.stream(self._comm_stream):
                tail = dist.all_reduce(self.grads[self._tail_off:], op=avg, async_op=True)
            dist.all_reduce(self.grads[:self._tail_off], op=avg)
            tail.wait()                                   # the current stream waits for the side-stream reduction
            cur.wait_stream(self._comm_stream)
        else:
            dist.all_reduce(self.grads, op=avg)
        if div:
            self.grads.div_(div)

    def sgd_step(self):
        _lib.check(self._lib.sb_net_sgd_step(self._h, _lib.ptr(self.grads), self.lr, self.momentum, self.weight_decay,
                                             _lib.stream_ptr()), "sb_net_sgd_step")
        self.step_count += 1

    def step_async(self, inputs, targets, masks=None):
        """Enqueues one optimisation step and returns without synchronising; read the loss with ``loss_value()``
        (before the next step overwrites it).  Lets the host draw the next batch while the GPU steps."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._enable_overlap()
        self.forward_backward(inputs, targets, masks)
        self.allreduce_grads()
        self.sgd_step()

    def loss_value(self):
        return float(self.loss.item())

    def step(self, inputs, targets, masks=None):
        """One optimisation step; returns the mean BCE loss as a python float (device -> host read)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._enable_overlap()
        self.forward_backward(inputs, targets, masks)
        self.allreduce_grads()
        self.sgd_step()
        return float(self.loss.item())

    def layer_params(self, layer, grads=False):
        wshape, bshape = self._shapes[layer]
        w = np.zeros(wshape, dtype=np.float32)
        b = np.zeros(bshape, dtype=np.float32)
        _lib.check(self._lib.sb_net_get_layer_params(self._h, int(layer), _lib.ptr(self.grads) if grads else None,
                                                     _lib.ptr(w), _lib.ptr(b)), "sb_net_get_layer_params")
        return w, b

    def export_layers(self, template_layers):
        """Layer list with the current weights (for a fresh inference ``B200Net`` or a checkpoint)."""
        out = []
        for i, l in enumerate(template_layers):
            w, b = self.layer_params(i)
            out.append(dict(l, weight=w, bias=b))
        return out


class TrainModel(object):
    """The step-driving subset of selene_sdk.TrainModel (train_model.py:36-508): ``train()`` performs one
    sample + optimisation step, ``train_and_validate()`` loops to ``max_steps`` logging the loss."""

    def __init__(self, model, data_sampler, loss_criterion=None, optimizer_class=None, optimizer_kwargs=None,
                 batch_size=64, max_steps=1000, report_stats_every_n_steps=100, sequence_length=1000,
                 use_cuda=True, device=None, precision="fp32", **_ignored):
        if not use_cuda:
            raise ValueError("selene_b200 has no CPU path: use_cuda must be True")
        kw = dict(optimizer_kwargs or {})
        self.sampler = data_sampler
        self.batch_size = batch_size
        self.max_steps = max_steps
        self.nth_step_report_stats = report_stats_every_n_steps
        self.trainer = B200Trainer(model, sequence_length, lr=kw.get("lr", 0.01), momentum=kw.get("momentum", 0.0),
                                   weight_decay=kw.get("weight_decay", 0.0),
                                   dropout=B200Trainer.dropout_of_module(model), device=device, precision=precision)
        self.step = 0
        self._time_per_step, self._train_loss = [], []

    def train(self):
        """train_model.py:469-508.  With ``self.prefetch`` (off by default: it draws batch i+1 before the loss of
        step i is read, i.e. one batch ahead of the reference's call order) the host-side sampling overlaps the
        GPU step."""
        t_i = time()
        if getattr(self, "prefetch", False):
            if getattr(self, "_next_batch", None) is None:
                self._next_batch = self.sampler.sample_device(self.batch_size, seq_dtype="float32", label_dtype="float32")
            inputs, targets = self._next_batch
            self.trainer.step_async(inputs, targets)
            self._next_batch = self.sampler.sample_device(self.batch_size, seq_dtype="float32", label_dtype="float32")
            t_sampling = time() - t_i
            loss = self.trainer.loss_value()
            self._train_loss.append(loss)
            self._time_per_step.append(time() - t_i)
            self.step += 1
            return loss, t_sampling
        inputs, targets = self.sampler.sample_device(self.batch_size, seq_dtype="float32", label_dtype="float32")
        t_sampling = time() - t_i
        loss = self.trainer.step(inputs, targets)
        self._train_loss.append(loss)
        self._time_per_step.append(time() - t_i)
        self.step += 1
        return loss, t_sampling

    def train_and_validate(self):
        for _ in range(self.step, self.max_steps):
            loss, _t = self.train()
            if self.step % self.nth_step_report_stats == 0:
                logger.info("[STEP {0}] average number of steps per second: {1:.1f}; training loss: {2}".format(
                    self.step, 1. / np.average(self._time_per_step), np.average(self._train_loss)))
                self._time_per_step, self._train_loss = [], []

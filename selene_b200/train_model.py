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
                 dropout=None, perm=(0, 1, 2, 3), device=None, seed=0, precision="fp32", grad_comm_dtype=None):
        import torch
        # data-parallel gradient exchange: fp32 (exact mean of the ranks' gradients) or bf16 (half the bytes; default of
        # the bf16 training path, whose operands are rounded to bf16 anyway)
        self.grad_comm_dtype = grad_comm_dtype or ("bf16" if precision == "bf16" else "fp32")
        if self.grad_comm_dtype not in ("fp32", "bf16"):
            raise ValueError("grad_comm_dtype must be 'fp32' or 'bf16', got %r" % (grad_comm_dtype,))
        self._grads16 = None
        self._tail_sgd_done = False
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
        from .nets import unwrap_non_strand_specific
        out, idx = {}, -1
        for m in unwrap_non_strand_specific(model)[0].modules():
            if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear, nn.LSTM)):     # the layer kinds of layers_from_module
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

    def allreduce_grads(self, overlap=True, sgd_tail=False):
        """Mean of the flat gradient vector over the ranks (NCCL).  With ``overlap`` the classifier tail is reduced
        on a side stream as soon as ``sb_net_forward_backward`` has produced it (the event of sb_net_set_grad_event),
        i.e. while the convolution gradients are still being computed; the head follows on the current stream.
        ``sgd_tail`` (what ``step`` / ``step_async`` pass): the classifier's optimiser step is enqueued right behind its
        all-reduce on the side stream too, and the following ``sgd_step()`` only updates the remaining layers."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        div = None if avg == dist.ReduceOp.AVG else dist.get_world_size()
        half = self.grad_comm_dtype == "bf16"
        if half and self._grads16 is None:
            self._grads16 = torch.empty((self.n_params,), dtype=torch.bfloat16, device=self.device)

        def reduce(lo, hi, **kw):
            """all-reduce of grads[lo:hi] on the current stream, through the bf16 copy if configured"""
            if not half:
                return dist.all_reduce(self.grads[lo:hi], op=avg, **kw)
            # slices start at multiples of 4 elements or at a cudaMalloc'd base: the casts' alignment holds
            _lib.check(self._lib.sb_cast_f32_to_bf16(_lib.ptr(self.grads[lo:hi]), hi - lo, _lib.ptr(self._grads16[lo:hi]),
                                                     _lib.stream_ptr()), "sb_cast_f32_to_bf16")
            return dist.all_reduce(self._grads16[lo:hi], op=avg, **kw)

        def widen(lo, hi):
            if half:
                _lib.check(self._lib.sb_cast_bf16_to_f32(_lib.ptr(self._grads16[lo:hi]), hi - lo, _lib.ptr(self.grads[lo:hi]),
                                                         _lib.stream_ptr()), "sb_cast_bf16_to_f32")

        n = self.n_params
        if overlap and self._tail_event is not None and self._tail_off % 4 == 0:
            cur = torch.cuda.current_stream()
            self._comm_stream.wait_event(self._tail_event)
            with torch.cuda.stream(self._comm_stream):
                tail = reduce(self._tail_off, n, async_op=True)
                tail.wait()
                if sgd_tail:
                    # the classifier's optimiser step (89 % of the parameters) right behind its all-reduce, still on
                    # the side stream: it overlaps the convolution backward and the head's all-reduce.  With bf16
                    # communication it reads the reduced bf16 vector directly (self.grads keeps this rank's own tail).
                    _lib.check(self._lib.sb_net_sgd_step_layers(self._h, _lib.ptr(self._grads16 if half else self.grads),
                                                                1 if half else 0, self.lr, self.momentum, self.weight_decay,
                                                                self._tail_layer, self.n_layers, 0, _lib.stream_ptr()),
                               "sb_net_sgd_step_layers")
                    self._tail_sgd_done = True
                else:
                    widen(self._tail_off, n)
            reduce(0, self._tail_off)
            widen(0, self._tail_off)
            cur.wait_stream(self._comm_stream)            # the current stream waits for the side-stream reduction
        else:
            reduce(0, n)
            widen(0, n)
        if div:
            self.grads.div_(div)

    def sgd_step(self):
        if self._tail_sgd_done:      # allreduce_grads already updated the classifier on the communication stream
            _lib.check(self._lib.sb_net_sgd_step_layers(self._h, _lib.ptr(self.grads), 0, self.lr, self.momentum,
                                                        self.weight_decay, 0, self._tail_layer, 1, _lib.stream_ptr()),
                       "sb_net_sgd_step_layers")
            self._tail_sgd_done = False
        else:
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
        self.allreduce_grads(sgd_tail=True)
        self.sgd_step()

    def loss_value(self):
        return float(self.loss.item())

    def step(self, inputs, targets, masks=None):
        """One optimisation step; returns the mean BCE loss as a python float (device -> host read)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._enable_overlap()
        self.forward_backward(inputs, targets, masks)
        self.allreduce_grads(sgd_tail=True)
        self.sgd_step()
        return float(self.loss.item())

    def layer_params(self, layer, grads=False):
        wshape, bshape = self._shapes[layer]
        w = np.zeros(wshape, dtype=np.float32)
        b = np.zeros(bshape, dtype=np.float32)
        _lib.check(self._lib.sb_net_get_layer_params(self._h, int(layer), _lib.ptr(self.grads) if grads else None,
                                                     _lib.ptr(w), _lib.ptr(b)), "sb_net_get_layer_params")
        return w, b

    def layer_momentum(self, layer):
        """SGD momentum buffers of one layer in torch layout (zeros before the first step)."""
        wshape, bshape = self._shapes[layer]
        w = np.zeros(wshape, dtype=np.float32)
        b = np.zeros(bshape, dtype=np.float32)
        _lib.check(self._lib.sb_net_get_layer_momentum(self._h, int(layer), _lib.ptr(w), _lib.ptr(b)),
                   "sb_net_get_layer_momentum")
        return w, b

    def load_momentum(self, state):
        """{layer: (weight buffer, bias buffer)} in torch layout -> device (checkpoint resume)."""
        for layer, (w, b) in state.items():
            wshape, bshape = self._shapes[layer]
            w = np.ascontiguousarray(np.asarray(w, dtype=np.float32).reshape(wshape))
            b = np.ascontiguousarray(np.asarray(b, dtype=np.float32).reshape(bshape))
            _lib.check(self._lib.sb_net_set_layer_momentum(self._h, int(layer), _lib.ptr(w), _lib.ptr(b)),
                       "sb_net_set_layer_momentum")
        if state:
            self.step_count = max(self.step_count, 1)

    def export_layers(self, template_layers):
        """Layer list with the current weights (for a fresh inference ``B200Net`` or a checkpoint)."""
        out = []
        for i, l in enumerate(template_layers):
            w, b = self.layer_params(i)
            out.append(dict(l, weight=w, bias=b))
        return out


def _trainable_modules(model):
    """Conv / Linear modules of ``model`` in layer order, with the checks that make the B200 training step valid.
    Raises for modules whose TRAIN-mode behaviour the kernels do not implement (``layers_from_module`` folds BatchNorm
    with eval-mode statistics and applies HeartENN's renorm once — correct for inference, wrong for training)."""
    import torch.nn as nn
    from .nets import unwrap_non_strand_specific
    inner, strand_mode = unwrap_non_strand_specific(model)
    if strand_mode:
        raise ValueError("training a NonStrandSpecific-wrapped model is not supported on the B200 training step")
    if inner.__class__.__name__ == "HeartENN" or hasattr(inner, "_renorm"):
        raise ValueError("models that re-normalise their weights on every forward (HeartENN) cannot be trained here")
    out = []
    for m in inner.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            raise ValueError("BatchNorm layers are not supported by the B200 training step (batch statistics are not "
                             "implemented; folding running statistics would train a different model)")
        if isinstance(m, (nn.LSTM, nn.GRU, nn.RNN)):
            raise ValueError("recurrent layers are not supported by the B200 training step")
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            out.append(m)
    return out


class TrainModel(object):
    """``selene_sdk.TrainModel`` (train_model.py:36-686) on the B200 training step: same constructor arguments,
    ``train`` / ``validate`` / ``evaluate`` / ``train_and_validate``, the reference's checkpoint files
    (``checkpoint.pth.tar`` / ``best_model.pth.tar`` holding ``step, arch, state_dict, min_loss, optimizer``,
    :430-448,638-686) and ``checkpoint_resume`` (:296-326).

    What the kernels implement is ``BCELoss`` + ``torch.optim.SGD(lr, momentum, weight_decay)`` (the pair of
    models/deepsea.py:60-74) on Conv / Linear / ReLU / MaxPool / Dropout stacks; any other ``loss_criterion``,
    ``optimizer_class`` or optimizer option raises instead of silently training something else.  Extra arguments:
    ``sequence_length``, ``device``, ``precision`` ("fp32" parity path or "bf16" tcgen05 path)."""

    def __init__(self, model, data_sampler, loss_criterion=None, optimizer_class=None, optimizer_kwargs=None,
                 batch_size=64, max_steps=1000, report_stats_every_n_steps=100, output_dir=None,
                 save_checkpoint_every_n_steps=1000, save_new_checkpoints_after_n_steps=None,
                 report_gt_feature_n_positives=10, n_validation_samples=None, n_test_samples=None, cpu_n_threads=1,
                 use_cuda=True, data_parallel=False, logging_verbosity=2, checkpoint_resume=None, metrics=None,
                 use_scheduler=True, deterministic=False, scheduler_kwargs=None, stopping_criteria=None,
                 sequence_length=1000, device=None, precision="fp32"):
        import os
        import torch
        if not use_cuda:
            raise ValueError("selene_b200 has no CPU path: use_cuda must be True")
        if data_parallel:
            raise ValueError("data_parallel (nn.DataParallel) is replaced by one process per GPU under "
                             "torch.distributed: launch with torchrun, gradients are all-reduced over NCCL")
        if loss_criterion is not None and not (isinstance(loss_criterion, torch.nn.BCELoss) and
                                               loss_criterion.reduction == "mean" and loss_criterion.weight is None):
            raise ValueError("the B200 training step implements torch.nn.BCELoss(reduction='mean') only, got %r"
                             % (loss_criterion,))
        if optimizer_class is not None and optimizer_class is not torch.optim.SGD:
            raise ValueError("the B200 training step implements torch.optim.SGD only, got %r" % (optimizer_class,))
        kw = dict(optimizer_kwargs or {})
        extra = set(kw) - set(["lr", "momentum", "weight_decay", "dampening", "nesterov"])
        if extra or kw.get("dampening", 0) or kw.get("nesterov", False):
            raise ValueError("unsupported SGD options %s (lr, momentum, weight_decay are implemented)"
                             % sorted(extra or ["dampening/nesterov"]))
        self.model = model
        self._modules = _trainable_modules(model)
        self.sampler = data_sampler
        self.batch_size = batch_size
        self.max_steps = max_steps
        self.nth_step_report_stats = report_stats_every_n_steps
        self.nth_step_save_checkpoint = save_checkpoint_every_n_steps or report_stats_every_n_steps
        self._save_new_checkpoints = save_new_checkpoints_after_n_steps
        self.output_dir = output_dir
        if output_dir:
            os.makedirs(output_dir, exist_ok=True)
        self._report_gt_feature_n_positives = report_gt_feature_n_positives
        if metrics is None:
            from sklearn.metrics import average_precision_score, roc_auc_score
            metrics = dict(roc_auc=roc_auc_score, average_precision=average_precision_score)
        self._metrics = metrics
        self._metric_history = {k: [] for k in metrics}
        self._n_validation_samples = n_validation_samples
        self._n_test_samples = n_test_samples
        self._sequence_length, self._device, self._precision = sequence_length, device, precision
        self._start_step = self.step = 0
        self._min_loss = float("inf")
        self._time_per_step, self._train_loss = [], []
        self._validation_data = self._test_data = None
        self._opt_kwargs = dict(lr=kw.get("lr", 0.01), momentum=kw.get("momentum", 0.0),
                                weight_decay=kw.get("weight_decay", 0.0))
        momentum_state = None
        if checkpoint_resume is not None:
            momentum_state = self._load_checkpoint(checkpoint_resume)
        self.trainer = B200Trainer(model, sequence_length, dropout=B200Trainer.dropout_of_module(model),
                                   device=device, precision=precision, **self._opt_kwargs)
        if momentum_state:
            self.trainer.load_momentum(momentum_state)
        # ReduceLROnPlateau (:335-338) runs on a one-parameter stand-in optimizer; its learning rate is the trainer's
        self._scheduler = None
        if use_scheduler:
            sk = dict(scheduler_kwargs if scheduler_kwargs is not None else dict(patience=16, factor=0.8))
            sk.pop("verbose", None)        # removed from torch >= 2.4; the reference's default still passes it
            self._sched_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=self.trainer.lr)
            self._scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self._sched_opt, 'min', **sk)
        self._early_stopping = False
        if type(stopping_criteria) is list and len(stopping_criteria) == 2 and stopping_criteria[0] in self._metrics:
            self._stopping_metric, self._stopping_patience = stopping_criteria
            self._early_stopping, self._stopping_reached = True, False

    # ------------------------------------------------------------------------------------------ weights <-> module
    def sync_model(self):
        """Copies the trained fp32 weights from the device handle back into ``self.model`` (so ``state_dict()``,
        checkpoints and a fresh inference ``B200Net`` see them)."""
        import torch
        with torch.no_grad():
            for i, m in enumerate(self._modules):
                w, b = self.trainer.layer_params(i)
                m.weight.copy_(torch.from_numpy(w).reshape(m.weight.shape))
                if m.bias is not None:
                    m.bias.copy_(torch.from_numpy(b))
        return self.model

    def _optimizer_state(self):
        """``torch.optim.SGD.state_dict()`` layout: parameter ids in ``model.parameters()`` order."""
        import torch
        ids = {id(p): i for i, p in enumerate(self.model.parameters())}
        state = {}
        if self.trainer.step_count > 0:
            for i, m in enumerate(self._modules):
                mw, mb = self.trainer.layer_momentum(i)
                state[ids[id(m.weight)]] = {"momentum_buffer": torch.from_numpy(mw).reshape(m.weight.shape).clone()}
                if m.bias is not None:
                    state[ids[id(m.bias)]] = {"momentum_buffer": torch.from_numpy(mb).clone()}
        group = dict(lr=self.trainer.lr, momentum=self.trainer.momentum, dampening=0,
                     weight_decay=self.trainer.weight_decay, nesterov=False, params=sorted(ids.values()))
        return {"state": state, "param_groups": [group]}

    def _checkpoint_dict(self):
        self.sync_model()
        return {"step": self.step, "arch": self.model.__class__.__name__, "state_dict": self.model.state_dict(),
                "min_loss": self._min_loss, "optimizer": self._optimizer_state()}

    def _save_checkpoint(self, state, is_best, filename="checkpoint"):
        """:638-686."""
        import os
        import shutil
        import torch
        if not self.output_dir:
            return
        cp = os.path.join(self.output_dir, filename)
        torch.save(state, "{0}.pth.tar".format(cp))
        if is_best:
            shutil.copyfile("{0}.pth.tar".format(cp), "{0}.pth.tar".format(os.path.join(self.output_dir, "best_model")))

    def _checkpoint(self):
        """:430-448."""
        from time import strftime
        name = "checkpoint"
        if self._save_new_checkpoints is not None and self._save_new_checkpoints >= self.step:
            name = "checkpoint-{0}".format(strftime("%m%d%H%M%S"))
        self._save_checkpoint(self._checkpoint_dict(), False, filename=name)

    def _load_checkpoint(self, checkpoint_resume):
        """:296-326.  Returns {layer index: (momentum weight, momentum bias)} for the trainer."""
        import torch
        ck = torch.load(checkpoint_resume, map_location="cpu", weights_only=False)
        if "state_dict" not in ck:
            raise ValueError(("'state_dict' not found in file {0} loaded with method `torch.load`. Selene does not "
                              "support continued training of models that were not originally trained using "
                              "Selene.").format(checkpoint_resume))
        own = self.model.state_dict()
        if len(own) != len(ck["state_dict"]):
            raise ValueError("Loaded state dict does not match the model architecture specified")
        self.model.load_state_dict({k: v for k, v in zip(own.keys(), ck["state_dict"].values())})
        self._start_step = self.step = ck["step"]
        if self._start_step >= self.max_steps:
            self.max_steps += self._start_step
        self._min_loss = ck["min_loss"]
        opt = ck.get("optimizer") or {}
        groups = opt.get("param_groups") or [{}]
        for k in ("lr", "momentum", "weight_decay"):
            if k in groups[0]:
                self._opt_kwargs[k] = groups[0][k]
        ids = {id(p): i for i, p in enumerate(self.model.parameters())}
        out = {}
        for i, m in enumerate(self._modules):
            sw = opt.get("state", {}).get(ids[id(m.weight)], {}).get("momentum_buffer")
            sb = opt.get("state", {}).get(ids[id(m.bias)], {}).get("momentum_buffer") if m.bias is not None else None
            if sw is not None:
                out[i] = (sw.detach().cpu().numpy().astype(np.float32),
                          (sb.detach().cpu().numpy() if sb is not None else np.zeros(m.weight.shape[0])).astype(np.float32))
        logger.info("Resuming from checkpoint: step {0}, min loss {1}".format(self._start_step, self._min_loss))
        return out

    # ------------------------------------------------------------------------------------------ the loop
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
        else:
            inputs, targets = self.sampler.sample_device(self.batch_size, seq_dtype="float32", label_dtype="float32")
            t_sampling = time() - t_i
            loss = self.trainer.step(inputs, targets)
        self._train_loss.append(loss)
        self._time_per_step.append(time() - t_i)
        self.step += 1
        if self.step % self.nth_step_report_stats == 0:
            logger.info("[STEP {0}] average number of steps per second: {1:.1f}; training loss: {2}".format(
                self.step, 1. / np.average(self._time_per_step), np.average(self._train_loss)))
            self._time_per_step, self._train_loss = [], []
        return loss, t_sampling

    def _evaluate_on_data(self, data_in_batches):
        """:511-549 through a fresh inference network built from the current weights (eval mode: no dropout)."""
        import torch
        from .nets import B200Net, layers_from_module
        self.sync_model()
        net = B200Net(layers_from_module(self.model), self._sequence_length, device=self._device)
        losses, preds = [], []
        dev = torch.device("cuda", int(net.device))
        for inputs, targets in data_in_batches:
            x = torch.as_tensor(np.asarray(inputs), dtype=torch.float32, device=dev)
            y = torch.as_tensor(np.asarray(targets), dtype=torch.float32, device=dev)
            p = net.forward_onehot(x, precision=self._precision)
            losses.append(float(torch.nn.functional.binary_cross_entropy(p, y).item()))
            preds.append(p.cpu().numpy())
        return float(np.average(losses)), np.vstack(preds)

    def _scores(self, predictions, targets):
        """Per-feature metric averaged over the features with enough positives
        (utils/performance_metrics.py:213-237 semantics: features with fewer positives are skipped)."""
        out = {}
        for name, fn in self._metrics.items():
            vals = []
            for f in range(targets.shape[1]):
                col = targets[:, f]
                if len(np.unique(col)) > 1 and col.sum() > self._report_gt_feature_n_positives:
                    vals.append(fn(col, predictions[:, f]))
            out[name] = float(np.average(vals)) if vals else None
        return out

    def validate(self):
        """:551-602: validation loss and metrics, scheduler step, ``best_model.pth.tar`` on a new minimum."""
        import math
        if self._validation_data is None:
            self._validation_data, self._all_validation_targets = self.sampler.get_validation_set(
                self.batch_size, n_samples=self._n_validation_samples)
        loss, preds = self._evaluate_on_data(self._validation_data)
        scores = self._scores(preds, np.asarray(self._all_validation_targets))
        for name, sc in scores.items():
            logger.info("validation {0}: {1}".format(name, sc))
            self._metric_history[name].append(sc if sc is not None else float("-inf"))
        scores["loss"] = loss
        if self._scheduler is not None:
            self._scheduler.step(math.ceil(loss * 1000.0) / 1000.0)
            self.trainer.lr = float(self._sched_opt.param_groups[0]["lr"])
        if loss < self._min_loss:
            self._min_loss = loss
            self._save_checkpoint(self._checkpoint_dict(), True)
        logger.info("validation loss: {0}".format(loss))
        if self._early_stopping:
            hist = self._metric_history[self._stopping_metric]
            if self._stopping_patience - (len(hist) - int(np.argmax(hist)) - 1) <= 0:
                self._stopping_reached = True
        return scores

    def evaluate(self):
        """:604-636: loss and metrics on the test set; predictions saved next to the checkpoints."""
        import os
        if self._test_data is None:
            self._test_data, self._all_test_targets = self.sampler.get_test_set(self.batch_size,
                                                                                n_samples=self._n_test_samples)
        loss, preds = self._evaluate_on_data(self._test_data)
        scores = self._scores(preds, np.asarray(self._all_test_targets))
        if self.output_dir:
            np.savez_compressed(os.path.join(self.output_dir, "test_predictions.npz"), data=preds)
        scores["loss"] = loss
        return scores

    def train_and_validate(self):
        """:450-467."""
        for step in range(self._start_step, self.max_steps):
            self.step = step
            self.train()
            self.step = step            # train() counts steps taken; the reference checkpoints on the loop index
            if step % self.nth_step_save_checkpoint == 0:
                self._checkpoint()
            if step and step % self.nth_step_report_stats == 0:
                self.validate()
            if self._early_stopping and self._stopping_reached:
                logger.debug("Patience ran out. Stopping early.")
                break
        self.step = min(self.max_steps, self.step + 1)
        self.sync_model()
        if hasattr(self.sampler, "save_dataset_to_file"):
            self.sampler.save_dataset_to_file("train", close_filehandle=True)

"""Architectures of the reference's example models, rebuilt compactly for tests and benches (the
reference's ``models/*.py`` are user-supplied files and do not travel to the GPU box).  Layer order,
shapes and module construction order follow the cited files, so ``torch.manual_seed(s)`` + default
init gives the same weights as constructing the reference class."""
import math

import numpy as np
import torch
import torch.nn as nn


def _conv_block(cin, cout, k, pool=None, drop=None):
    mods = [nn.Conv1d(cin, cout, kernel_size=k), nn.ReLU(inplace=True)]
    if pool:
        mods.append(nn.MaxPool1d(kernel_size=pool, stride=pool))
    if drop is not None:
        mods.append(nn.Dropout(p=drop))
    return mods


class DeepSEA(nn.Module):
    """models/deepsea.py:21-58."""

    def __init__(self, sequence_length=1000, n_genomic_features=919):
        super().__init__()
        self.conv_net = nn.Sequential(*_conv_block(4, 320, 8, 4, 0.2), *_conv_block(320, 480, 8, 4, 0.2),
                                      *_conv_block(480, 960, 8, None, 0.5))
        self.n_channels = int(np.floor((np.floor((sequence_length - 7) / 4.) - 7) / 4.) - 7)
        self.classifier = nn.Sequential(nn.Linear(960 * self.n_channels, n_genomic_features), nn.ReLU(inplace=True),
                                        nn.Linear(n_genomic_features, n_genomic_features), nn.Sigmoid())

    def forward(self, x):
        o = self.conv_net(x)
        return self.classifier(o.view(o.size(0), 960 * self.n_channels))


class HeartENN(nn.Module):
    """models/heartenn.py:21-82 (forward re-normalises every conv/linear weight row to L2 <= 0.9)."""

    def __init__(self, sequence_length=1000, n_genomic_features=90):
        super().__init__()
        self.conv_net = nn.Sequential(
            *_conv_block(4, 60, 8), *_conv_block(60, 60, 8, 4), nn.BatchNorm1d(60),
            *_conv_block(60, 80, 8), *_conv_block(80, 80, 8, 4), nn.BatchNorm1d(80), nn.Dropout(p=0.4),
            *_conv_block(80, 240, 8), *_conv_block(240, 240, 8), nn.BatchNorm1d(240), nn.Dropout(p=0.6))
        self._n_channels = int(np.floor((np.floor((sequence_length - 14) / 4.) - 14) / 4.) - 14)
        self.classifier = nn.Sequential(
            nn.Linear(240 * self._n_channels, n_genomic_features), nn.ReLU(inplace=True),
            nn.BatchNorm1d(n_genomic_features), nn.Linear(n_genomic_features, n_genomic_features), nn.Sigmoid())

    def forward(self, x):
        for layer in list(self.conv_net.children()) + list(self.classifier.children()):
            if isinstance(layer, (nn.Conv1d, nn.Linear)):
                layer.weight.data.renorm_(2, 0, 0.9)
        o = self.conv_net(x)
        return self.classifier(o.view(o.size(0), 240 * self._n_channels))


class DeeperDeepSEA(nn.Module):
    """utils/example_model.py:19-93 — the model of the paper's VEP case study: DeepSEA with every
    conv doubled and BatchNorm after each pool / before the classifier."""

    def __init__(self, sequence_length=1000, n_targets=919):
        super().__init__()
        self.conv_net = nn.Sequential(
            *_conv_block(4, 320, 8), *_conv_block(320, 320, 8, 4), nn.BatchNorm1d(320),
            *_conv_block(320, 480, 8), *_conv_block(480, 480, 8, 4), nn.BatchNorm1d(480), nn.Dropout(p=0.2),
            *_conv_block(480, 960, 8), *_conv_block(960, 960, 8), nn.BatchNorm1d(960), nn.Dropout(p=0.2))
        self._n_channels = int(np.floor((np.floor((sequence_length - 14) / 4.) - 14) / 4.) - 14)
        self.classifier = nn.Sequential(
            nn.Linear(960 * self._n_channels, n_targets), nn.ReLU(inplace=True),
            nn.BatchNorm1d(n_targets), nn.Linear(n_targets, n_targets), nn.Sigmoid())

    def forward(self, x):
        o = self.conv_net(x)
        return self.classifier(o.view(o.size(0), 960 * self._n_channels))


class Seqweaver(nn.Module):
    """models/seqweaver.py:25-61: Conv2d (1,8) stack; forward unsqueezes the height axis itself."""

    def __init__(self, n_classes=217):
        super().__init__()
        self.model = nn.Sequential(
            nn.Conv2d(4, 160, (1, 8)), nn.ReLU(), nn.MaxPool2d((1, 4), (1, 4)), nn.Dropout(0.1),
            nn.Conv2d(160, 320, (1, 8)), nn.ReLU(), nn.MaxPool2d((1, 4), (1, 4)), nn.Dropout(0.1),
            nn.Conv2d(320, 480, (1, 8)), nn.ReLU(), nn.Dropout(0.3))
        self.fc = nn.Sequential(nn.Linear(25440, n_classes), nn.ReLU(), nn.Linear(n_classes, n_classes), nn.Sigmoid())

    def forward(self, x):
        o = self.model(x.unsqueeze(2))
        return self.fc(torch.reshape(o, (o.size(0), 25440)))


class DanQ(nn.Module):
    """models/danQ.py:22-52, including its batch_first quirk: the LSTM is given (75, B, 320) but is
    ``batch_first=True``, so it treats 75 as the batch and recurs over the B samples."""

    def __init__(self, sequence_length=1000, n_genomic_features=919):
        super().__init__()
        self.nnet = nn.Sequential(nn.Conv1d(4, 320, kernel_size=26), nn.ReLU(inplace=True),
                                  nn.MaxPool1d(kernel_size=13, stride=13), nn.Dropout(0.2))
        self.bdlstm = nn.Sequential(nn.LSTM(320, 320, num_layers=1, batch_first=True, bidirectional=True))
        self._n_channels = math.floor((sequence_length - 25) / 13)
        self.classifier = nn.Sequential(nn.Dropout(0.5), nn.Linear(self._n_channels * 640, 925), nn.ReLU(inplace=True),
                                        nn.Linear(925, n_genomic_features), nn.Sigmoid())

    def forward(self, x):
        out = self.nnet(x)
        reshape_out = out.transpose(0, 1).transpose(0, 2)
        out, _ = self.bdlstm(reshape_out)
        out = out.transpose(0, 1)
        return self.classifier(out.contiguous().view(out.size(0), 640 * self._n_channels))


class NonStrandSpecific(nn.Module):
    """utils/non_strand_specific_module.py:26-85 for models that take (B, 4, L): the wrapped model also sees the
    input with its channel axis and its position axis flipped; outputs are averaged or maxed."""

    def __init__(self, model, mode="mean"):
        super().__init__()
        if mode != "mean" and mode != "max":
            raise ValueError("Mode should be one of 'mean' or 'max' but was{0}.".format(mode))
        self.model = model
        self.mode = mode

    def forward(self, x):
        out = self.model.forward(x)
        out_rev = self.model.forward(torch.flip(x, dims=(1, 2)))
        return (out + out_rev) / 2 if self.mode == "mean" else torch.max(out, out_rev)


def randomize_batchnorm(model, seed=0):
    """Gives eval-mode BatchNorm layers non-trivial running statistics / affine terms."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.copy_(torch.rand(m.num_features, generator=g) - 0.5)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.rand(m.num_features, generator=g) - 0.5)

"""``CNN`` module definition, state_dict-compatible with the reference (training/cnn_trainer.py:30-139).

``net = Sequential(features, classifier)`` registers every parameter twice, so saved state dicts carry
both ``features.*``/``classifier.*`` and ``net.0.*``/``net.1.*`` keys; keeping the same
aliasing lets reference-written checkpoints load with strict=True.  ``init_args`` omits
``use_batchnorm`` exactly as the reference does (:59-69), so reloaded models always have BatchNorm.
"""
from __future__ import annotations

import torch.nn as nn


class CNN(nn.Module):
    def __init__(self, num_classes, in_channels=1, base_channels=32, num_blocks=3, hidden_dim=256,
                 dropout=0.1, kernel_size=3, use_batchnorm=True, use_maxpool=True, adaptive_pool=(4, 4)):
        super().__init__()
        self.init_args = dict(num_classes=num_classes, in_channels=in_channels, base_channels=base_channels,
                              num_blocks=num_blocks, hidden_dim=hidden_dim, dropout=dropout,
                              kernel_size=kernel_size, use_maxpool=use_maxpool, adaptive_pool=adaptive_pool)
        widths = [in_channels] + [base_channels << b for b in range(num_blocks)]
        feats = []
        for c_in, c_out in zip(widths[:-1], widths[1:]):
            feats.append(nn.Conv2d(c_in, c_out, kernel_size=kernel_size, padding=kernel_size // 2))
            if use_batchnorm:
                feats.append(nn.BatchNorm2d(c_out))
            feats.append(nn.LeakyReLU(inplace=True))
            if use_maxpool:
                feats.append(nn.MaxPool2d(2))
            if dropout > 0.0:
                feats.append(nn.Dropout(dropout))
        feats.append(nn.AdaptiveAvgPool2d(adaptive_pool))
        self.features = nn.Sequential(*feats)
        flat = widths[-1] * adaptive_pool[0] * adaptive_pool[1]
        head = [nn.Flatten()]
        if hidden_dim is not None and hidden_dim > 0:
            head += [nn.Linear(flat, hidden_dim), nn.LeakyReLU(inplace=True)]
            if dropout > 0.0:
                head.append(nn.Dropout(dropout))
            head.append(nn.Linear(hidden_dim, num_classes))
        else:
            head.append(nn.Linear(flat, num_classes))
        self.classifier = nn.Sequential(*head)
        self.net = nn.Sequential(self.features, self.classifier)

    def forward(self, x):
        return self.net(x)

"""``MLP`` module definition, state_dict-compatible with the reference (training/mlp_trainer.py:32-105).

Only the class the model-loading layer needs (note_predictor.py:40-49); the forward pass used for
inference runs in the CUDA extension (csrc/infer.cu), this module is the weight container.
"""
from __future__ import annotations

import torch.nn as nn


def hidden_dims(hidden_dim: int, num_hidden_layers: int) -> list[int]:
    """Widths halve per layer and stop before dropping under 8 (mlp_trainer.py:49-54)."""
    dims = [hidden_dim]
    while len(dims) < num_hidden_layers and dims[-1] // 2 >= 8:
        dims.append(dims[-1] // 2)
    return dims


class MLP(nn.Module):
    def __init__(self, num_features, hidden_dim, num_hidden_layers, num_classes, dropout=0.1):
        super().__init__()
        self.init_args = dict(num_features=num_features, hidden_dim=hidden_dim,
                              num_hidden_layers=num_hidden_layers, num_classes=num_classes, dropout=dropout)
        dims = hidden_dims(hidden_dim, num_hidden_layers)
        mods, fan_in = [], num_features
        for width in dims:
            mods += [nn.Linear(fan_in, width), nn.LayerNorm(width), nn.LeakyReLU(0.1)]
            if dropout > 0:
                mods.append(nn.Dropout(dropout))
            fan_in = width
        mods.append(nn.Linear(fan_in, num_classes))
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)

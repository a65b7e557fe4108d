"""Value / Q network: the reference's layer shape (``model/model.py:5-20``: five ``nn.Linear`` layers,
in -> 128 -> 128 -> 128 -> 128 -> out, ReLU between them) with a constructor that works (the reference's
``super(Model, self).__init()`` at ``model/model.py:7`` raises AttributeError).  Stays in PyTorch: the north star
keeps the MLP value head out of the hand-written kernels."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class Model(nn.Module):
    def __init__(self, state_space_size: int, action_space_size: int, hidden: int = 128):
        super().__init__()
        self.layer1 = nn.Linear(state_space_size, hidden)
        self.layer2 = nn.Linear(hidden, hidden)
        self.layer3 = nn.Linear(hidden, hidden)
        self.layer4 = nn.Linear(hidden, hidden)
        self.layer5 = nn.Linear(hidden, action_space_size)

    def forward(self, x):
        x = F.relu(self.layer1(x))
        x = F.relu(self.layer2(x))
        x = F.relu(self.layer3(x))
        x = F.relu(self.layer4(x))
        return self.layer5(x)


# scale of the four afterstate features (rows cleared, holes, bumpiness, aggregate height) -> O(1) inputs
FEATURE_SCALE = (1.0 / 4.0, 1.0 / 40.0, 1.0 / 40.0, 1.0 / 100.0)


class ValueNet(Model):
    """Afterstate value V(features): ``Model(4, 1)`` on the normalised feature row."""

    def __init__(self, hidden: int = 128):
        super().__init__(4, 1, hidden)
        self.register_buffer("scale", torch.tensor(FEATURE_SCALE, dtype=torch.float32))

    def forward(self, feats):
        return super().forward(feats.to(torch.float32) * self.scale).squeeze(-1)

    @torch.no_grad()
    def rank_bf16(self, feats):
        """Inference-only forward in bf16 for RANKING afterstates (action selection over 40 * N rows): each hidden layer is one
        cuBLASLt GEMM with the bias + ReLU epilogue fused (``torch._addmm_activation``), i.e. no separate ReLU pass over the
        [40 N, 128] activations -- a quarter of the eager forward's GPU time.  Same weights as ``forward``."""
        x = (feats.to(torch.float32) * self.scale).to(torch.bfloat16)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            x = torch._addmm_activation(layer.bias.to(torch.bfloat16), x, layer.weight.to(torch.bfloat16).t())
        return torch.addmm(self.layer5.bias.to(torch.bfloat16), x, self.layer5.weight.to(torch.bfloat16).t()).squeeze(-1)

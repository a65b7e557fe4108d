"""``ValueKernel``: the value net's RANKING forward pass as one fused tensor-core kernel (``tpl_value_rows``), and the per-env
action selection over the distinct placements (``tpl_select_action``).

The network and its training stay in PyTorch (``model.ValueNet``: the reference's layer shape ``model/model.py:9-20``); this
class only mirrors the current parameters into the packed bf16 blob the kernel keeps in shared memory (``sync``), and is what
the rollout calls every step instead of five library GEMMs over [40 N, 128] activations.  Numerics: bf16 operands, fp32
accumulation, bf16 activations between layers -- the same as ``ValueNet.rank_bf16``."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .model import ValueNet

BLOB_BYTES = 119328          # TPL_VALUE_BLOB_BYTES


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class ValueKernel:
    def __init__(self, net: ValueNet):
        self._L = _lib.lib()
        self.net = net
        p = next(net.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("ValueKernel is CUDA-only (tcgen05 tensor cores, sm_100a); there is no CPU path")
        if net.layer1.weight.shape != (128, 4) or net.layer5.weight.shape != (1, 128):
            raise ValueError("ValueKernel is built for the 4 -> 128 -> 128 -> 128 -> 128 -> 1 net")
        self.device = p.device
        self.blob = torch.zeros(BLOB_BYTES, dtype=torch.uint8, device=self.device)
        self._scale = (ctypes.c_float * 4)(*[float(v) for v in net.scale.tolist()])
        self.sync()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, fn, what, *args):
        if torch.cuda.current_device() == self.device.index:
            _lib.check(fn(*args), what)
        else:
            with torch.cuda.device(self.device):
                _lib.check(fn(*args), what)

    def sync(self) -> None:
        """Re-pack the net's current parameters (one small kernel; call after every optimiser step that should be seen)."""
        n = self.net
        ps = [n.layer1.weight, n.layer1.bias, n.layer2.weight, n.layer2.bias, n.layer3.weight, n.layer3.bias,
              n.layer4.weight, n.layer4.bias, n.layer5.weight, n.layer5.bias]
        if all(p.dtype == torch.float32 and p.is_contiguous() for p in ps):       # the usual case: the optimiser updates them in place
            ptrs = [ctypes.c_void_p(p.data_ptr()) for p in ps]
        else:
            self._keep = [p.detach().to(torch.float32).contiguous() for p in ps]
            ptrs = [_ptr(p) for p in self._keep]
        self._call(self._L.tpl_value_pack, "tpl_value_pack", *ptrs, ctypes.cast(self._scale, ctypes.c_void_p), _ptr(self.blob), self._stream())

    def values(self, rows: torch.Tensor, count: torch.Tensor = None, out: torch.Tensor = None) -> torch.Tensor:
        """rows: int32/uint32 CUDA tensor of feature words (distinct-placements form); count: optional 0-d / 1-element int32
        device tensor = number of valid rows (read on the device).  Returns float32 [len(rows)] (entries >= count untouched)."""
        if out is None:
            out = torch.zeros(rows.numel(), dtype=torch.float32, device=self.device)
        self._call(self._L.tpl_value_rows, "tpl_value_rows", _ptr(rows), _ptr(count), rows.numel(), _ptr(self.blob), _ptr(out), self._stream())
        return out

    def select(self, rows, runs, values, gamma: float, eps: float, seed: int, step: int, env_base: int = 0, reward_win: float = 10.0,
               reward_lose: float = -10.0, out=None):
        """Epsilon-greedy arg-max over each env's distinct placements of (rows cleared + win / lose reward + gamma * V).
        Returns (rot uint8[N], loc uint8[N], chosen int32[N] feature word, q float32[N])."""
        n = runs.numel()
        if out is None:
            out = (torch.empty(n, dtype=torch.uint8, device=self.device), torch.empty(n, dtype=torch.uint8, device=self.device),
                   torch.empty(n, dtype=torch.int32, device=self.device), torch.empty(n, dtype=torch.float32, device=self.device))
        rot, loc, chosen, q = out
        self._call(self._L.tpl_select_action, "tpl_select_action", _ptr(rows), _ptr(runs), _ptr(values), n, float(gamma), float(reward_win),
                   float(reward_lose), float(eps), int(seed), int(env_base), int(step) & 0xFFFFFFFF, _ptr(rot), _ptr(loc), _ptr(chosen), _ptr(q),
                   self._stream())
        return rot, loc, chosen, q


@torch.no_grad()
def reference_values(net: ValueNet, rows: torch.Tensor) -> torch.Tensor:
    """The same forward pass in plain PyTorch with the kernel's numerics spelled out (bf16 operands and activations, fp32
    accumulation, last layer in fp32 from the bf16-rounded activations): what the parity test compares ``tpl_value_rows`` with."""
    w = rows.view(torch.uint8).view(-1, 4).to(torch.float32)
    w[:, 0] = (rows.view(torch.int32) & 7).to(torch.float32)
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)   # noqa: E731
    x = bf(w * net.scale)
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        x = bf(torch.relu(x @ bf(layer.weight).t() + bf(layer.bias)))
    return (x @ bf(net.layer5.weight).t()).squeeze(-1) + bf(net.layer5.bias)

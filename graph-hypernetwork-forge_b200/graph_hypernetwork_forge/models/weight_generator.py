"""WeightGenerator - hypernetwork that turns relation text embeddings into GNN weights.

Drop-in for the reference class of the same name
(`graph_hypernetwork_forge/models/weight_generator.py:33-143`): same constructor,
attributes, parameter names (`generators.{W_msg,W_self,bias}.<i>.{weight,bias}`,
`log_scales.{W_msg,W_self,bias}`), initialisation stream and return convention.
The forward pass runs on the B200 through `ghf_linear` (ReLU and the `exp(log_scale)`
factor fused into the store); there is no CPU path.  Gradients: `autograd.LinearFn`.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn

from .. import _native, autograd

_KINDS = ("W_msg", "W_self", "bias")


class WeightGenerator(nn.Module):
    """Three small MLPs, one per generated tensor, sharing nothing but the input.

    Args mirror the reference (weight_generator.py:50-59): ``text_dim``, ``d_in``,
    ``d_out``, ``hidden_dim=128``, ``num_hidden=2``, ``dropout=0.0``,
    ``init_scale=0.01``.
    """

    def __init__(self, text_dim: int, d_in: int, d_out: int, hidden_dim: int = 128, num_hidden: int = 2,
                 dropout: float = 0.0, init_scale: float = 0.01) -> None:
        super().__init__()
        if min(text_dim, d_in, d_out) <= 0:
            raise ValueError("text_dim, d_in, d_out must all be positive integers")
        self.text_dim, self.d_in, self.d_out = text_dim, d_in, d_out
        self.init_scale = init_scale
        self._dropout = float(dropout)
        self.grad_tf32 = False     # backward GEMMs of the MLPs in TF32 (HyperGNN sets it in tensor-core precision modes)
        self._shapes = {"W_msg": (d_in, d_out), "W_self": (d_in, d_out), "bias": (d_out,)}

        # Parameter creation order matters: with the same torch seed the drop-in must draw the
        # same random stream as the reference (all MLPs first, then the last-layer re-init).
        self.generators = nn.ModuleDict()
        for kind in _KINDS:
            stack, width = [], text_dim
            for _ in range(num_hidden):
                stack += [nn.Linear(width, hidden_dim), nn.ReLU()]
                if dropout > 0.0:
                    stack.append(nn.Dropout(dropout))  # keeps the reference's Sequential indices
                width = hidden_dim
            stack.append(nn.Linear(width, math.prod(self._shapes[kind])))
            self.generators[kind] = nn.Sequential(*stack)
        self.log_scales = nn.ParameterDict(
            {kind: nn.Parameter(torch.full((1,), math.log(init_scale))) for kind in _KINDS})
        for kind in _KINDS:  # near-zero generated weights at init (weight_generator.py:109-114)
            head = [m for m in self.generators[kind] if isinstance(m, nn.Linear)][-1]
            nn.init.zeros_(head.bias)
            nn.init.normal_(head.weight, std=0.01)

    # ------------------------------------------------------------------
    def _run_mlp(self, kind: str, x: torch.Tensor) -> torch.Tensor:
        """flat = MLP_kind(x) * exp(log_scale_kind), every Linear (+ReLU, + the scale) one fused kernel; with
        gradients enabled the same kernels run under `autograd.LinearFn`."""
        mods = list(self.generators[kind])
        linears = [i for i, m in enumerate(mods) if isinstance(m, nn.Linear)]
        ls = self.log_scales[kind]
        grad = torch.is_grad_enabled() and autograd.wants_grad(x, ls, *self.generators[kind].parameters())
        for pos, i in enumerate(linears):
            last = pos == len(linears) - 1
            if grad:
                x = autograd.linear(x, mods[i].weight, mods[i].bias, relu=not last, log_scale=ls if last else None,
                                    tf32=self.grad_tf32)
            else:
                x = _native.linear(x, mods[i].weight, mods[i].bias, relu=not last, log_scale=ls if last else None)
            if not last and self.training and self._dropout > 0.0:
                x = nn.functional.dropout(x, self._dropout)       # the Dropout module after each hidden ReLU
        return x

    def hidden(self, kind: str, x: torch.Tensor):
        """(input of the last Linear of MLP `kind`, that Linear) - inference only; what the generator -> operand-image
        fusion consumes instead of the generated fp32 tensor (`_native.weight_images`)."""
        linears = [m for m in self.generators[kind] if isinstance(m, nn.Linear)]
        for lin in linears[:-1]:
            x = _native.linear(x, lin.weight, lin.bias, relu=True)
        return x, linears[-1]

    def forward(self, text_emb: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``[text_dim]`` or ``[B, text_dim]`` -> {"W_msg", "W_self", "bias"} (unbatched in, unbatched out)."""
        _native.require_cuda(text_emb, self.log_scales["W_msg"])
        single = text_emb.dim() == 1
        x = text_emb.unsqueeze(0) if single else text_emb
        out: Dict[str, torch.Tensor] = {}
        for kind in _KINDS:
            w = self._run_mlp(kind, x).view(x.size(0), *self._shapes[kind])
            out[kind] = w.squeeze(0) if single else w
        return out

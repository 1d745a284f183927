"""Autograd for the B200 path: `torch.autograd.Function`s whose forward is the native kernel and whose backward
is the native gradient kernel (message passing, text encoder) or a plain library GEMM (the Linear layers).

The reference trains through `HyperGNN.forward` (tests/test_hypergnn.py:183-226, demo.py:79-101);
with these the drop-in does too: when gradients are enabled and something requires them, `HyperGNN.forward`,
`WeightGenerator.forward` and `TextEncoder.forward` route through here, otherwise through the forward-only calls.
"""
from __future__ import annotations

import torch

from . import _native


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class LinearFn(torch.autograd.Function):
    """y = exp(log_scale) * act(x W^T + b) (ghf_linear).  Backward: three GEMMs through torch.matmul.
    `shadow` (a list, optional) receives the fp16 Shadow of y made by the same kernel (the input projection)."""

    @staticmethod
    def forward(ctx, x, weight, bias, log_scale, relu: bool, shadow):
        if shadow is not None:
            y, y16 = _native.linear(x, weight, bias, relu=relu, log_scale=log_scale, want_f16=True)
            shadow.append(y16)
        else:
            y = _native.linear(x, weight, bias, relu=relu, log_scale=log_scale)
        ctx.relu = relu
        ctx.save_for_backward(x, weight, log_scale, y)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, weight, log_scale, y = ctx.saved_tensors
        g_pre = g_y
        if ctx.relu:
            g_pre = g_pre * (y > 0)          # exp(log_scale) > 0: y and the pre-activation share their sign
        g_ls = None
        if log_scale is not None:
            if ctx.needs_input_grad[3]:
                g_ls = (g_y * y).sum().reshape(log_scale.shape)
            g_pre = g_pre * log_scale.exp()
        g_x = g_pre @ weight if ctx.needs_input_grad[0] else None
        g_w = g_pre.t() @ x if ctx.needs_input_grad[1] else None
        g_b = g_pre.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return g_x, g_w, g_b, g_ls, None, None


def linear(x, weight, bias, relu=False, log_scale=None, shadow=None):
    return LinearFn.apply(x, weight, bias, log_scale, relu, shadow)


class TextEncodeFn(torch.autograd.Function):
    """tanh(mean_i Emb[tok_i] Wp^T + bp) over packed strings (ghf_text_encode / ghf_text_encode_backward)."""

    @staticmethod
    def forward(ctx, char_emb, proj_w, proj_b, utf8, offsets, index, num: int):
        out = _native.text_encode(utf8, offsets, index, num, char_emb, proj_w, proj_b)
        ctx.save_for_backward(char_emb, proj_w, out)
        ctx.packed = (utf8, offsets, index, num)
        return out

    @staticmethod
    def backward(ctx, g_out):
        char_emb, proj_w, out = ctx.saved_tensors
        utf8, offsets, index, num = ctx.packed
        g_emb, g_w, g_b = _native.text_encode_backward(utf8, offsets, index, num, char_emb, proj_w, out,
                                                       g_out.contiguous())
        return g_emb, g_w, g_b, None, None, None, None


class MPLayerFn(torch.autograd.Function):
    """One message-passing layer (ghf_mp_layer_f16) with its gradients w.r.t. h, the generated relation tensors
    and the LayerNorm parameters."""

    @staticmethod
    def forward(ctx, h, W_msg, W_self, bias, ln_w, ln_b, graph, eps: float, precision: int, h16, out16):
        out, upd = graph.mp_layer(h, W_msg, W_self, bias, ln_w, ln_b, eps, precision, want_upd=True, h16=h16,
                                  out16=out16)
        ctx.graph, ctx.eps, ctx.precision, ctx.h16 = graph, eps, precision, h16
        ctx.save_for_backward(h, W_msg, W_self, ln_w, upd)
        return out

    @staticmethod
    def backward(ctx, g_out):
        h, W_msg, W_self, ln_w, upd = ctx.saved_tensors
        graph, prec = ctx.graph, ctx.precision
        f16 = prec == _native.PREC_F16
        # g_acc travels as ONE fp16 shadow to the two contractions and to the weight gradients (its max comes out
        # of the epilogue kernel itself)
        g_pre, g_acc, g_ln_w, g_ln_b, g16 = graph.epilogue_backward(g_out.contiguous(), upd, h, ln_w, ctx.eps,
                                                                    want_shadow=f16 and graph.hidden_dim == 128)
        g_h = None
        if ctx.needs_input_grad[0]:
            zero_w = torch.zeros_like(W_msg)
            zero_b = torch.zeros(W_msg.shape[:2], dtype=W_msg.dtype, device=W_msg.device)
            g_h = g_pre                                    # the residual's share; the other two are added in place
            # messages: g_acc_v W_msg[r]^T lands on the SOURCE u - the same contraction over the reversed edges
            graph.reversed().contract(g_acc, W_msg.transpose(1, 2).contiguous(), zero_w, zero_b, prec, x16=g16,
                                      out=g_h, accumulate=True)
            # self-loop: g_acc_v W_self[r]^T summed over v's in-edges stays at v
            graph.contract(g_acc, zero_w, W_self.transpose(1, 2).contiguous(), zero_b, prec, x16=g16, out=g_h,
                           accumulate=True)
        g_wm = g_ws = g_b = None
        if any(ctx.needs_input_grad[1:4]):
            g_wm, g_ws, g_b = graph.weight_grad(h, g_acc, prec, h16=ctx.h16, g16=g16)
        return g_h, g_wm, g_ws, g_b, g_ln_w, g_ln_b, None, None, None, None, None


def mp_layer(graph, h, W_msg, W_self, bias, ln_w, ln_b, eps, precision, h16=None, out16=None):
    return MPLayerFn.apply(h, W_msg, W_self, bias, ln_w, ln_b, graph, eps, precision, h16, out16)

"""Autograd for the B200 path: `torch.autograd.Function`s whose forward is the native kernel and whose backward
is the native gradient kernel (message passing, text encoder, Linear layers).

The reference trains through `HyperGNN.forward` (tests/test_hypergnn.py:183-226, demo.py:79-101);
with these the drop-in does too: when gradients are enabled and something requires them, `HyperGNN.forward`,
`WeightGenerator.forward` and `TextEncoder.forward` route through here, otherwise through the forward-only calls.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _native


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class LinearFn(torch.autograd.Function):
    """y = exp(log_scale) * act(x W^T + b) (ghf_linear).  Backward: `ghf_linear_backward` - two fp32 kernels
    (dL/dx; dL/dW with the bias and log-scale gradients summed from the same tiles), the ReLU mask applied while the
    tiles are loaded.  `shadow` (a list, optional) receives the fp16 Shadow of y made by the same kernel (the input
    projection).  `tf32` is kept for callers of round 1 (the library GEMMs it selected are gone) and ignored."""

    @staticmethod
    def forward(ctx, x, weight, bias, log_scale, relu: bool, shadow, tf32: bool):
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
    @once_differentiable
    def backward(ctx, g_y):
        x, weight, log_scale, y = ctx.saved_tensors
        need = ctx.needs_input_grad
        g_x, g_w, g_b, g_ls = _native.linear_backward(
            x, weight, log_scale, y, g_y.contiguous(), ctx.relu, need_x=need[0], need_w=need[1],
            need_b=ctx.has_bias and need[2], need_ls=log_scale is not None and need[3])
        if g_ls is not None:
            g_ls = g_ls.reshape(log_scale.shape)
        return g_x, g_w, g_b, g_ls, None, None, None


def linear(x, weight, bias, relu=False, log_scale=None, shadow=None, tf32=False):
    return LinearFn.apply(x, weight, bias, log_scale, relu, shadow, tf32)


class TextEncodeFn(torch.autograd.Function):
    """tanh(mean_i Emb[tok_i] Wp^T + bp) over packed strings (ghf_text_encode / ghf_text_encode_backward)."""

    @staticmethod
    def forward(ctx, char_emb, proj_w, proj_b, utf8, offsets, index, num: int):
        out = _native.text_encode(utf8, offsets, index, num, char_emb, proj_w, proj_b)
        ctx.save_for_backward(char_emb, proj_w, out)
        ctx.packed = (utf8, offsets, index, num)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        char_emb, proj_w, out = ctx.saved_tensors
        utf8, offsets, index, num = ctx.packed
        g_emb, g_w, g_b = _native.text_encode_backward(utf8, offsets, index, num, char_emb, proj_w, out,
                                                       g_out.contiguous())
        return g_emb, g_w, g_b, None, None, None, None


def _contraction_backward(graph, prec, h, W_msg, W_self, g_acc, g16, h16, needs, g_h):
    """Gradients of acc_v = sum_{(u->v), r} (h_u W_msg[r] + h_v W_self[r] + bias[r]) given g_acc = dL/d acc.
    `g_h` (or None): buffer that already holds the other shares of dL/dh; the two contraction terms are added to it.
    -> (g_h, g_W_msg, g_W_self, g_bias)"""
    if needs[0]:
        rev = graph.reversed()
        if prec == _native.PREC_F16 and graph.hidden_dim == 128:
            # the f16 engine transposes while packing and skips the absent half of K (no zero tensors, no copies)
            g_h = rev.contract(g_acc, W_msg, None, None, prec, x16=g16, out=g_h, accumulate=g_h is not None,
                               transposed=True)
            graph.contract(g_acc, None, W_self, None, prec, x16=g16, out=g_h, accumulate=True, transposed=True)
        else:
            zero_w = torch.zeros_like(W_msg)
            zero_b = torch.zeros(W_msg.shape[:2], dtype=W_msg.dtype, device=W_msg.device)
            # messages: g_acc_v W_msg[r]^T lands on the SOURCE u - the same contraction over the reversed edges
            g_h = rev.contract(g_acc, W_msg.transpose(1, 2).contiguous(), zero_w, zero_b, prec, x16=g16, out=g_h,
                               accumulate=g_h is not None)
            # self-loop: g_acc_v W_self[r]^T summed over v's in-edges stays at v
            graph.contract(g_acc, zero_w, W_self.transpose(1, 2).contiguous(), zero_b, prec, x16=g16, out=g_h,
                           accumulate=True)
    else:
        g_h = None
    g_wm = g_ws = g_b = None
    if any(needs[1:4]):
        g_wm, g_ws, g_b = graph.weight_grad(h, g_acc, prec, h16=h16, g16=g16)
    return g_h, g_wm, g_ws, g_b


class MPLayerFn(torch.autograd.Function):
    """One message-passing layer (ghf_mp_layer_f16) with its gradients w.r.t. h, the generated relation tensors
    and the LayerNorm parameters."""

    @staticmethod
    def forward(ctx, h, W_msg, W_self, bias, ln_w, ln_b, graph, eps: float, precision: int, h16, out16, dropout=None):
        out, upd = graph.mp_layer(h, W_msg, W_self, bias, ln_w, ln_b, eps, precision, want_upd=True, h16=h16,
                                  out16=out16, dropout=dropout)
        ctx.graph, ctx.eps, ctx.precision, ctx.h16, ctx.dropout = graph, eps, precision, h16, dropout
        ctx.save_for_backward(h, W_msg, W_self, ln_w, upd)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        h, W_msg, W_self, ln_w, upd = ctx.saved_tensors
        graph, prec = ctx.graph, ctx.precision
        f16 = prec == _native.PREC_F16
        # g_acc travels as ONE fp16 shadow to the two contractions and to the weight gradients (its max comes out
        # of the epilogue kernel itself); g_pre, the residual's share of dL/dh, is the buffer the others add to
        g_pre, g_acc, g_ln_w, g_ln_b, g16 = graph.epilogue_backward(g_out.contiguous(), upd, h, ln_w, ctx.eps,
                                                                    want_shadow=f16 and graph.hidden_dim == 128,
                                                                    dropout=ctx.dropout)
        g_h, g_wm, g_ws, g_b = _contraction_backward(graph, prec, h, W_msg, W_self, g_acc, g16, ctx.h16,
                                                     ctx.needs_input_grad, g_pre)
        return g_h, g_wm, g_ws, g_b, g_ln_w, g_ln_b, None, None, None, None, None, None


def mp_layer(graph, h, W_msg, W_self, bias, ln_w, ln_b, eps, precision, h16=None, out16=None, dropout=None):
    """`dropout` (_native.DropoutState): training-mode dropout inside the row epilogue (HG:293-294)."""
    return MPLayerFn.apply(h, W_msg, W_self, bias, ln_w, ln_b, graph, eps, precision, h16, out16, dropout)


class MPUpdateFn(torch.autograd.Function):
    """Only the pre-residual update of a layer, upd_v = acc_v / max(indeg_v, 1) (HG:160-230, what the reference's
    `_message_passing` returns).  Used by the reference-shaped `_message_passing` entry, and for training-mode dropout
    (HG:293-294) on tensors whose size torch's vectorised dropout kernel does not cover (numel % 4 != 0), where the rest
    of the layer runs as torch ops; otherwise dropout happens inside the native row epilogue (MPLayerFn)."""

    @staticmethod
    def forward(ctx, h, W_msg, W_self, bias, graph, precision: int, h16):
        acc = graph.contract(h, W_msg, W_self, bias, precision, x16=h16)
        inv = 1.0 / graph.in_degree().clamp(min=1).to(acc.dtype)
        ctx.graph, ctx.precision, ctx.h16 = graph, precision, h16
        ctx.save_for_backward(h, W_msg, W_self, inv)
        return acc.mul_(inv.unsqueeze(1))

    @staticmethod
    @once_differentiable
    def backward(ctx, g_upd):
        h, W_msg, W_self, inv = ctx.saved_tensors
        g_acc = g_upd * inv.unsqueeze(1)
        g_h, g_wm, g_ws, g_b = _contraction_backward(ctx.graph, ctx.precision, h, W_msg, W_self, g_acc, None, ctx.h16,
                                                     ctx.needs_input_grad, None)
        return g_h, g_wm, g_ws, g_b, None, None, None


def mp_update(graph, h, W_msg, W_self, bias, precision, h16=None):
    return MPUpdateFn.apply(h, W_msg, W_self, bias, graph, precision, h16)


class ScorePairsFn(torch.autograd.Function):
    """out[b] = <emb[heads[b]], emb[tails[b]]> (ghf_score_pairs / ghf_score_pairs_backward)."""

    @staticmethod
    def forward(ctx, emb, heads, tails):
        ctx.save_for_backward(emb, heads, tails)
        return _native.score_pairs(emb, heads, tails)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out):
        emb, heads, tails = ctx.saved_tensors
        return _native.score_pairs_backward(emb, heads, tails, g_out.contiguous()), None, None

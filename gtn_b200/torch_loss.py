"""PyTorch criteria on CUDA tensors, end to end on the device.

The reference's example (bindings/python/examples/pytorch_loss.py:19-102) moves the emissions to
the CPU, builds one Graph per utterance inside gtn.parallel_for and copies the gradients back.
Here the [B, T, C] CUDA tensor goes straight into the C ABI (`gtnb_ctc_loss`, device pointers in,
device gradients out) on the tensor's own CUDA stream; only the B losses visit the host.
"""
import ctypes as C

import numpy as np
import torch

from . import capi

_ctx_cache = {}


def _ctx_for(device, stream):
    key = (device.index, stream)
    if key not in _ctx_cache:
        _ctx_cache[key] = capi.Ctx(device.index, stream)
    return _ctx_cache[key]


class CTCLossFunction(torch.autograd.Function):
    """loss_b = forwardScore(e_b) - forwardScore(intersect(ctcGraph(target_b), e_b))  (benchmarks/ctc.cpp:150-158).

    emissions: float32 CUDA tensor [B, T, C] (unnormalised scores are fine: the normaliser is part of the loss);
    targets: list of 1-D int sequences; returns the mean loss over the batch (reduction="mean") or the [B] vector.
    """

    @staticmethod
    def forward(ctx, emissions, targets, blank=0, reduction="mean"):
        assert emissions.is_cuda and emissions.dtype == torch.float32 and emissions.dim() == 3
        e = emissions.contiguous()
        B, T, Cn = e.shape
        stream = torch.cuda.current_stream(e.device).cuda_stream
        g = _ctx_for(e.device, stream)
        lens = np.asarray([len(t) for t in targets], np.int32)
        cat = np.ascontiguousarray(
            np.concatenate([np.asarray(t, np.int32) for t in targets]) if B else np.zeros(0, np.int32), np.int32)
        losses = np.zeros(B, np.float32)
        grad = torch.empty_like(e)
        g._check(capi.lib().gtnb_ctc_loss(
            g.h, B, T, Cn, e.data_ptr(), 1, None, cat.ctypes.data_as(capi._i32p),
            lens.ctypes.data_as(capi._i32p), int(blank), losses.ctypes.data_as(capi._f32p),
            grad.data_ptr(), 1))
        out = torch.from_numpy(losses).to(e.device)
        ctx.save_for_backward(grad)
        ctx.reduction = reduction
        ctx.B = B
        return out.mean() if reduction == "mean" else out

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        if ctx.reduction == "mean":
            scale = grad_output / ctx.B
            return grad * scale, None, None, None
        return grad * grad_output.view(-1, 1, 1), None, None, None


def ctc_loss(emissions, targets, blank=0, reduction="mean"):
    return CTCLossFunction.apply(emissions, targets, blank, reduction)

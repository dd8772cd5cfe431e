"""cuDNN fused convolution epilogues for the R-CNN (a4 glue, inference only).

After `fold_batchnorm_into_convs` every backbone convolution carries a bias, and PyTorch's eager conv adds it with a
separate full-tensor kernel, then ReLU (and the residual add) with more of them -- tools/rcnn_kernels.py showed those
elementwise passes costing more device time than the convolutions.  cuDNN can run conv + bias + ReLU and
conv + bias + residual + ReLU as ONE kernel (`torch.cudnn_convolution_relu`, `torch.cudnn_convolution_add_relu`: library
calls, like the convolutions themselves); this module routes the ResNet bottlenecks, the stem and every (Conv2d, ReLU) pair
of the heads through them.  `enable_fused_convs(model)` patches in place; CPU tensors and unusual convolutions keep the
eager path.
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as F


def _fusable(conv: torch.nn.Conv2d) -> bool:
    return isinstance(conv, torch.nn.Conv2d) and conv.padding_mode == 'zeros' and not isinstance(conv.padding, str)


class _CastCache:
    """Weights / biases in the compute dtype, converted once (the model is frozen for inference)."""

    def __init__(self, conv: torch.nn.Conv2d):
        self.conv, self.cache = conv, {}

    def get(self, dtype: torch.dtype):
        key = (dtype, self.conv.weight.data_ptr(), self.conv.weight.stride())
        hit = self.cache.get(key)
        if hit is None:
            c = self.conv
            hit = (c.weight.detach().to(dtype), None if c.bias is None else c.bias.detach().to(dtype))
            self.cache = {key: hit}
        return hit


def _compute_dtype(x: torch.Tensor) -> torch.dtype:
    return torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype


def conv_bias_relu(conv: torch.nn.Conv2d, cache: _CastCache, x: torch.Tensor) -> torch.Tensor:
    if not (x.is_cuda and _fusable(conv)) or torch.is_grad_enabled() and conv.weight.requires_grad:
        return F.relu(conv(x))
    dt = _compute_dtype(x)
    w, b = cache.get(dt)
    return torch.cudnn_convolution_relu(x.to(dt), w, b, conv.stride, conv.padding, conv.dilation, conv.groups)


def conv_bias_add_relu(conv: torch.nn.Conv2d, cache: _CastCache, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    if not (x.is_cuda and _fusable(conv)) or torch.is_grad_enabled() and conv.weight.requires_grad:
        return F.relu(conv(x) + z)
    dt = _compute_dtype(x)
    w, b = cache.get(dt)
    z = z.to(dt)
    if z.is_contiguous(memory_format=torch.channels_last) != x.is_contiguous(memory_format=torch.channels_last):
        z = z.contiguous(memory_format=torch.channels_last if x.is_contiguous(memory_format=torch.channels_last) else torch.contiguous_format)
    return torch.cudnn_convolution_add_relu(x.to(dt), w, z, 1.0, b, conv.stride, conv.padding, conv.dilation, conv.groups)


class FusedConvReLU(torch.nn.Module):
    """Conv2d followed by ReLU as one cuDNN call; stands where the Conv2d stood, the ReLU becomes an Identity."""

    def __init__(self, conv: torch.nn.Conv2d):
        super().__init__()
        self.conv = conv
        self._cache = _CastCache(conv)

    def forward(self, x):
        return conv_bias_relu(self.conv, self._cache, x)


def _bottleneck_forward(self, x):
    """torchvision.models.resnet.Bottleneck.forward with folded normalisation layers: three fused cuDNN calls."""
    out = conv_bias_relu(self.conv1, self._msq_cache[0], x)
    out = conv_bias_relu(self.conv2, self._msq_cache[1], out)
    identity = x if self.downsample is None else self.downsample(x)
    return conv_bias_add_relu(self.conv3, self._msq_cache[2], out, identity)


def enable_fused_convs(model: torch.nn.Module) -> int:
    """Patch `model` (eval mode, BatchNorm already folded) in place; returns the number of convolutions fused."""
    from torchvision.models.resnet import Bottleneck
    fused = 0
    for mod in list(model.modules()):
        if isinstance(mod, Bottleneck) and not hasattr(mod, '_msq_cache'):
            if all(isinstance(getattr(mod, n), torch.nn.Identity) for n in ('bn1', 'bn2', 'bn3')) and \
                    all(_fusable(getattr(mod, n)) for n in ('conv1', 'conv2', 'conv3')):
                mod._msq_cache = [_CastCache(mod.conv1), _CastCache(mod.conv2), _CastCache(mod.conv3)]
                mod.forward = types.MethodType(_bottleneck_forward, mod)
                fused += 3
    for parent in list(model.modules()):
        names = list(parent._modules.keys())
        ordered = isinstance(parent, torch.nn.Sequential)
        # the ResNet stem lives in an IntermediateLayerGetter (a ModuleDict run in order): conv1, bn1 (Identity), relu
        if type(parent).__name__ == 'IntermediateLayerGetter' and names[:3] == ['conv1', 'bn1', 'relu'] and \
                isinstance(parent._modules['bn1'], torch.nn.Identity) and _fusable(parent._modules['conv1']):
            parent._modules['conv1'] = FusedConvReLU(parent._modules['conv1'])
            parent._modules['relu'] = torch.nn.Identity()
            fused += 1
        if not ordered:
            continue
        for a, b in zip(names, names[1:]):
            conv, act = parent._modules[a], parent._modules[b]
            if _fusable(conv) and type(conv) is torch.nn.Conv2d and isinstance(act, torch.nn.ReLU):
                parent._modules[a] = FusedConvReLU(conv)
                parent._modules[b] = torch.nn.Identity()
                fused += 1
    return fused

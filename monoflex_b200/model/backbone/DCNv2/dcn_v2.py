"""DCNv2 modules with the reference's names, constructor signatures and state_dict keys
(model/backbone/DCNv2/dcn_v2.py:16-128). `_DCNv2.forward` calls `_ext.dcn_v2_forward` with the reference's positional
signature; inside the detector the fused NHWC tensor-core kernel is used instead (engine.Plan.dcn)."""
import math

import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _ext as _backend


class _DCNv2(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
        ctx.cfg = (_pair(weight.shape[2:4]), _pair(stride), _pair(padding), _pair(dilation), deformable_groups)
        k, s, p, d, g = ctx.cfg
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return _backend.dcn_v2_forward(input, weight, bias, offset, mask, k[0], k[1], s[0], s[1], p[0], p[1], d[0], d[1], g)

    @staticmethod
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        k, s, p, d, g = ctx.cfg
        grads = _backend.dcn_v2_backward(input, weight, bias, offset, mask, grad_output.contiguous(), k[0], k[1], s[0],
                                         s[1], p[0], p[1], d[0], d[1], g)
        gi, go, gm, gw, gb = grads
        return gi, go, gm, gw, gb, None, None, None, None


dcn_v2_conv = _DCNv2.apply


class DCNv2(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super(DCNv2, self).__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        with torch.no_grad():
            self.weight.uniform_(-stdv, stdv)
            self.bias.zero_()

    def forward(self, input, offset, mask):
        kk = self.deformable_groups * self.kernel_size[0] * self.kernel_size[1]
        assert 2 * kk == offset.shape[1] and kk == mask.shape[1]
        return dcn_v2_conv(input, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)


class DCN(DCNv2):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super(DCN, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        kk = self.deformable_groups * self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(in_channels, 3 * kk, kernel_size=self.kernel_size, stride=self.stride,
                                          padding=self.padding, bias=True)
        nn.init.zeros_(self.conv_offset_mask.weight)      # dcn_v2.py:114-116
        nn.init.zeros_(self.conv_offset_mask.bias)

    def forward(self, input):
        """Stand-alone module forward (fp32 NCHW in/out). The offset conv is torch's; the DCN op is ours."""
        out = self.conv_offset_mask(input)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        return dcn_v2_conv(input, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias, self.stride,
                           self.padding, self.dilation, self.deformable_groups)

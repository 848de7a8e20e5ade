"""Drop-in for the reference's native operator module `_ext` (model/backbone/DCNv2/src/vision.cpp:4-9), same positional
signatures as src/dcn_v2.h:9-23,48-59, backed by the C ABI of libmonoflex_b200.so.

    output = _ext.dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, deformable_group)

Inputs must be fp32 CUDA tensors; like the reference (dcn_v2_cuda.cu:60-84) anything else raises RuntimeError.
Outputs are freshly allocated tensors owned by the caller; the kernel runs asynchronously on the current stream."""
import torch

from ...._lib import call, load, stream


def _check(name, t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)
    return t.contiguous()


def dcn_v2_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                   dilation_w, deformable_group):
    x, w, b = _check("input", input), _check("weight", weight), _check("bias", bias)
    off, m = _check("offset", offset), _check("mask", mask)
    B, C, H, W = x.shape
    Co, Ck, kh_, kw_ = w.shape
    if (kh_, kw_) != (kernel_h, kernel_w):
        raise RuntimeError("Input shape and kernel shape wont match: (%d x %d vs %d x %d)." % (kernel_h, kernel_w, kh_, kw_))
    if C != Ck:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (C, Ck))
    Ho = (H + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    Wo = (W + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1
    if off.shape != (B, 2 * deformable_group * kernel_h * kernel_w, Ho, Wo) or \
            m.shape != (B, deformable_group * kernel_h * kernel_w, Ho, Wo):
        raise RuntimeError("offset/mask shape does not match the output size")
    y = torch.empty(B, Co, Ho, Wo, dtype=torch.float32, device=x.device)
    call("mf_dcn_v2_forward", x.data_ptr(), w.data_ptr(), b.data_ptr(), off.data_ptr(), m.data_ptr(), y.data_ptr(), B, C,
         H, W, Co, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, deformable_group, None, 0,
         stream())
    return y


def dcn_v2_backward(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                    dilation_h, dilation_w, deformable_group):
    """-> (grad_input, grad_offset, grad_mask, grad_weight, grad_bias), src/dcn_v2.h:48-59."""
    x, w, b = _check("input", input), _check("weight", weight), _check("bias", bias)
    off, m, gy = _check("offset", offset), _check("mask", mask), _check("grad_output", grad_output)
    B, C, H, W = x.shape
    Co = w.shape[0]
    gx, gw, gb = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
    go, gm = torch.empty_like(off), torch.empty_like(m)
    geom = (B, C, H, W, Co, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, deformable_group)
    ws_bytes = load().mf_dcn_v2_backward_workspace(*geom)
    if ws_bytes == 0:
        raise RuntimeError("dcn_v2_backward: " + load().mf_last_error().decode())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)          # grad-columns buffer + split-K partials of dW
    call("mf_dcn_v2_backward", x.data_ptr(), w.data_ptr(), b.data_ptr(), off.data_ptr(), m.data_ptr(), gy.data_ptr(),
         gx.data_ptr(), go.data_ptr(), gm.data_ptr(), gw.data_ptr(), gb.data_ptr(), B, C, H, W, Co, kernel_h, kernel_w,
         stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, deformable_group, ws.data_ptr(), ws_bytes, stream())
    return gx, go, gm, gw, gb


def dcn_v2_psroi_pooling_forward(*args):
    call("mf_dcn_v2_psroi_pooling_forward")


def dcn_v2_psroi_pooling_backward(*args):
    call("mf_dcn_v2_psroi_pooling_backward")

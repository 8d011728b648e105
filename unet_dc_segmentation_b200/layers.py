"""Single-layer entry points (dc_stem / dc_conv_tc) on torch tensors -- used by the per-layer parity tests
and the dilation / size sweep; the network itself goes through dc_forward (model.py).

Activations are NHWC bf16 CUDA tensors; weights are the packed blobs of model.pack_conv3x3 / pack_upconv.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def stem(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, dilation: int = 1) -> torch.Tensor:
    """First conv (Cin = 3) + folded BN + ReLU.  x: f32 [B,3,H,W], u8 [B,H,W] or u8 [B,H,W,3];
    weight: f32 [64,27] (co, ci*9+ky*3+kx); returns bf16 [B,H,W,64]."""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    if x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3:
        kind, (B, _, H, W) = 0, x.shape
    elif x.dtype == torch.uint8 and x.dim() == 3:
        kind, (B, H, W) = 1, x.shape
    elif x.dtype == torch.uint8 and x.dim() == 4 and x.shape[3] == 3:
        kind, (B, H, W, _) = 2, x.shape
    else:
        raise ValueError(f"unsupported stem input {x.dtype} {tuple(x.shape)}")
    out = torch.empty((B, H, W, 64), dtype=torch.bfloat16, device=x.device)
    args = _lib.StemArgs(kind, B, H, W, 64, int(dilation), x.data_ptr(), weight.data_ptr(), bias.data_ptr(),
                         out.data_ptr(), 64, 0)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().dc_stem(C.byref(args), _lib.stream_ptr(x.device)))
    return out


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, dilation: int = 1, relu: bool = True,
            cin: int | None = None, out: torch.Tensor | None = None, out_offset: int = 0, pool: bool = False,
            weight_par: torch.Tensor | None = None):
    """3x3 dilated conv (+bias, +ReLU) on tcgen05.  x: bf16 [B,H,W,S] whose channels [0,cin) are read;
    weight: bf16 [Cout, 9*cin]; out (optional): bf16 [B,H,W,S_out], written at channel out_offset.
    Returns out, or (out, pooled bf16 [B,H/2,W/2,Cout]) when pool."""
    _lib.require_cuda(x, "x")
    B, H, W, S = x.shape
    cin = S if cin is None else cin
    cout = weight.shape[0]
    if out is None:
        out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=x.device)
    pooled = torch.empty((B, H // 2, W // 2, cout), dtype=torch.bfloat16, device=x.device) if pool else None
    a = _lib.ConvArgs()
    a.kind, a.epilogue, a.relu = _lib.DC_KIND_CONV3X3, (_lib.DC_EPI_STORE_POOL if pool else _lib.DC_EPI_STORE), int(relu)
    a.B, a.H, a.W, a.Cin, a.Cout, a.dilation = B, H, W, cin, cout, int(dilation)
    a.in_, a.in_stride = x.data_ptr(), S
    a.weight, a.bias = weight.data_ptr(), bias.data_ptr()
    a.out, a.out_stride, a.out_offset = out.data_ptr(), out.shape[3], int(out_offset)
    if weight_par is not None:
        a.weight_par = weight_par.data_ptr()
    if pool:
        a.pool_out, a.pool_stride = pooled.data_ptr(), cout
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().dc_conv_tc(C.byref(a), _lib.stream_ptr(x.device)))
    return (out, pooled) if pool else out


def upconv2x2(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out: torch.Tensor | None = None,
              out_offset: int = 0) -> torch.Tensor:
    """ConvTranspose2d(k=2, s=2).  x: bf16 [B,H,W,Cin]; weight: bf16 [4*Cout, Cin]; out: bf16 [B,2H,2W,S_out]."""
    _lib.require_cuda(x, "x")
    B, H, W, cin = x.shape
    cout = weight.shape[0] // 4
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, cout), dtype=torch.bfloat16, device=x.device)
    a = _lib.ConvArgs()
    a.kind, a.epilogue, a.relu = _lib.DC_KIND_UPCONV2, _lib.DC_EPI_UPSCATTER, 0
    a.B, a.H, a.W, a.Cin, a.Cout, a.dilation = B, H, W, cin, cout, 1
    a.in_, a.in_stride = x.data_ptr(), cin
    a.weight, a.bias = weight.data_ptr(), bias.data_ptr()
    a.out, a.out_stride, a.out_offset = out.data_ptr(), out.shape[3], int(out_offset)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().dc_conv_tc(C.byref(a), _lib.stream_ptr(x.device)))
    return out


def upconv_conv3x3(x: torch.Tensor, skip: torch.Tensor, weight: torch.Tensor, bias9: torch.Tensor, relu: bool = True,
                   skip_offset: int = 0, out: torch.Tensor | None = None, out_offset: int = 0,
                   weight_skip: torch.Tensor | None = None) -> torch.Tensor:
    """ConvTranspose2d(2C, C, 2, 2) -> cat([up, skip]) -> 3x3 conv (+bias, +ReLU) as one launch (dc_conv_upfused).
    x: bf16 [B,H,W,2C]; skip: bf16 [B,2H,2W,S] read at channels [skip_offset, skip_offset + C); C = bias9.shape[1].
    C = 64: weight from model.pack_upfused; C = 128 / 256 / 512: weight, weight_skip from model.pack_upfused_wide.
    Returns bf16 [B,2H,2W,C] (or `out`, written at channel out_offset)."""
    _lib.require_cuda(x, "x")
    B, H, W, S = x.shape
    c = int(bias9.shape[1])
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, c), dtype=torch.bfloat16, device=x.device)
    a = _lib.UpfuseArgs()
    a.B, a.H, a.W = B, H, W
    a.channels = c
    if weight_skip is not None:
        a.weight_skip = weight_skip.data_ptr()
    a.x, a.x_stride = x.data_ptr(), S
    a.skip, a.skip_stride = skip.data_ptr() + 2 * int(skip_offset), skip.shape[3]
    a.weight, a.bias9, a.relu = weight.data_ptr(), bias9.data_ptr(), int(relu)
    a.out, a.out_stride, a.out_offset = out.data_ptr(), out.shape[3], int(out_offset)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().dc_conv_upfused(C.byref(a), _lib.stream_ptr(x.device)))
    return out


def conv3x3_head(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, head_w: torch.Tensor, head_b: float,
                 thresh: float, dilation: int = 1, weight_par: torch.Tensor | None = None):
    """Last layer: conv3x3 (64 -> 64) + ReLU, then 1x1 conv + sigmoid + threshold in the epilogue.
    Returns (probs f32 [B,H,W], mask u8 [B,H,W])."""
    _lib.require_cuda(x, "x")
    B, H, W, S = x.shape
    prob = torch.empty((B, H, W), dtype=torch.float32, device=x.device)
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
    a = _lib.ConvArgs()
    a.kind, a.epilogue, a.relu = _lib.DC_KIND_CONV3X3, _lib.DC_EPI_HEAD, 1
    a.B, a.H, a.W, a.Cin, a.Cout, a.dilation = B, H, W, S, 64, int(dilation)
    a.in_, a.in_stride = x.data_ptr(), S
    a.weight, a.bias = weight.data_ptr(), bias.data_ptr()
    a.head_w, a.head_b, a.thresh = head_w.data_ptr(), float(head_b), float(thresh)
    if weight_par is not None:
        a.weight_par = weight_par.data_ptr()
    a.prob_out, a.mask_out = prob.data_ptr(), mask.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().dc_conv_tc(C.byref(a), _lib.stream_ptr(x.device)))
    return prob, mask

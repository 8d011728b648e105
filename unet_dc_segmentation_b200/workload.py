"""Algorithmic work of the hot path (what roofline.achieved is computed from; DESIGN.md "Measurement").

FLOPs follow reference models/model_2.py:5-80 layer by layer; bytes follow BASELINE.md section 3.
"""
from __future__ import annotations

WIDTHS = (64, 128, 256, 512, 1024)


def _inb(n: int, d: int) -> float:
    """Fraction of the 3 taps along one axis of length n that fall inside the image at dilation d."""
    return (n + 2 * max(n - d, 0)) / (3.0 * n)


def conv_layers(H: int, W: int, dilations=(1, 2, 4, 8, 16)):
    """(name, kind, h, w, cin, cout, dilation) for the 23 layers of UNetDC at input H x W, in launch order."""
    out = []
    cin = 3
    for lvl, (c, d) in enumerate(zip(WIDTHS, dilations)):
        h, w = H >> lvl, W >> lvl
        name = f"enc{lvl + 1}" if lvl < 4 else "bottleneck"
        out.append((f"{name}.0", "conv3x3", h, w, cin, c, d))
        out.append((f"{name}.3", "conv3x3", h, w, c, c, d))
        cin = c
    for lvl in (3, 2, 1, 0):
        h, w, c = H >> lvl, W >> lvl, WIDTHS[lvl]
        out.append((f"upconv{lvl + 1}", "upconv2x2", h // 2, w // 2, 2 * c, c, 1))
        out.append((f"dec{lvl + 1}.0", "conv3x3", h, w, 2 * c, c, 1))
        out.append((f"dec{lvl + 1}.3", "conv3x3", h, w, c, c, 1))
    out.append(("out_conv", "conv1x1", H, W, 64, 1, 1))
    return out


def layer_flops(layer, in_bounds: bool = True) -> float:
    _, kind, h, w, cin, cout, d = layer
    if kind == "conv3x3":
        f = _inb(h, d) * _inb(w, d) if in_bounds else 1.0
        return 2.0 * 9 * cin * cout * h * w * f
    if kind == "upconv2x2":
        return 2.0 * 4 * cin * cout * h * w
    return 2.0 * cin * cout * h * w


def forward_flops(H: int, W: int, dilations=(1, 2, 4, 8, 16), in_bounds: bool = True) -> float:
    """FLOPs of one UNetDC forward on one H x W image (nominal: 2 x 734,976 x H x W)."""
    return sum(layer_flops(l, in_bounds) for l in conv_layers(H, W, dilations))


def forward_flops_issued(H: int, W: int, dilations=(1, 2, 4, 8, 16), fused_levels=(), in_bounds: bool = True) -> float:
    """FLOPs the kernels actually issue for one forward: a decoder level whose upconv is composed into the following
    conv (2x2 taps over the 2C-channel half-resolution input + 3x3 over the C-channel skip) runs K = 8 C + 9 C per
    output pixel and channel instead of the 2 C + 18 C of the two layers it replaces."""
    total = 0.0
    for l in conv_layers(H, W, dilations):
        name, kind, h, w, cin, cout, d = l
        lvl = int(name[6]) if name.startswith("upconv") else (int(name[3]) if name.startswith("dec") and name.endswith(".0") else 0)
        if lvl in fused_levels and kind == "upconv2x2":
            continue                                         # rides in dec{lvl}.0
        if lvl in fused_levels:
            f = _inb(h, 1) * _inb(w, 1) if in_bounds else 1.0
            total += 2.0 * cout * h * w * (4 * cin + 9 * cout * f)      # cin = 2 C here: x taps 4 x 2C, skip taps 9 x C
        else:
            total += layer_flops(l, in_bounds)
    return total


def launch_flops(H: int, W: int, dilations=(1, 2, 4, 8, 16), in_bounds: bool = True, fused_level1: bool = False,
                 fused_levels=()):
    """FLOPs per kernel launch of dc_forward (22 launches: out_conv is fused into the last one; one fewer per decoder
    level whose upconv rides in the following conv -- ``fused_level1`` / ``fused_levels`` -- counted with the FLOPs of
    the two layers it replaces: the composed form executes K = 17 C instead of 20 C per output pixel and channel)."""
    ls = conv_layers(H, W, dilations)
    fl = [layer_flops(l, in_bounds) for l in ls]
    fl[-2] += fl[-1]
    names, fl = [l[0] for l in ls[:-1]], fl[:-1]
    for lvl in sorted(set(fused_levels) | ({1} if fused_level1 else set())):
        i = names.index(f"upconv{lvl}")
        assert names[i + 1] == f"dec{lvl}.0"
        names[i:i + 2] = [f"upconv{lvl}+dec{lvl}.0"]
        fl[i:i + 2] = [fl[i] + fl[i + 1]]
    return names, fl


# bytes per pixel
ROLLING_BALL_BYTES_PER_PX = 2      # u8 read + u8 write, one plane for grayscale input
LABEL_BYTES_PER_PX = 5             # 1 B mask read + 4 B int32 label write
STATS_BYTES_PER_PX = 4             # int32 label read
STATS_BYTES_PER_DROPLET = 56       # i32 label, i64 area, 5 x f64

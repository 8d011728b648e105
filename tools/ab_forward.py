"""A/B of forward schedules in ONE process on the same box (the boxes differ by more than the effects measured):
interleaved rounds of N forwards per variant at the bench shape, CUDA events, median per variant."""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from unet_dc_segmentation_b200 import UNetDC                                    # noqa: E402
from unet_dc_segmentation_b200.synth import calibrated_state_dict              # noqa: E402

B, S = 32, 1024
dev = torch.device("cuda:0")
sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=1)
# name -> (parity_layers, fuse_level1, fuse_levels)
variants = {"enc1.3 + dec1.3 by parity class": (("enc1", "dec1"), True, (2, 3, 4)), "dec1.3 only": (("dec1",), True, (2, 3, 4)),
            "enc1.3 only": (("enc1",), True, (2, 3, 4)), "neither": ((), True, (2, 3, 4)),
            "fused: level 1 only": (("enc1", "dec1"), True, ()), "fused: none (22 launches)": ((), False, ())}
if len(sys.argv) > 1:
    variants = {k: v for k, v in variants.items() if any(k.startswith(a) for a in sys.argv[1].split(","))}
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
models = {}
for name, (layers, f1, fl) in variants.items():
    m = UNetDC(3, 1)
    m.load_state_dict(sd)
    m.parity_layers, m.fuse_level1, m.fuse_levels = layers, f1, fl
    models[name] = m.to(dev).eval()
frames = torch.randint(0, 256, (B, S, S), dtype=torch.uint8, device=dev)
times = {k: [] for k in variants}
for k, m in models.items():
    for _ in range(3):
        m.predict_u8(frames, 0.3)
torch.cuda.synchronize()
for rnd in range(ROUNDS):
    for k, m in models.items():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            m.predict_u8(frames, 0.3)
        b.record()
        torch.cuda.synchronize()
        times[k].append(a.elapsed_time(b) / 5)
for k, v in times.items():
    print(f"{k:36s} median {statistics.median(v):.3f} ms  (rounds: {' '.join(f'{x:.2f}' for x in v)})")

import cProfile, pstats, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, '.')
import torch
from PIL import Image
from unet_dc_segmentation_b200 import cli
from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
n, size, batch, img = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
d = Path(tempfile.mkdtemp()); (d/'in').mkdir()
for i in range(n): Image.fromarray(synthetic_image(size, i)).save(d/'in'/f'f{i:04d}.png')
torch.save(calibrated_state_dict(seed=0), d/'ckpt.pth')
common = ['--img_dir', str(d/'in'), '--ckpt_path', str(d/'ckpt.pth'), '--batch', str(batch), '--img_size', str(img), '--skip_excel', '--skip_histogram', '--px_per_micron', '3.45']
cli.main(common + ['--out_dir', str(d/'o0')])
pr = cProfile.Profile(); pr.enable(); t0=time.perf_counter()
cli.main(common + ['--out_dir', str(d/'o1')]); torch.cuda.synchronize()
pr.disable(); print('wall', time.perf_counter()-t0)
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)

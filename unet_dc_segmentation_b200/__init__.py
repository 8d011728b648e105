"""unet_dc_segmentation_b200 -- the droplet-quantification hot path of malani86/unet-DC-segmentation
(quantify_droplets_batch.py) as hand-written sm_100a CUDA kernels behind the reference's own call surface.

    UNetDC / UNet                  <- models/model_2.py:5 / models/model.py:7      (model.py)
    rolling_ball_correction_rgb    <- utils/data_loader.py:11                       (morphology.py)
    quantify                       <- quantify_droplets_batch.py:81                 (quantify.py)
    load_model / preprocess / run_batch / main  <- quantify_droplets_batch.py       (cli.py)
    overlay_stencil_device         <- findContours + drawContours, qdb:74-79        (overlay.py)
    generate_roi_mask / get_targets / density_maps  <- quantify_pipline.py:44,61,93 (density.py)
    DropletPipeline                fused batched device path                        (pipeline.py)

All compute goes through lib/libunetdc_b200.so (C ABI: include/unetdc_b200.h).  There is no CPU or
PyTorch fallback: without the library or an sm_100 GPU the calls raise.
"""
from .model import UNet, UNetDC  # noqa: F401
from .morphology import resize_linear_u8_device, rolling_ball_correction_rgb, rolling_ball_device  # noqa: F401
from .overlay import draw_overlay, overlay_stencil_device  # noqa: F401
from .pipeline import DropletPipeline  # noqa: F401
from .quantify import label_stats_device, quantify, quantify_arrays  # noqa: F401

__all__ = ["UNetDC", "UNet", "rolling_ball_correction_rgb", "rolling_ball_device", "resize_linear_u8_device", "quantify", "quantify_arrays",
           "label_stats_device", "DropletPipeline", "overlay_stencil_device", "draw_overlay"]

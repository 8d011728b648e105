"""GPU: whole-network and whole-path parity.

Tolerance (north_star: "probabilities must agree within a stated bf16 tolerance; mask disagreements are
counted only at pixels within tolerance of prob_thresh").  The CUDA path stores activations and weights in
bf16 and accumulates in fp32; the reference is fp32 throughout.  Two checks:

 1. against the fp32 reference (golden / oracle):  max |dp| <= PROB_TOL = 0.08 and mean |dp| <= MEAN_TOL = 0.004.
    A CPU emulation that rounds to bf16 at exactly the kernel's storage points (oracle.unetdc_forward(...,
    emulate_bf16=True)) differs from fp32 by max 0.05-0.07 / mean 0.0008 on this checkpoint: 23 random-weight
    layers leave ~2 % relative noise in the last feature map and the fitted head (logit range -6..+4) turns
    0.3-0.5 of logit noise into <= 0.07 of probability at the steepest pixels.  Masks may differ from the
    reference's only where |p_ref - thresh| <= PROB_TOL.
 2. against that bf16 emulation: what is left is fp32 accumulation order (tensor core vs CPU) flipping
    single bf16 roundings: max |dp| <= EMU_TOL = 0.02, mean |dp| <= EMU_MEAN_TOL = 0.001 (measured on B200:
    max 0.013, mean 0.0003).  This is the check that says the kernels compute the network they claim to.

THE STATED TOLERANCE: |dp| <= 0.03 on the survey's checkpoint -- default init + BatchNorm
calibration + an out_conv bias shift (``calibrated_state_dict(fit_head=False)``, logit std ~0.4) -- and masks may
differ from the reference's only where |p_ref - thresh| <= 0.03: test_survey_checkpoint_stated_tolerance.  The bench
and most tests use the fitted-head checkpoint (``fit_head=True``: out_conv is least-squares fitted so that the masks
look like droplets; its logits span -6..+4, ~10x the gain), for which the same bf16 noise in the last feature map
becomes up to 0.06-0.07 of probability at the steepest pixels: PROB_TOL = 0.08 applies to THAT checkpoint only, and
bench.py reports how many mask pixels / droplets actually differ from the fp32 reference on the bench frames.

Droplet tables are bit-exact GIVEN THE SAME MASK, so every table check feeds the oracle the kernel's own mask."""
import numpy as np
import pytest

import oracle
from conftest import assert_table_equal, load_golden

pytestmark = pytest.mark.gpu


def _fused(sd):
    """Composed weights of the decoder levels the CUDA forward runs as one launch each by default (UNetDC.fuse_level1,
    UNetDC.fuse_levels), for the emulation."""
    from unet_dc_segmentation_b200.model import fused_level_blobs
    cpu = {k: v.detach().float().cpu() for k, v in sd.items() if k.startswith(("dec", "upconv"))}
    return {lvl: fused_level_blobs(cpu, lvl) for lvl in (1, 2, 3, 4)}


PROB_TOL = 0.08
MEAN_TOL = 0.004
EMU_MEAN_TOL = 0.001
EMU_TOL = 0.02


SURVEY_TOL = 0.03          # the stated bf16 tolerance, on the survey's checkpoint
SURVEY_MEAN_TOL = 0.002


def _check_probs(got, ref_fp32, emu=None, what=""):
    err = np.abs(got - ref_fp32)
    msg = f"{what}: vs fp32 max {err.max():.4f} mean {err.mean():.5f}"
    if emu is not None:
        e2 = np.abs(got - emu)
        msg += f"; vs bf16 emulation max {e2.max():.4f} mean {e2.mean():.5f}"
    print(msg)
    assert err.max() <= PROB_TOL and err.mean() <= MEAN_TOL, msg
    if emu is not None:
        assert e2.mean() <= EMU_MEAN_TOL and e2.max() <= EMU_TOL, msg


def _model(cls_name, sd, device):
    import unet_dc_segmentation_b200 as pkg
    m = getattr(pkg, cls_name)(3, 1)
    m.load_state_dict(sd)
    return m.to(device).eval()


@pytest.mark.parametrize("tag,cls", [("unetdc", "UNetDC"), ("unet", "UNet")])
def test_forward_vs_reference_golden(cuda_device, tag, cls):
    import torch
    from unet_dc_segmentation_b200.synth import calibrated_state_dict
    g = load_golden("forward.npz")
    dil = tuple(int(v) for v in g[f"{tag}/dilations"])
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2, dilations=dil)
    m = _model(cls, sd, cuda_device)
    x = torch.from_numpy(np.repeat(g[f"{tag}/images"][:, None], 3, 1).astype(np.float32) / 255.0)
    y = m(x.to(cuda_device))
    assert y.shape == (2, 1, 64, 64) and y.dtype == torch.float32
    emu = oracle.unetdc_forward(sd, x, dil, emulate_bf16=True, fused_levels=_fused(sd)).numpy()
    _check_probs(y.cpu().numpy(), g[f"{tag}/probs"], emu, tag)


@pytest.mark.parametrize("B,H,W", [(1, 16, 16), (2, 128, 128), (1, 48, 80)])
def test_forward_vs_oracle_sizes(cuda_device, B, H, W):
    import torch
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    m = _model("UNetDC", sd, cuda_device)
    imgs = np.stack([synthetic_image(max(H, W), 300 + b)[:H, :W] for b in range(B)])
    x = torch.from_numpy(np.repeat(imgs[:, None], 3, 1).astype(np.float32) / 255.0)
    want = oracle.unetdc_forward(sd, x).numpy()
    emu = oracle.unetdc_forward(sd, x, emulate_bf16=True, fused_levels=_fused(sd)).numpy()
    emu_gray = oracle.unetdc_forward(sd, x, emulate_bf16=True, fused_levels=_fused(sd), gray_input=True).numpy()
    _check_probs(m(x.to(cuda_device)).cpu().numpy(), want, emu, f"f32 NCHW {B}x{H}x{W}")
    # u8 entry points (the /255 happens in the first kernel)
    mask_g, prob_g = m.predict_u8(torch.from_numpy(imgs).to(cuda_device), 0.3, return_prob=True)
    _check_probs(prob_g.cpu().numpy(), want, emu_gray, f"u8 gray {B}x{H}x{W}")
    rgb = torch.from_numpy(np.repeat(imgs[..., None], 3, -1)).to(cuda_device)
    mask_c, prob_c = m.predict_u8(rgb, 0.3, return_prob=True)
    # grayscale frames fold the three (identical) input channels into one K=9 tap set before the bf16
    # rounding of the weights; the RGB entry rounds 27 weights separately: two bf16 roundings of one fp32 layer
    _check_probs(prob_c.cpu().numpy(), want, None, f"u8 HWC {B}x{H}x{W}")
    assert torch.equal(mask_g.cpu(), (prob_g[:, 0].cpu() > 0.3).to(torch.uint8))


def test_forward_with_level1_unfused(cuda_device):
    """UNetDC.fuse_level1 = False: upconv1 and dec1.0 as two launches (`up` stored in bf16), held to the emulation of
    THAT schedule; the two schedules agree with each other within the bf16 noise of one stored tensor."""
    import torch
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    imgs = np.stack([synthetic_image(96, 410 + b)[:80, :96] for b in range(2)])
    x = torch.from_numpy(np.repeat(imgs[:, None], 3, 1).astype(np.float32) / 255.0)
    want = oracle.unetdc_forward(sd, x).numpy()
    m = _model("UNetDC", sd, cuda_device)
    y_fused = m(x.to(cuda_device)).cpu().numpy()
    m.fuse_level1 = m.parity_level1 = False          # the generic kernels for enc1.3 / dec1.3 as well
    m.fuse_levels = ()
    m.invalidate()
    assert m.num_launches() == 22
    y = m(x.to(cuda_device)).cpu().numpy()
    _check_probs(y, want, oracle.unetdc_forward(sd, x, emulate_bf16=True).numpy(), "level 1 unfused")
    assert np.abs(y - y_fused).max() <= PROB_TOL


@pytest.mark.parametrize("layers,levels", [(("enc1",), (2,)), (("dec1",), (3, 4)), ((), (2, 3, 4)), (("enc1", "dec1"), ())])
def test_forward_schedule_switches(cuda_device, layers, levels):
    """Every combination of the schedule switches (UNetDC.parity_layers, UNetDC.fuse_levels) is the same network: each
    is held to the emulation of ITS schedule."""
    import torch
    from unet_dc_segmentation_b200.model import fused_level_blobs
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    imgs = np.stack([synthetic_image(96, 510 + b)[:64, :96] for b in range(2)])
    x = torch.from_numpy(np.repeat(imgs[:, None], 3, 1).astype(np.float32) / 255.0)
    m = _model("UNetDC", sd, cuda_device)
    m.parity_layers, m.fuse_levels = layers, levels
    m.invalidate()
    assert m.num_launches() == 21 - len(levels)
    cpu = {k: v.detach().float().cpu() for k, v in sd.items() if k.startswith(("dec", "upconv"))}
    emu = oracle.unetdc_forward(sd, x, emulate_bf16=True, fused_levels={l: fused_level_blobs(cpu, l) for l in (1,) + tuple(levels)})
    _check_probs(m(x.to(cuda_device)).cpu().numpy(), oracle.unetdc_forward(sd, x).numpy(), emu.numpy(), f"{layers} {levels}")


def test_module_contract(cuda_device):
    import torch
    import unet_dc_segmentation_b200 as pkg
    m = pkg.UNetDC(3, 1)
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 3, 16, 16))                        # CPU tensors: no fallback
    m = m.to(cuda_device)
    with pytest.raises(RuntimeError):
        m.train()(torch.zeros(1, 3, 16, 16, device=cuda_device))   # eval-mode only
    with pytest.raises(ValueError):
        m.eval()(torch.zeros(1, 3, 24, 16, device=cuda_device))    # H, W multiples of 16
    assert m.num_launches() == 18          # stem + 17 conv3x3 (the four upconvs ride in dec{l}.0)


def test_whole_path_vs_reference_golden(cuda_device):
    """preprocess -> forward -> threshold -> quantify at native size against the golden run of the
    reference functions (tests/golden/end_to_end.npz)."""
    import torch
    from unet_dc_segmentation_b200 import DropletPipeline
    from unet_dc_segmentation_b200.synth import calibrated_state_dict
    g = load_golden("end_to_end.npz")
    sd = calibrated_state_dict(seed=0, calib_size=64, n_calib=2)
    m = _model("UNetDC", sd, cuda_device)
    pipe = DropletPipeline(m, background_radius=50, prob_thresh=0.3, min_area=1, px_per_micron=3.45)
    imgs = torch.from_numpy(g["images"]).to(cuda_device)
    res = pipe.run_device(imgs, return_prob=True, want_labels=True)
    probs = res.probs.cpu().numpy()
    _check_probs(probs, g["probs"], None, "whole path")
    masks = res.masks.cpu().numpy()
    tables = res.tables.to_host()
    n_far_mismatch = 0
    for i in range(2):
        want_mask = np.unpackbits(g[f"mask{i}"])[: 96 * 96].reshape(96, 96)
        far = np.abs(g["probs"][i, 0] - 0.3) > PROB_TOL
        n_far_mismatch += int(((masks[i] != want_mask) & far).sum())
        labels, cols = oracle.quantify_arrays(masks[i], 1, 3.45)        # same mask -> bit-exact table
        cols["n"] = len(cols["label"])
        assert_table_equal(tables[i], cols, f"image {i}")
        np.testing.assert_array_equal(res.tables.labels[i].cpu().numpy(), labels)
    assert n_far_mismatch == 0
    # host entry: same masks and tables with the H2D / D2H inside the call
    masks_h, tables_h = pipe.run_host(torch.from_numpy(g["images"]).pin_memory())
    np.testing.assert_array_equal(masks_h, masks)
    for i in range(2):
        for c in tables[i]:
            np.testing.assert_array_equal(tables_h[i][c], tables[i][c])
    # streaming host entry (three streams, two slots): five batches, alternating two inputs, results in order
    flipped = torch.from_numpy(np.ascontiguousarray(g["images"][:, ::-1, :])).pin_memory()
    want_f_masks, want_f_tables = pipe.run_host(flipped)
    want_f_masks = want_f_masks.copy()
    seq = [torch.from_numpy(g["images"]).pin_memory(), flipped]
    n = 0
    for k, (mk, tb) in enumerate(pipe.run_host_pipelined(seq[k % 2] for k in range(5))):
        wm, wt = (masks, tables) if k % 2 == 0 else (want_f_masks, want_f_tables)
        np.testing.assert_array_equal(mk, wm, err_msg=f"batch {k}")
        for i in range(2):
            assert set(tb[i]) == set(wt[i])
            for c in wt[i]:
                np.testing.assert_array_equal(tb[i][c], wt[i][c], err_msg=f"batch {k} image {i} column {c}")
        n += 1
    assert n == 5
    assert list(pipe.run_host_pipelined(iter(()))) == []
    # results are owned by the caller: holding ALL of them (not just the latest) must keep every batch intact
    held = list(pipe.run_host_pipelined(seq[k % 2] for k in range(5)))
    for k, (mk, tb) in enumerate(held):
        np.testing.assert_array_equal(mk, masks if k % 2 == 0 else want_f_masks, err_msg=f"held batch {k}")
    assert not np.shares_memory(held[0][0], held[2][0])
    # CUDA-graph replay of the per-batch launches (one graph per buffer slot, captured on the slot's second use)
    gpipe = DropletPipeline(m, background_radius=50, prob_thresh=0.3, min_area=1, px_per_micron=3.45, use_graphs=True)
    for k, (mk, tb) in enumerate(gpipe.run_host_pipelined(seq[k % 2] for k in range(7))):
        wm, wt = (masks, tables) if k % 2 == 0 else (want_f_masks, want_f_tables)
        np.testing.assert_array_equal(mk, wm, err_msg=f"graph batch {k}")
        for i in range(2):
            for c in wt[i]:
                np.testing.assert_array_equal(tb[i][c], wt[i][c], err_msg=f"graph batch {k} image {i} column {c}")
    assert all(sl["graph"] is not None for sl in gpipe._slots)
    # tables archived on the device (sharded jobs gather them elsewhere): same rows, batch after batch
    from unet_dc_segmentation_b200.quantify import alloc_tables
    archive = alloc_tables(10, pipe.capacity, True, cuda_device)
    got_masks = [mk for mk, tb in pipe.run_host_pipelined((seq[k % 2] for k in range(5)), tables_archive=archive) if tb is None]
    assert len(got_masks) == 5
    counts, rows = archive.compact_rows()
    rows = rows.cpu().numpy()
    pos = 0
    for k in range(5):
        wt = tables if k % 2 == 0 else want_f_tables
        for i in range(2):
            n = len(wt[i]["label"])
            assert int(counts[2 * k + i]) == n
            np.testing.assert_array_equal(rows[pos:pos + n, 0].view(np.int64), wt[i]["area"])
            for j, c in enumerate(["equivalent_diameter", "centroid-0", "centroid-1", "area_sqmicron", "eq_diam_micron"]):
                np.testing.assert_array_equal(rows[pos:pos + n, j + 1], wt[i][c], err_msg=f"archive batch {k} image {i} {c}")
            pos += n
    assert pos == rows.shape[0]


def test_config1_eight_256_images(cuda_device):
    """BASELINE config 1 (kernel-parity variant, identity resizes): 8 x 256^2, batch 8, thresh 0.3, min_area 1."""
    import torch
    from unet_dc_segmentation_b200 import DropletPipeline
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0)
    m = _model("UNetDC", sd, cuda_device)
    imgs = np.stack([synthetic_image(256, i) for i in range(8)])
    probs_ref, masks_ref, _ = oracle.run_path(sd, [np.repeat(im[:, :, None], 3, 2) for im in imgs], radius=50,
                                              prob_thresh=0.3, min_area=1, px_per_um=None, use_cv2=True)
    pipe = DropletPipeline(m, 50, 0.3, 1, None)
    res = pipe.run_device(torch.from_numpy(imgs).to(cuda_device), return_prob=True)
    probs = res.probs[:, 0].cpu().numpy()
    masks = res.masks.cpu().numpy()
    _check_probs(probs, probs_ref, None, "config 1")
    far = np.abs(probs_ref - 0.3) > PROB_TOL
    assert int(((masks != masks_ref) & far).sum()) == 0
    tables = res.tables.to_host()
    for i in range(8):
        _, cols = oracle.quantify_arrays(masks[i], 1, None)
        cols["n"] = len(cols["label"])
        assert_table_equal(tables[i], cols, f"image {i}")


@pytest.mark.parametrize("dil", [(1, 2, 2, 4, 4), (2, 4, 8, 16, 32), (3, 1, 5, 2, 7)])
def test_forward_other_dilation_sets(cuda_device, dil):
    """BASELINE config 5 sweeps dilation sets: each moves layers between kernel families (halo regions for dilation
    <= 4 at every level, the per-tap kernel above that, dilation 32 on an 8x8 bottleneck = centre tap only)."""
    import torch
    import unet_dc_segmentation_b200 as pkg
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=1, calib_size=64, n_calib=2, dilations=dil)
    cls = type("UNetDCx", (pkg.UNetDC,), {"dilations": dil})
    m = cls(3, 1)
    m.load_state_dict(sd)
    m = m.to(cuda_device).eval()
    imgs = np.stack([synthetic_image(128, 700 + b) for b in range(2)])
    x = torch.from_numpy(np.repeat(imgs[:, None], 3, 1).astype(np.float32) / 255.0)
    want = oracle.unetdc_forward(sd, x, dil).numpy()
    emu = oracle.unetdc_forward(sd, x, dil, emulate_bf16=True, fused_levels=_fused(sd)).numpy()
    _check_probs(m(x.to(cuda_device)).cpu().numpy(), want, emu, f"dilations {dil}")


def test_survey_checkpoint_stated_tolerance(cuda_device):
    """SURVEY.md 8(d)'s checkpoint and bar: |dp| <= 0.03 against the fp32 reference path (cv2 rolling ball + the
    reference network), mask mismatches only within 0.03 of prob_thresh, tables bit-exact given the kernel's mask."""
    import torch
    from unet_dc_segmentation_b200 import DropletPipeline
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0, fit_head=False)
    m = _model("UNetDC", sd, cuda_device)
    imgs = np.stack([synthetic_image(256, 40 + i) for i in range(4)])
    probs_ref, masks_ref, _ = oracle.run_path(sd, [np.repeat(im[:, :, None], 3, 2) for im in imgs], radius=50,
                                              prob_thresh=0.3, min_area=1, px_per_um=None, use_cv2=True)
    pipe = DropletPipeline(m, 50, 0.3, 1, None, capacity=32768)
    res = pipe.run_device(torch.from_numpy(imgs).to(cuda_device), return_prob=True)
    probs = res.probs[:, 0].cpu().numpy()
    masks = res.masks.cpu().numpy()
    err = np.abs(probs - probs_ref)
    diff = masks != masks_ref
    print(f"survey checkpoint: max |dp| {err.max():.4f} mean {err.mean():.5f}; mask pixels differing {int(diff.sum())} "
          f"({100.0 * diff.mean():.3f} %)")
    assert err.max() <= SURVEY_TOL and err.mean() <= SURVEY_MEAN_TOL
    assert int((diff & (np.abs(probs_ref - 0.3) > SURVEY_TOL)).sum()) == 0
    tables = res.tables.to_host()
    for i in range(4):
        _, cols = oracle.quantify_arrays(masks[i], 1, None)
        cols["n"] = len(cols["label"])
        assert_table_equal(tables[i], cols, f"image {i}")


@pytest.mark.parametrize("cin,cout", [(1, 1), (4, 2), (3, 3), (7, 1)])
def test_other_channel_counts(cuda_device, cin, cout):
    """UNetDC(in_channels, out_channels) (models/model_2.py:6,10,32) beyond the (3, 1) the inference script builds:
    enc1.0 then runs as an ordinary 3x3 layer on the zero-padded input and out_conv as a 1x1 kernel (csrc/generic.cu)."""
    import torch
    import unet_dc_segmentation_b200 as pkg
    torch.manual_seed(cin * 10 + cout)
    m = pkg.UNetDC(cin, cout)
    for name, mod in m.named_modules():                      # non-trivial BatchNorm statistics, as a trained net has
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.uniform_(-0.2, 0.2)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.6)
            mod.bias.data.uniform_(-0.1, 0.1)
    m.out_conv.weight.data.mul_(25.0)                        # logits of a few units: a wrong channel would show
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    assert sd["enc1.0.weight"].shape == (64, cin, 3, 3) and sd["out_conv.weight"].shape == (cout, 64, 1, 1)
    m = m.to(cuda_device).eval()
    x = torch.rand(2, cin, 48, 64)
    y = m(x.to(cuda_device))
    assert tuple(y.shape) == (2, cout, 48, 64)
    want = oracle.unetdc_forward(sd, x).numpy()
    emu = oracle.unetdc_forward(sd, x, emulate_bf16=True, fused_levels=_fused(sd), round_last=(cout != 1)).numpy()
    _check_probs(y.cpu().numpy(), want, emu if (cin, cout) != (3, 1) else None, f"UNetDC({cin},{cout})")
    assert float(want.std()) > 0.05, "degenerate test: the reference output is flat"
    assert m.num_launches() == 18 + (cin != 3) + (cout != 1)
    if cin != 3:
        with pytest.raises(ValueError):
            m.predict_u8(torch.zeros(1, 48, 64, dtype=torch.uint8, device=cuda_device), 0.3)

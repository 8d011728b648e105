"""GPU: repeat-run determinism of every stage (compute-sanitizer is not available on this pool, so races in the
union-find / cluster-barrier / double-buffered shared-memory code are hunted by repetition instead): the same inputs
through the same kernels many times must give bit-identical outputs, and batched results must not depend on how the
frames were grouped or streamed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _masks(n, size, seed):
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n):
        kind = i % 4
        if kind == 0:
            m = (rs.rand(size, size) < rs.uniform(0.2, 0.7)).astype(np.uint8)
        elif kind == 1:
            from unet_dc_segmentation_b200.synth import synthetic_mask
            m = synthetic_mask(size, int(rs.randint(50, 600)), seed=seed * 1000 + i)
        elif kind == 2:
            m = np.zeros((size, size), np.uint8)
            m[::2, :] = 1
            m[1::2, rs.randint(0, size, size // 2)] = 1            # long serpentine-like merges
        else:
            m = (np.add.outer(np.arange(size), np.arange(size)) % 2).astype(np.uint8)
            m[rs.randint(0, size, 40), rs.randint(0, size, 40)] ^= 1
        out.append(m)
    return np.stack(out)


def test_label_stats_repeatable_on_1000_masks(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import label_stats_device
    for chunk in range(4):
        masks = torch.from_numpy(_masks(256, 192, chunk)).to(cuda_device)           # 1024 masks in four batches
        ref = None
        for rep in range(6):
            t = label_stats_device(masks, 1 + (chunk % 2) * 3, 3.45, capacity=192 * 96, want_labels=True)
            counts = t.counts.cpu().numpy()
            valid = torch.arange(t.capacity, device=cuda_device)[None, :] < t.counts[:, None]      # rows in use per image
            cur = [t.counts.clone(), t.labels.clone()] + [torch.where(valid, c.view(torch.int64), 0) for c in
                                                          (t.area, t.centroid0, t.centroid1, t.eq_diam, t.diam_um)]
            assert counts.max() <= t.capacity
            if ref is None:
                ref = cur
            else:
                for a, b in zip(ref, cur):
                    assert torch.equal(a, b), f"chunk {chunk} rep {rep}"


def test_rolling_ball_and_overlay_repeatable(cuda_device):
    import torch
    from unet_dc_segmentation_b200 import overlay_stencil_device, rolling_ball_device
    rs = np.random.RandomState(5)
    frames = torch.from_numpy(rs.randint(0, 256, (24, 300, 280)).astype(np.uint8)).to(cuda_device)
    masks = torch.from_numpy(_masks(24, 300, 9)[:, :, :280].copy()).to(cuda_device)
    r0, o0 = rolling_ball_device(frames, 50).clone(), overlay_stencil_device(masks).clone()
    for rep in range(10):
        assert torch.equal(rolling_ball_device(frames, 50), r0), f"rolling ball rep {rep}"
        assert torch.equal(overlay_stencil_device(masks), o0), f"overlay rep {rep}"


def test_pipeline_repeatable_and_grouping_independent(cuda_device):
    """32 frames of 512^2: device path ten times, then the streamed host path with the frames regrouped into
    batches of 8 / 5 / 32 -- every frame's mask and table must be the same each time."""
    import torch
    from unet_dc_segmentation_b200 import DropletPipeline, UNetDC
    from unet_dc_segmentation_b200.synth import calibrated_state_dict, synthetic_image
    sd = calibrated_state_dict(seed=0)
    m = UNetDC(3, 1)
    m.load_state_dict(sd)
    m = m.to(cuda_device).eval()
    base = np.stack([synthetic_image(512, 60 + i) for i in range(4)])
    frames = np.concatenate([np.roll(base, 29 * r, axis=2) for r in range(8)])
    pipe = DropletPipeline(m, 50, 0.3, 1, 3.45)
    dev = torch.from_numpy(frames).to(cuda_device)
    first = pipe.run_device(dev)
    masks0, tabs0 = first.masks.clone(), first.tables.to_host()
    for rep in range(10):
        r = pipe.run_device(dev)
        assert torch.equal(r.masks, masks0), f"rep {rep}"
        for a, b in zip(r.tables.to_host(), tabs0):
            for c in b:
                np.testing.assert_array_equal(a[c], b[c], err_msg=f"rep {rep} column {c}")
    m0 = masks0.cpu().numpy()
    for bs in (8, 5, 32):
        groups = [frames[i:i + bs] for i in range(0, 32, bs)]
        k = 0
        for mk, tb in DropletPipeline(m, 50, 0.3, 1, 3.45).run_host_pipelined(iter(groups), cuda_device):
            for j in range(len(tb)):
                np.testing.assert_array_equal(mk[j], m0[k], err_msg=f"batch size {bs}, frame {k}")
                for c in tabs0[k]:
                    np.testing.assert_array_equal(tb[j][c], tabs0[k][c], err_msg=f"batch size {bs}, frame {k}, column {c}")
                k += 1
        assert k == 32

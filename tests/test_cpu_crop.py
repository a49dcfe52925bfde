"""CPU checks of the visual data-pipeline oracle (oracle/crop_oracle.py, test infrastructure): bit-exact with
torchvision's PIL backend run live, with the committed golden digests, and the host-side random draws consume
torch's RNG exactly like the reference transform (dataset/CramedDataset.py:76-89)."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle.crop_oracle import crop_resize_flip_normalize, precompute_coeffs  # noqa: E402


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "crop_golden.npz"))


def test_oracle_matches_golden_digests():
    from make_crop_golden import image_of
    g = _golden()
    for k, (H, W, i, j, h, w, flip) in enumerate(g["cases"].tolist()):
        got = np.ascontiguousarray(crop_resize_flip_normalize(image_of(k, H, W), i, j, h, w, flip))
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(g["digests"][k]), k
        assert np.array_equal(got.reshape(-1)[:: got.size // 16][:16], g["probes"][k])


def test_oracle_matches_torchvision_live():
    from make_crop_golden import reference
    rs = np.random.RandomState(7)
    for _ in range(12):
        H, W = int(rs.randint(20, 400)), int(rs.randint(20, 400))
        h, w = int(rs.randint(1, H + 1)), int(rs.randint(1, W + 1))
        i, j, flip = int(rs.randint(0, H - h + 1)), int(rs.randint(0, W - w + 1)), int(rs.randint(2))
        img = rs.randint(0, 256, size=(H, W, 3), dtype=np.uint8)
        ref = reference(img, i, j, h, w, flip)
        got = crop_resize_flip_normalize(img, i, j, h, w, flip)
        assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), (H, W, i, j, h, w, flip)


def test_coefficients_sum_to_one_in_fixed_point():
    for in_size in (1, 57, 224, 360, 480, 900):
        bounds, kk = precompute_coeffs(in_size, 224)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= in_size).all()
        assert np.abs(kk.sum(1) - (1 << 22)).max() <= kk.shape[1]  # rounding of each coefficient


@pytest.mark.parametrize("mode", ["train", "test"])
def test_host_draws_follow_the_reference_transform(mode):
    """draw_frame_params consumes torch's global RNG exactly like the reference's Compose([...]) and, fed to
    the oracle, reproduces its output bit for bit."""
    from gdl_b200.datapipe import draw_frame_params
    from gdl_b200.synthetic import _transform, synth_image
    for seed in range(4):
        img = synth_image("crop/%d" % seed, (480, 360))
        torch.manual_seed(seed)
        ref = _transform(mode)(img)
        state = torch.get_rng_state()
        torch.manual_seed(seed)
        p = draw_frame_params(5, 360, 480, mode)
        assert torch.equal(torch.get_rng_state(), state)
        assert p[0] == 5
        got = crop_resize_flip_normalize(np.asarray(img), *p[1:])
        assert np.array_equal(ref.numpy().view(np.uint32), got.view(np.uint32))


def test_device_dataset_delivers_the_reference_batch():
    """synthetic.SyntheticCramedDevice (crop boxes + a uint8 frame store) describes exactly the batch that
    SyntheticCramed (the reference's per-item torchvision transform) produces from the same RNG state."""
    import argparse
    import pickle
    from gdl_b200.synthetic import SyntheticCramed, SyntheticCramedDevice
    args = argparse.Namespace(fps=2)
    for mode in ("train", "test"):
        ref_ds = SyntheticCramed(args, mode, 3)
        dev_ds = SyntheticCramedDevice(args, mode, 3)
        store = dev_ds.frame_store().numpy()
        assert store.shape == (6, 360, 480, 3)
        for idx in range(3):
            torch.manual_seed(20 + idx)
            np.random.seed(20 + idx)
            spec_r, images, label_r = ref_ds[idx]
            state = (torch.get_rng_state(), np.random.get_state()[1].copy())
            torch.manual_seed(20 + idx)
            np.random.seed(20 + idx)
            spec_d, params, label_d = dev_ds[idx]
            assert torch.equal(torch.get_rng_state(), state[0]) and np.array_equal(np.random.get_state()[1], state[1])
            assert np.array_equal(spec_r, spec_d) and label_r == label_d
            assert params.dtype == torch.int32 and tuple(params.shape) == (2, 6)
            for t, (src, i, j, h, w, flip) in enumerate(params.tolist()):
                assert src == idx * 2 + t
                got = crop_resize_flip_normalize(store[src], i, j, h, w, flip)
                assert np.array_equal(images[:, t].numpy().view(np.uint32), got.view(np.uint32))
        assert "device_pipeline" not in pickle.loads(pickle.dumps(dev_ds)).__dict__

"""GPU parity of the visual data-pipeline kernel (csrc/datapipe.cu) through the C-ABI: bit-exact against the CPU
oracle (itself pinned to torchvision / Pillow) and against the golden digests generated from torchvision."""
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

pytestmark = pytest.mark.gpu


def _pipe(frames_u8, T):
    from gdl_b200.datapipe import DeviceFrameStore, VisualPipeline
    return VisualPipeline(DeviceFrameStore(torch.from_numpy(frames_u8)), T)


def test_golden_cases_bit_exact():
    from make_crop_golden import image_of
    g = np.load(os.path.join(ROOT, "tests", "golden", "crop_golden.npz"))
    for k, (H, W, i, j, h, w, flip) in enumerate(g["cases"].tolist()):
        pipe = _pipe(image_of(k, H, W)[None], 1)
        params = torch.tensor([[0, i, j, h, w, flip]], dtype=torch.int32, device="cuda")
        out = pipe(params)[0, :, 0].cpu().numpy()  # [3, 224, 224]
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g["digests"][k]), k


@pytest.mark.parametrize("H,W,T", [(360, 480, 3), (256, 340, 3), (224, 224, 1), (61, 45, 2)])
def test_random_crops_match_oracle(H, W, T):
    from oracle.crop_oracle import crop_resize_flip_normalize
    rs = np.random.RandomState(H + W)
    n_store, B = 7, 4
    store = rs.randint(0, 256, size=(n_store, H, W, 3), dtype=np.uint8)
    rows = []
    for f in range(B * T):
        h, w = int(rs.randint(1, H + 1)), int(rs.randint(1, W + 1))
        if f == 0:
            h, w = H, W  # Resize((224,224)) of the whole frame (test split)
        rows.append([int(rs.randint(n_store)), int(rs.randint(0, H - h + 1)), int(rs.randint(0, W - w + 1)), h, w,
                     int(rs.randint(2))])
    pipe = _pipe(store, T)
    out = pipe(torch.tensor(rows, dtype=torch.int32, device="cuda"))
    assert out.shape == (B, 3, T, 224, 224) and out.dtype == torch.float32
    got = out.cpu().numpy()
    for f, (src, i, j, h, w, flip) in enumerate(rows):
        ref = crop_resize_flip_normalize(store[src], i, j, h, w, flip)
        b, t = divmod(f, T)
        assert np.array_equal(got[b, :, t].view(np.uint32), ref.view(np.uint32)), (f, rows[f])
    # deterministic, and writing into a caller-provided tensor gives the same bits
    out2 = torch.empty_like(out)
    pipe(torch.tensor(rows, dtype=torch.int32, device="cuda"), out=out2)
    assert torch.equal(out, out2)


def test_reference_dataset_batch_reproduced():
    """A batch of the reference-order synthetic dataset (torchvision transform on the host) == the device
    pipeline fed with the host-drawn crop boxes of the same RNG stream."""
    import argparse
    from gdl_b200.datapipe import draw_frame_params
    from gdl_b200.synthetic import SyntheticCramed, synth_image
    args = argparse.Namespace(fps=3)
    ds = SyntheticCramed(args, 'train', 4)
    torch.manual_seed(3)
    np.random.seed(3)
    host = torch.stack([ds[i][1] for i in range(4)])  # [4, 3, 3, 224, 224]
    frames, rows = [], []
    torch.manual_seed(3)
    for i in range(4):
        for t in range(3):
            frames.append(np.asarray(synth_image("%s/%d" % (ds.key(i), t), ds.frame_size)))
            rows.append(draw_frame_params(len(frames) - 1, 360, 480, 'train'))
    pipe = _pipe(np.stack(frames), 3)
    dev = pipe(torch.tensor(rows, dtype=torch.int32, device="cuda"))
    assert torch.equal(dev.cpu(), host)


def test_bad_arguments_are_rejected():
    from gdl_b200._lib import GdlError
    pipe = _pipe(np.zeros((1, 2000, 100, 3), dtype=np.uint8), 1)  # 2000 / 224 > 7.5
    with pytest.raises(GdlError):
        pipe(torch.tensor([[0, 0, 0, 2000, 100, 0]], dtype=torch.int32, device="cuda"))
    pipe = _pipe(np.zeros((1, 50, 50, 3), dtype=np.uint8), 2)
    with pytest.raises(ValueError):
        pipe(torch.zeros(3, 6, dtype=torch.int32, device="cuda"))  # not a multiple of T


def test_training_step_with_device_pipeline_is_bit_identical():
    """DGLStep.prefetch(..., pipeline=) == prefetch of the host-transformed frames: same losses, same weights."""
    import argparse
    import gdl_b200
    from gdl_b200.datapipe import DeviceFrameStore, VisualPipeline
    from gdl_b200.step import DGLStep
    from oracle.crop_oracle import crop_resize_flip_normalize
    from oracle.synth import SHAPES, make_batch
    Fq, Tt, T, H, W = SHAPES["tiny"]
    B = 4
    rs = np.random.RandomState(5)
    store = rs.randint(0, 256, size=(B * T, 90, 120, 3), dtype=np.uint8)
    rows = []
    for f in range(B * T):
        h, w = int(rs.randint(20, 91)), int(rs.randint(20, 121))
        rows.append([f, int(rs.randint(0, 90 - h + 1)), int(rs.randint(0, 120 - w + 1)), h, w, int(rs.randint(2))])
    host = torch.zeros(B, 3, T, H, W)
    for f, (src, i, j, h, w, flip) in enumerate(rows):
        host[f // T, :, f % T] = torch.from_numpy(crop_resize_flip_normalize(store[src], i, j, h, w, flip, size=H))
    spec, _, label = make_batch(B, 6, "tiny", seed=2)
    results = []
    for use_pipe in (False, True):
        args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modality="full")
        gdl_b200.setup_seed(0)
        model = gdl_b200.AVClassifier_DGL(args)
        model.apply(gdl_b200.weight_init)
        model.cuda().train()
        step = DGLStep(model, B, (Fq, Tt), (T, H, W), alpha=4.0, lr=0.01, use_graph=False)
        if use_pipe:
            pipe = VisualPipeline(DeviceFrameStore(torch.from_numpy(store)), T, size=H)
            step.prefetch(spec.pin_memory(), torch.tensor(rows, dtype=torch.int32).pin_memory(), label.pin_memory(),
                          pipeline=pipe)
        else:
            step.prefetch(spec.pin_memory(), host.pin_memory(), label.pin_memory())
        step.step()
        torch.cuda.synchronize()
        results.append((step.read_stats(), step.arena.param.clone()))
    assert results[0][0] == results[1][0]
    assert torch.equal(results[0][1], results[1][1])

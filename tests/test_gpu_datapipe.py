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


# ------------------------------------------------------------------------------------------------------------
# audio half: gdl_log_stft vs the oracle restating librosa.stft (oracle/stft_oracle.py,
# pinned to scipy.signal.stft by tests/test_cpu_stft.py).  Floating point: |log-spectrogram difference| <= 5e-5
# (the reference computes the FFT in float64 and stores complex64; the kernel does the same on the device, so what
# is left is the order of the FFT butterflies and a float32 hypot / log ulp).
# ------------------------------------------------------------------------------------------------------------
def _audio(n_clips, lens, seed):
    import numpy as np
    rs = np.random.RandomState(seed)
    waves = []
    for i in range(n_clips):
        n = lens[i % len(lens)]
        t = np.arange(n) / 16000.0
        w = 0.6 * rs.randn(n) * np.exp(-t * rs.uniform(0.2, 3.0)) + 0.9 * np.sin(2 * np.pi * rs.uniform(80, 4000) * t)
        waves.append((w * rs.uniform(0.3, 1.6)).astype(np.float32))   # some clips exceed [-1, 1]: clipping matters
    return waves


@pytest.mark.parametrize("geom", [("CREMAD", 22050 * 3, 512, 353, (257, 188)), ("KS", 16000 * 5, 256, 128, (129, 626))])
@pytest.mark.parametrize("pad_mode", ["reflect", "constant"])
def test_log_stft_matches_oracle(geom, pad_mode):
    import numpy as np
    from gdl_b200.datapipe import AudioPipeline, DeviceWaveStore
    from oracle import stft_oracle as S
    name, L, n_fft, hop, shape = geom
    waves = _audio(6, [41234, 55125, 64000, 170000], seed=11)
    store = DeviceWaveStore(waves, "cuda")
    pipe = AudioPipeline(store, L, n_fft, hop, pad_mode)
    rs = np.random.RandomState(2)
    B = 9
    clips = rs.randint(0, 6, size=B)
    starts = np.zeros(B, dtype=np.int64) if name == "CREMAD" else rs.randint(0, 16000 * 5 + 1, size=B)
    params = torch.tensor(np.stack([clips, starts], 1), dtype=torch.int32, device="cuda")
    out = pipe(params)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B,) + shape
    worst = 0.0
    for b in range(B):
        ref = S.log_spectrogram(waves[clips[b]], int(starts[b]), L, n_fft, hop, pad_mode)
        worst = max(worst, float(np.abs(out[b].cpu().numpy() - ref).max()))
    assert worst <= 5e-5, worst


def test_log_stft_golden_and_determinism():
    """The committed scipy-generated vectors (tests/golden/stft_golden.npz), straight against the kernel."""
    import numpy as np
    from gdl_b200.datapipe import AudioPipeline, DeviceWaveStore
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stft_golden.npz"))
    for name in ("noise", "tones", "clipped"):
        w = g["wave_" + name]
        store = DeviceWaveStore([w], "cuda")
        for n_fft, hop in ((512, 353), (256, 128)):
            pipe = AudioPipeline(store, len(w), n_fft, hop)
            p = torch.zeros(1, 2, dtype=torch.int32, device="cuda")
            a, b = pipe(p), pipe(p)
            torch.cuda.synchronize()
            assert torch.equal(a, b)
            ref = g["spec_%s_%d_%d" % (name, n_fft, hop)]
            assert np.abs(a[0].cpu().numpy() - ref).max() <= 5e-5, (name, n_fft)


def test_step_fed_by_the_device_audio_pipeline():
    """A training step whose spectrograms are produced on the device from resident waveforms equals (bit for bit in
    its inputs' bf16 rounding aside, to 1e-3 in the losses) a step fed with host-computed spectrograms."""
    import argparse
    import numpy as np
    import gdl_b200
    from gdl_b200.datapipe import AudioPipeline, DeviceWaveStore
    from gdl_b200.step import DGLStep
    from oracle import stft_oracle as S
    B, L, n_fft, hop = 4, 4000, 128, 67          # tiny geometry: spectrogram 65 x 60 = the "tiny" test shape
    waves = _audio(5, [3000, 5100], seed=4)
    clips, starts = np.array([0, 3, 2, 4]), np.array([0, 17, 2999, 801])
    spec_host = torch.tensor(np.stack([S.log_spectrogram(waves[c], int(s), L, n_fft, hop) for c, s in zip(clips, starts)]))
    assert tuple(spec_host.shape) == (B, 65, 60)
    g = torch.Generator().manual_seed(3)
    image = torch.randn(B, 3, 2, 64, 64, generator=g)
    label = torch.randint(0, 6, (B,), generator=g)
    res = []
    for device_audio in (False, True):
        args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modality="full")
        gdl_b200.setup_seed(0)
        model = gdl_b200.AVClassifier_DGL(args)
        model.apply(gdl_b200.weight_init)
        model.cuda().train()
        step = DGLStep(model, B, (65, 60), (2, 64, 64), lr=0.01, use_graph=False)
        if device_audio:
            pipe = AudioPipeline(DeviceWaveStore(waves, "cuda"), L, n_fft, hop)
            params = torch.tensor(np.stack([clips, starts], 1), dtype=torch.int32).pin_memory()
            step.prefetch(params, image.pin_memory(), label.pin_memory(), audio_pipeline=pipe)
            step.step()
        else:
            step.step(spec_host.cuda(), image.cuda(), label.cuda())
        res.append(step.read_stats())
    for a, b in zip(res[0][:4], res[1][:4]):
        assert abs(a - b) <= 1e-3 * abs(a), (res[0], res[1])


def test_device_dataset_delivers_audio_params_and_matches_the_host_dataset():
    """SyntheticCramedDevice with the audio pipeline attached: items carry {clip, start}; the spectrograms computed
    on the device equal the ones SyntheticCramed (the reference's sample contract, CramedDataset.py:57-110) computes
    on the host, and train_epoch / valid consume them through DGLStep.prefetch(audio_pipeline=...)."""
    import argparse
    from gdl_b200.synthetic import SyntheticCramed, SyntheticCramedDevice
    args = argparse.Namespace(dataset="CREMAD", fps=2, use_video_frames=3)
    host = SyntheticCramed(args, "test", 4)
    dev = SyntheticCramedDevice(args, "test", 4)
    dev.attach_pipeline(torch.device("cuda"))
    assert dev.device_audio_pipeline is not None
    params = torch.stack([dev[i][0] for i in range(4)]).cuda()
    assert params.dtype == torch.int32 and tuple(params.shape) == (4, 2)
    spec = dev.device_audio_pipeline(params).cpu()
    assert tuple(spec.shape) == (4, 257, 188)
    for i in range(4):
        ref = torch.as_tensor(host[i][0])
        assert (spec[i] - ref).abs().max().item() <= 2e-4   # the host stand-in multiplies the window in float32

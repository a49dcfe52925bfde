"""Synthetic datasets with the reference's sample contract AND random-draw order
(dataset/CramedDataset.py:57-110, KSDataset.py:136-201, VGGSoundDataset.py:109-159):
`(spectrogram f32[F,Tt], images f32[3,T,224,224], label int)`.

There are no datasets in the build environment (no network, no librosa), so only the FILE DECODING is
replaced — a seeded waveform instead of `librosa.load`, a seeded PIL image instead of `Image.open` — while
everything that consumes random numbers is kept exactly as the reference does it, in the same order:

  CREMA-D   np.random.choice(n_frames, size=fps, replace=False)  (drawn, result unused: frames 0..fps-1 are read),
            then per frame torchvision RandomResizedCrop(224) (torch global RNG: up to 10 x (area, log-ratio),
            then i, j) and RandomHorizontalFlip (one torch.rand(1)).
  KS / VGG  random.randint(0, rate*5) for the 5-second audio crop (python RNG) first, then the same as above.

so that, given the same seeds, the crop offsets / flips / audio offsets are bit-identical to the reference's
(tests/test_cpu_sampling.py compares against the reference classes with their file I/O mocked, and against
golden digests where the reference tree is not available).  Sampler and worker seeding are torch's own
DataLoader, used unchanged by main_dgl.py.

`SyntheticAV` is the light-weight variant (pre-normalised random tensors, no PIL work) used for throughput runs.
"""
import hashlib
import random

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset
from torchvision import transforms

SHAPES = {"CREMAD": (257, 188), "KineticSound": (129, 626), "VGGSound": (129, 626), "AVE": (257, 1004),
          "kinect400": (129, 626)}
N_LABELS = {"CREMAD": 6, "KineticSound": 31, "VGGSound": 309, "AVE": 28, "kinect400": 400}


class SyntheticAV(Dataset):
    def __init__(self, args, mode='train', length=None):
        self.F, self.Tt = SHAPES[args.dataset]
        self.T = args.fps if args.dataset == 'CREMAD' else args.use_video_frames  # CramedDataset.py:44
        self.n = N_LABELS[args.dataset]
        self.len = length or (6698 if mode == 'train' else 744)  # CREMA-D split sizes
        self.seed = 0 if mode == 'train' else 1

    def __len__(self):
        return self.len

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        spec = torch.randn(self.F, self.Tt, generator=g) * 2.0 - 3.0
        image = torch.randn(3, self.T, 224, 224, generator=g)
        label = int(torch.randint(0, self.n, (1,), generator=g))
        return spec, image, label


# ------------------------------------------------------------------------------------------------
# decoding stand-ins (consume NO global random numbers)
# ------------------------------------------------------------------------------------------------
def _seed_of(key):
    return int.from_bytes(hashlib.sha256(str(key).encode()).digest()[:4], "little")


def synth_wave(key, seconds, rate):
    """Stand-in for librosa.load(path, sr=rate): a seeded waveform in [-1.2, 1.2] (so clipping matters)."""
    rs = np.random.RandomState(_seed_of(("wav", key)))
    return (rs.standard_normal(int(seconds * rate)) * 0.4).astype(np.float32), rate


def synth_image(key, size=(480, 360)):
    """Stand-in for Image.open(path).convert('RGB'): a seeded RGB image of the dataset's frame size (W, H)."""
    rs = np.random.RandomState(_seed_of(("img", key)))
    return Image.fromarray(rs.randint(0, 256, size=(size[1], size[0], 3), dtype=np.uint8), "RGB")


def stft(x, n_fft, hop_length):
    """|librosa.stft| geometry (center=True, reflect padding, periodic hann window): [1 + n_fft/2, 1 + len/hop]."""
    x = np.pad(np.asarray(x, dtype=np.float32), n_fft // 2, mode="reflect")
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)
    n_frames = 1 + (len(x) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    return np.fft.rfft(x[idx] * win[None, :], axis=1).T.astype(np.complex64)


def _transform(mode):
    # identical to the reference's per-item transform (CramedDataset.py:76-89, KSDataset.py:160-173)
    if mode == 'train':
        return transforms.Compose([
            transforms.RandomResizedCrop(224),
            transforms.RandomHorizontalFlip(),
            transforms.ToTensor(),
            transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    return transforms.Compose([
        transforms.Resize(size=(224, 224)),
        transforms.ToTensor(),
        transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])


class SyntheticCramed(Dataset):
    """CREMA-D sample contract and draw order (dataset/CramedDataset.py:57-110)."""

    def __init__(self, args, mode='train', length=64, frames_in_dir=3, frame_size=(480, 360)):
        self.args, self.mode, self.len = args, mode, length
        self.frames_in_dir, self.frame_size = max(frames_in_dir, args.fps), frame_size

    def __len__(self):
        return self.len

    def key(self, idx):
        return "%s/%d" % (self.mode, idx)

    def __getitem__(self, idx):
        samples, rate = synth_wave(self.key(idx), 2.5, 22050)
        resamples = np.tile(samples, 3)[:22050 * 3]          # 3 s at 22 050 Hz, clipped (CramedDataset.py:60-63)
        resamples[resamples > 1.] = 1.
        resamples[resamples < -1.] = -1.
        spectrogram = np.log(np.abs(stft(resamples, n_fft=512, hop_length=353)) + 1e-7)
        transform = _transform(self.mode)
        fps = self.args.fps
        select_index = np.random.choice(self.frames_in_dir, size=fps, replace=False)  # drawn, unused (:92-93)
        select_index.sort()
        images = torch.zeros((fps, 3, 224, 224))
        for i in range(fps):
            images[i] = transform(synth_image("%s/%d" % (self.key(idx), i), self.frame_size))
        images = torch.permute(images, (1, 0, 2, 3))
        label = _seed_of(("label", self.key(idx))) % N_LABELS["CREMAD"]
        return spectrogram, images, label


class SyntheticCramedDevice(SyntheticCramed):
    """SyntheticCramed for the device-side visual pipeline (gdl_b200/datapipe.py): the item carries the crop
    boxes / flips instead of the pixels — drawn from the same RNG stream, in the same order, as
    SyntheticCramed's torchvision transform — and the decoded frames live in one uint8 store (frame t of item i
    at index i * fps + t) that the caller uploads once (`attach_pipeline`)."""

    device_pipeline = None        # set by attach_pipeline; read by gdl_b200.train.train_epoch / valid
    device_audio_pipeline = None  # likewise: datapipe.AudioPipeline (spectrograms computed on the device)

    def __getstate__(self):  # DataLoader workers never see the CUDA-side pipelines
        d = dict(self.__dict__)
        d.pop("device_pipeline", None)
        d.pop("device_audio_pipeline", None)
        return d

    def frame_store(self):
        """uint8 [len * fps, H, W, 3]: what Image.open(...).convert('RGB') yields for every frame the dataset reads."""
        fps = self.args.fps
        out = np.empty((self.len * fps, self.frame_size[1], self.frame_size[0], 3), dtype=np.uint8)
        for idx in range(self.len):
            for i in range(fps):
                out[idx * fps + i] = np.asarray(synth_image("%s/%d" % (self.key(idx), i), self.frame_size))
        return torch.from_numpy(out)

    def wave_store(self):
        """What librosa.load(path, sr=22050) yields for every clip (clip idx at row idx)."""
        return [synth_wave(self.key(idx), 2.5, 22050)[0] for idx in range(self.len)]

    def attach_pipeline(self, device, audio=True):
        from .datapipe import AudioPipeline, DeviceFrameStore, DeviceWaveStore, VisualPipeline
        self.device_pipeline = VisualPipeline(DeviceFrameStore(self.frame_store(), device), self.args.fps)
        self.device_audio = bool(audio)
        if audio:  # CramedDataset.py:60-66: 3 s at 22 050 Hz, n_fft 512, hop 353
            self.device_audio_pipeline = AudioPipeline(DeviceWaveStore(self.wave_store(), device), 22050 * 3, 512, 353)
        return self.device_pipeline

    def __getitem__(self, idx):
        from .datapipe import draw_frame_params
        if getattr(self, "device_audio", False):
            spectrogram = torch.tensor([idx, 0], dtype=torch.int32)  # {clip, start}: np.tile(samples, 3)[:L] starts at 0
        else:
            samples, rate = synth_wave(self.key(idx), 2.5, 22050)
            resamples = np.tile(samples, 3)[:22050 * 3]
            resamples[resamples > 1.] = 1.
            resamples[resamples < -1.] = -1.
            spectrogram = np.log(np.abs(stft(resamples, n_fft=512, hop_length=353)) + 1e-7)
        fps = self.args.fps
        select_index = np.random.choice(self.frames_in_dir, size=fps, replace=False)  # drawn, unused (:92-93)
        select_index.sort()
        W, H = self.frame_size
        params = torch.tensor([draw_frame_params(idx * fps + i, H, W, self.mode) for i in range(fps)],
                              dtype=torch.int32)
        label = _seed_of(("label", self.key(idx))) % N_LABELS["CREMAD"]
        return spectrogram, params, label


class SyntheticKS(Dataset):
    """Kinetics-Sounds / VGGSound sample contract and draw order (dataset/KSDataset.py:136-201)."""

    def __init__(self, args, mode='train', length=64, frames_in_dir=10, frame_size=(340, 256), n_labels=31):
        self.args, self.mode, self.len = args, mode, length
        self.frames_in_dir, self.frame_size, self.n_labels = max(frames_in_dir, args.use_video_frames), frame_size, n_labels

    def __len__(self):
        return self.len

    def key(self, idx):
        return "%s/%d" % (self.mode, idx)

    def __getitem__(self, idx):
        sample, rate = synth_wave(self.key(idx), 4.0, 16000)
        while len(sample) / rate < 10.:
            sample = np.tile(sample, 2)
        start_point = random.randint(a=0, b=rate * 5)         # python RNG, drawn FIRST (KSDataset.py:143)
        new_sample = sample[start_point:start_point + rate * 5]
        new_sample[new_sample > 1.] = 1.
        new_sample[new_sample < -1.] = -1.
        spectrogram = np.log(np.abs(stft(new_sample, n_fft=256, hop_length=128)) + 1e-7)
        transform = _transform(self.mode)
        T = self.args.use_video_frames
        select_index = np.random.choice(self.frames_in_dir, size=T, replace=False)    # drawn, unused (:178-179)
        select_index.sort()
        images = torch.zeros((T, 3, 224, 224))
        for i in range(T):
            images[i] = transform(synth_image("%s/%d" % (self.key(idx), i), self.frame_size))
        images = torch.permute(images, (1, 0, 2, 3))
        label = _seed_of(("label", self.key(idx))) % self.n_labels
        return spectrogram, images, label

"""Device-side visual data pipeline: the reference's per-item transform

    RandomResizedCrop(224) | Resize((224, 224)) -> RandomHorizontalFlip -> ToTensor -> Normalize
    (dataset/CramedDataset.py:76-89,96-101; dataset/KSDataset.py:160-173,183-190)

with the random draws kept on the HOST in the reference's order (torchvision's own `get_params`, then one
`torch.rand(1)` per frame) and everything that touches pixels done by libgdl_b200.so's `gdl_crop_resize_normalize`
on decoded uint8 frames that stay resident in HBM.  The kernel is bit-exact with torchvision's PIL backend, so a
batch produced here equals the reference DataLoader's `image` tensor bit for bit (tests/test_gpu_datapipe.py),
while a training step uploads 24 bytes per frame instead of 602 KB of fp32 pixels.

No CPU fallback: the transform itself is only ever computed by the CUDA kernel (the CPU restatement lives in
oracle/crop_oracle.py and is test infrastructure).
"""
import ctypes as C

import torch
from torchvision import transforms

from . import _lib, ops

MEAN = (0.485, 0.456, 0.406)  # CramedDataset.py:81
STD = (0.229, 0.224, 0.225)
PARAM_INTS = 6  # store index, top, left, height, width, flip
_SCALE, _RATIO = (0.08, 1.0), (3.0 / 4.0, 4.0 / 3.0)  # RandomResizedCrop defaults used by the reference


def draw_frame_params(store_index, height, width, mode):
    """The random draws of ONE frame of the reference transform, in its order, from torch's global RNG:
    RandomResizedCrop.get_params (area, log-ratio, top, left; up to 10 attempts, then the centre-crop fallback),
    then RandomHorizontalFlip's `torch.rand(1) < 0.5`.  mode != 'train' is Resize((224, 224)): no draws."""
    if mode != 'train':
        return [store_index, 0, 0, height, width, 0]
    probe = torch.empty(3, height, width, device="meta")  # get_params only reads the size
    i, j, h, w = transforms.RandomResizedCrop.get_params(probe, list(_SCALE), list(_RATIO))
    flip = int(bool(torch.rand(1) < 0.5))
    return [store_index, i, j, h, w, flip]


class DeviceFrameStore:
    """Decoded RGB frames, uint8 [n, H, W, 3], resident on the GPU (CREMA-D at 3 frames per clip is ~10 GB)."""

    def __init__(self, frames_u8, device=None):
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
            raise ValueError("DeviceFrameStore needs uint8 frames [n, H, W, 3]")
        self.frames = frames_u8.to(device if device is not None else "cuda").contiguous()
        self.n, self.H, self.W = self.frames.shape[:3]


class VisualPipeline:
    """params (int32 [B*T, 6], device) -> fp32 [B, 3, T, S, S] on the current stream."""

    def __init__(self, store, T, size=224, max_frames=None):
        ops.init()
        self.store, self.T, self.S = store, T, size
        self._mean = (C.c_float * 3)(*MEAN)
        self._std = (C.c_float * 3)(*STD)
        self._table = None
        if max_frames:
            self._reserve(max_frames)

    def _reserve(self, frames):
        need = int(_lib.load().gdl_crop_table_ints(frames, self.S))
        if self._table is None or self._table.numel() < need:
            self._table = torch.empty(need, device=self.store.frames.device, dtype=torch.int32)

    def __call__(self, params, out=None):
        if params.dtype != torch.int32 or params.dim() != 2 or params.shape[1] != PARAM_INTS or not params.is_cuda:
            raise ValueError("params must be a CUDA int32 tensor [frames, 6]")
        frames = params.shape[0]
        if frames % self.T:
            raise ValueError("frames must be a multiple of T")
        self._reserve(frames)
        if out is None:
            out = torch.empty(frames // self.T, 3, self.T, self.S, self.S, device=params.device)
        ops.crop_resize_normalize(self.store.frames, self.store.n, self.store.H, self.store.W, params.contiguous(),
                                  frames, self.T, self.S, self._mean, self._std, out, self._table)
        return out


class DeviceWaveStore:
    """Decoded mono waveforms resident on the GPU: fp32 [n_clips, max_len] (rows zero padded) + their lengths."""

    def __init__(self, waves, device=None):
        dev = device if device is not None else "cuda"
        n = len(waves)
        self.lengths = torch.tensor([len(w) for w in waves], dtype=torch.int32)
        self.stride = int(self.lengths.max())
        buf = torch.zeros(n, self.stride)
        for i, w in enumerate(waves):
            buf[i, :len(w)] = torch.as_tensor(w, dtype=torch.float32)
        self.waves = buf.to(dev).contiguous()
        self.lengths = self.lengths.to(dev)
        self.n = n


class AudioPipeline:
    """params (int32 [B, 2] = {clip index, start sample}, device) -> fp32 [B, 1 + n_fft/2, 1 + L/hop] on the current
    stream: the reference's spectrogram np.log(np.abs(librosa.stft(clip(samples), n_fft, hop)) + 1e-7)
    (dataset/CramedDataset.py:60-66: L = 22050*3, n_fft 512, hop 353, start 0; KSDataset.py:138-150 /
    VGGSoundDataset.py:112-122: L = 16000*5, n_fft 256, hop 128, start = the host-drawn random.randint) computed by
    libgdl_b200.so's gdl_log_stft from waveforms that stay in HBM.  pad_mode "reflect" is librosa < 0.10 (the
    reference's era), "constant" librosa >= 0.10."""

    def __init__(self, store, L, n_fft, hop, pad_mode="reflect"):
        ops.init()
        if pad_mode not in ("reflect", "constant"):
            raise ValueError("pad_mode must be 'reflect' or 'constant'")
        self.store, self.L, self.n_fft, self.hop = store, int(L), int(n_fft), int(hop)
        self.pad = 0 if pad_mode == "reflect" else 1
        self.F, self.frames = 1 + self.n_fft // 2, 1 + self.L // self.hop

    def __call__(self, params, out=None):
        if params.dtype != torch.int32 or params.dim() != 2 or params.shape[1] != 2 or not params.is_cuda:
            raise ValueError("params must be a CUDA int32 tensor [B, 2]")
        B = params.shape[0]
        if out is None:
            out = torch.empty(B, self.F, self.frames, device=params.device)
        ops.log_stft(self.store.waves, self.store.stride, self.store.lengths, params.contiguous(), B, self.L, self.n_fft,
                     self.hop, self.pad, out)
        return out

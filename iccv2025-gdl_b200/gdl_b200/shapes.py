"""Input geometries of the reference's datasets and seeded synthetic batches of its data contract
(dataset/CramedDataset.py:57-110, KSDataset.py:136-201): (spectrogram f32[B,F,Tt], images f32[B,3,T,H,W],
label i64[B]).  Used by bench.py's GPU arm, main_dgl.py's synthetic loaders and (re-exported by oracle/synth.py)
the tests — so the product bench has no import from oracle/."""
import torch

BATCH_SHAPES = {
    # name: (F, Tt, T, H, W)
    "CREMAD": (257, 188, 3, 224, 224),        # 22 050 Hz * 3 s, n_fft 512, hop 353 (CramedDataset.py:60-66)
    "KineticSound": (129, 626, 3, 224, 224),  # 16 kHz * 5 s, n_fft 256, hop 128 (KSDataset.py:139-148)
    "VGGSound": (129, 626, 3, 224, 224),
    "tiny": (65, 60, 2, 64, 64),
}
HEAD_WIDTH = {"CREMAD": 6, "KineticSound": 34, "VGGSound": 309}   # reference models/basic_model.py:15-26
LABEL_MAX = {"CREMAD": 6, "KineticSound": 31, "VGGSound": 309}    # KineticSound: 34-wide head, 31 classes in the data


def make_batch(B, n_classes, shape="CREMAD", seed=1, label_max=None):
    Fq, Tt, T, H, W = BATCH_SHAPES[shape] if isinstance(shape, str) else shape
    g = torch.Generator().manual_seed(seed)
    spec = torch.randn(B, Fq, Tt, generator=g) * 2.0 - 3.0   # log(|STFT|+1e-7)-like range
    image = torch.randn(B, 3, T, H, W, generator=g)          # normalised-image statistics
    label = torch.randint(0, label_max or n_classes, (B,), generator=g)
    return spec, image, label

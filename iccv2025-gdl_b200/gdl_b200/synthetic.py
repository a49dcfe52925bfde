"""Synthetic datasets with the reference's sample contract (dataset/CramedDataset.py:57-110,
KSDataset.py:136-201): `(spectrogram f32[F,Tt], images f32[3,T,H,W], label int)`.  There are no
datasets in the build environment (no network), so the CLI trains on these when
`--audio_path synthetic` is given; shapes follow SURVEY.md §8a."""
import torch
from torch.utils.data import Dataset

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

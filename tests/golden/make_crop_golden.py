"""Golden vectors for the visual data pipeline, generated from torchvision / Pillow THEMSELVES (the third-party
code the reference's transform calls: dataset/CramedDataset.py:76-89):

    python tests/golden/make_crop_golden.py        # writes tests/golden/crop_golden.npz

Each case: a seeded uint8 RGB image, a crop box (i, j, h, w), a flip flag -> F.resized_crop(PIL, ..., [224, 224])
-> hflip -> to_tensor -> normalize.  The fp32 result is stored as a sha256 digest plus 16 probe values, so the
fixture stays small; tests compare the CPU oracle (and, on a GPU box, the CUDA kernel) against them bit for bit."""
import hashlib
import os

import numpy as np
import torchvision.transforms.functional as F
from PIL import Image

CASES = [  # (H, W, i, j, h, w, flip)
    (360, 480, 0, 0, 360, 480, 0),      # Resize((224,224)) of a CREMA-D frame (test split)
    (360, 480, 40, 0, 282, 329, 1),     # typical RandomResizedCrop draws
    (360, 480, 123, 242, 135, 108, 1),  # up-scaling in x, down-scaling in y
    (256, 340, 11, 82, 224, 224, 0),    # Kinetics-Sounds frame, crop already 224x224 (both passes skipped)
    (256, 340, 3, 9, 224, 100, 1),      # vertical pass skipped
    (256, 340, 7, 5, 57, 224, 0),       # horizontal pass skipped
    (97, 131, 5, 7, 1, 1, 1),           # 1-pixel crop
    (700, 900, 0, 0, 700, 900, 0),      # scale > 4: 9 coefficients per sample
]


def image_of(k, H, W):
    return np.random.RandomState(1000 + k).randint(0, 256, size=(H, W, 3), dtype=np.uint8)


def reference(img, i, j, h, w, flip):
    r = F.resized_crop(Image.fromarray(img, "RGB"), i, j, h, w, [224, 224])
    if flip:
        r = F.hflip(r)
    return F.normalize(F.to_tensor(r), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]).numpy()


def main():
    out = {"cases": np.array(CASES, dtype=np.int32)}
    digests, probes = [], []
    for k, (H, W, i, j, h, w, flip) in enumerate(CASES):
        ref = np.ascontiguousarray(reference(image_of(k, H, W), i, j, h, w, flip))
        digests.append(hashlib.sha256(ref.tobytes()).hexdigest())
        probes.append(ref.reshape(-1)[:: ref.size // 16][:16].copy())
    out["digests"] = np.array(digests)
    out["probes"] = np.stack(probes).astype(np.float32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "crop_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()

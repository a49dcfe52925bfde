"""gdl_b200 — B200-native drop-in for the DGL training step of shicaiwei123/ICCV2025-GDL.

Host-side mirror of the reference interface (models/basic_model.py, models/backbone.py,
models/fusion_modules.py, utils/utils.py, main_dgl.train_epoch) over the C-ABI of
libgdl_b200.so (include/gdl_b200.h).  There is no CPU or library fallback: importing is cheap,
but any compute call raises GdlError when the extension is missing.
"""
from ._lib import GdlError, LIB_PATH  # noqa: F401
from .basic_model import AVClassifier, AVClassifier_DGL  # noqa: F401
from .backbone import resnet18  # noqa: F401
from .fusion_modules import (ConcatFusion_DGL, FiLM_DGL, GatedFusion_DGL,  # noqa: F401
                             SumFusion_DGL)
from .utils import setup_seed, weight_init  # noqa: F401

__all__ = ["AVClassifier_DGL", "AVClassifier", "resnet18", "ConcatFusion_DGL", "SumFusion_DGL",
           "FiLM_DGL", "GatedFusion_DGL", "setup_seed", "weight_init", "GdlError"]

"""Seeded synthetic batches of the reference's data contract (dataset/CramedDataset.py:57-110,
KSDataset.py:136-201): (spectrogram f32[B,F,Tt], images f32[B,3,T,H,W], label i64[B]).
Test infrastructure; the table and the generator live in the package (gdl_b200/shapes.py) so that the product
bench imports nothing from oracle/ — this module re-exports them for tests/golden/make_golden.py and the tests."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iccv2025-gdl_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from gdl_b200.shapes import BATCH_SHAPES as SHAPES, make_batch  # noqa: E402,F401

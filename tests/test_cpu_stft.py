"""The spectrogram oracle (oracle/stft_oracle.py, a restatement of librosa.stft's published algorithm — librosa is a
third-party dependency absent from the reference tree and from this image) is pinned here against (a) golden vectors
computed by scipy.signal.stft, an independent implementation (tests/golden/make_stft_golden.py), (b) scipy run live,
and (c) the numpy stand-in the synthetic datasets use on the host (gdl_b200.synthetic.stft).  CPU only."""
import os

import numpy as np
import pytest

from oracle import stft_oracle as S

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "stft_golden.npz"))
GEOMS = ((512, 353), (256, 128))  # CramedDataset.py:65, KSDataset.py:148


@pytest.mark.parametrize("name", ["noise", "tones", "clipped"])
@pytest.mark.parametrize("geom", GEOMS)
def test_oracle_matches_scipy_golden(name, geom):
    n_fft, hop = geom
    w = GOLD["wave_" + name]
    got = S.log_spectrogram(w, 0, len(w), n_fft, hop)
    ref = GOLD["spec_%s_%d_%d" % (name, n_fft, hop)]
    assert got.shape == ref.shape == (1 + n_fft // 2, 1 + len(w) // hop)
    assert got.dtype == np.float32
    assert np.abs(got - ref).max() <= 2e-6      # float32 log of magnitudes spanning 100 dB


def test_oracle_matches_scipy_live_and_reference_geometry():
    import scipy.signal
    rs = np.random.RandomState(3)
    x = np.clip(rs.randn(22050 * 3) * 0.3, -1, 1).astype(np.float32)
    spec = S.log_spectrogram(x, 0, len(x), 512, 353)
    assert spec.shape == (257, 188)              # the CREMA-D spectrogram shape (BASELINE.json)
    win = scipy.signal.get_window("hann", 512, fftbins=True)
    _, _, z = scipy.signal.stft(x.astype(np.float64), window=win, nperseg=512, noverlap=512 - 353, boundary="even",
                                padded=False)
    ref = np.log(np.abs((z * win.sum()).astype(np.complex64)) + np.float32(1e-7))[:, :188]
    assert np.abs(spec - ref).max() <= 2e-6
    assert S.log_spectrogram(rs.randn(16000 * 4).astype(np.float32), 777, 16000 * 5, 256, 128).shape == (129, 626)


def test_tiling_window_and_clipping_semantics():
    """np.tile(samples, k)[start:start+L] == wave[(start + i) mod len]; values clipped to [-1, 1] before the STFT."""
    rs = np.random.RandomState(5)
    w = (rs.randn(5000) * 1.2).astype(np.float32)
    L, start = 12000, 3100
    manual = np.tile(w, 4)[start:start + L].copy()
    manual[manual > 1.] = 1.
    manual[manual < -1.] = -1.
    assert np.array_equal(S.item_samples(w, start, L), manual)
    a = S.log_spectrogram(w, start, L, 256, 128)
    b = np.log(np.abs(S.stft(manual, 256, 128)) + np.float32(1e-7))
    assert np.array_equal(a, b)
    # zero padding (librosa >= 0.10) differs from reflect only in the frames that overlap the edges
    z = S.log_spectrogram(w, start, L, 256, 128, pad_mode="constant")
    assert np.array_equal(a[:, 1:-1], z[:, 1:-1]) and not np.array_equal(a[:, 0], z[:, 0])


def test_host_stand_in_of_the_synthetic_datasets_agrees():
    from gdl_b200.synthetic import stft as host_stft
    rs = np.random.RandomState(7)
    x = np.clip(rs.randn(20000) * 0.5, -1, 1).astype(np.float32)
    for n_fft, hop in GEOMS:
        a = np.log(np.abs(host_stft(x, n_fft, hop)) + 1e-7)
        b = S.log_spectrogram(x, 0, len(x), n_fft, hop)
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-4   # the stand-in multiplies the window in float32

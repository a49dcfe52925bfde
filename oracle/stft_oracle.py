"""CPU restatement of the reference's spectrogram computation — TEST INFRASTRUCTURE, NOT THE PRODUCT.

Reference (dataset/CramedDataset.py:60-66, KSDataset.py:138-150, VGGSoundDataset.py:112-122):

    resamples = np.tile(samples, 3)[:22050 * 3]; clip to [-1, 1]
    spectrogram = np.log(np.abs(librosa.stft(resamples, n_fft=512, hop_length=353)) + 1e-7)

`librosa` is a third-party dependency that is absent from /root/reference and from this image (no requirements
file; README.md:6-10 pins only Ubuntu 20.04 / CUDA 11.1 / PyTorch 1.11 / Python 3.8.6, i.e. the librosa 0.8-0.9 era).
Its published algorithm (librosa/core/spectrum.py `stft`) is restated here:

    fft_window = scipy.signal.get_window("hann", n_fft, fftbins=True)        (periodic Hann, float64)
    y = np.pad(y, n_fft // 2, mode=pad_mode)       center=True; pad_mode "reflect" (< 0.10) or "constant" (>= 0.10)
    frames k: y[k*hop : k*hop + n_fft],  k = 0 .. (len(y) - n_fft) // hop       => 1 + len(x) // hop frames
    stft_matrix[:, k] = rfft(fft_window * frame)   float64 product and FFT, stored into a complex64 matrix
    np.abs -> float32, np.log -> float32

The oracle is PINNED against scipy.signal.stft (tests/test_cpu_stft.py: same window, boundary="even" == reflect
padding, un-normalised by the window sum) and against tests/golden/stft_golden.npz written by
tests/golden/make_stft_golden.py; only tests/ may import it."""
import numpy as np


def hann_periodic(n_fft):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft, dtype=np.float64) / n_fft)


def stft(y, n_fft, hop_length, pad_mode="reflect"):
    """complex64 [1 + n_fft/2, 1 + len(y)//hop] like librosa.stft(y, n_fft, hop_length) with its defaults."""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    frames = yp[idx].astype(np.float64) * hann_periodic(n_fft)[None, :]
    return np.fft.rfft(frames, axis=1).T.astype(np.complex64)


def item_samples(wave, start, L):
    """np.tile(wave, k)[start:start + L], clipped to [-1, 1] (CramedDataset.py:61-63, KSDataset.py:139-146)."""
    wave = np.asarray(wave, dtype=np.float32)
    idx = (start + np.arange(L, dtype=np.int64)) % len(wave)
    return np.clip(wave[idx], -1.0, 1.0)


def log_spectrogram(wave, start, L, n_fft, hop_length, pad_mode="reflect"):
    """float32 [1 + n_fft/2, 1 + L//hop]: the reference's `spectrogram`."""
    s = stft(item_samples(wave, start, L), n_fft, hop_length, pad_mode)
    return np.log(np.abs(s) + np.float32(1e-7)).astype(np.float32)

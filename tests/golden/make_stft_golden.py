"""Writes tests/golden/stft_golden.npz: seeded waveforms (white noise, a decaying two-tone signal with a 100 dB
dynamic range, clipped samples) and their reference spectrograms computed by scipy.signal.stft — an implementation
independent of oracle/stft_oracle.py — for the two geometries the reference uses (n_fft 512 / hop 353 and n_fft 256 /
hop 128).  Run in the build container: python tests/golden/make_stft_golden.py"""
import os

import numpy as np
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))


def scipy_log_spec(x, n_fft, hop):
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)
    # boundary="even" pads n_fft/2 samples by reflection (== np.pad mode="reflect"); scaling="spectrum" divides by
    # win.sum(): undo it to get librosa's un-normalised STFT
    _, _, z = scipy.signal.stft(x.astype(np.float64), window=win, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft,
                                boundary="even", padded=False, return_onesided=True)
    z = (z * win.sum()).astype(np.complex64)
    return np.log(np.abs(z) + np.float32(1e-7)).astype(np.float32)


def main():
    rs = np.random.RandomState(1234)
    out = {}
    cases = {"noise": (rs.randn(9000) * 0.4).astype(np.float32),
             "tones": (np.sin(2 * np.pi * 440 * np.arange(12000) / 22050.0) * np.exp(-np.arange(12000) / 900.0)
                       + 1e-5 * np.sin(2 * np.pi * 5000 * np.arange(12000) / 22050.0)).astype(np.float32),
             "clipped": (rs.randn(7000) * 1.5).astype(np.float32)}
    for name, w in cases.items():
        out["wave_" + name] = w
        for n_fft, hop in ((512, 353), (256, 128)):
            x = np.clip(w, -1.0, 1.0)
            spec = scipy_log_spec(x, n_fft, hop)
            n_frames = 1 + len(x) // hop
            out["spec_%s_%d_%d" % (name, n_fft, hop)] = spec[:, :n_frames]
    np.savez_compressed(os.path.join(HERE, "stft_golden.npz"), **out)
    print("wrote stft_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

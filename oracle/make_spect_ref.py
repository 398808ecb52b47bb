"""CPU oracle for the make_spect log-mel front-end.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this; the product path never does.

numpy/scipy restatement of the reference's ``Spect.spect`` spmel branch
(make_spect.py:50-83, :92-94) with its helpers ``butter_highpass`` (:30-34) and
``pySTFT`` (:36-48).  ``librosa`` (pinned 0.9.1, requirements.txt:1) is not installed in
this image, so the one librosa routine on the path, ``librosa.filters.mel(sr, n_fft,
fmin, fmax, n_mels)`` with its 0.9.1 defaults ``htk=False, norm='slaney',
dtype=float32``, is restated here from its published algorithm (Slaney auditory-toolbox
mel scale: linear below 1 kHz at 200/3 Hz per mel, logarithmic above with step
ln(6.4)/27; triangular filters between n_mels+2 band edges, each scaled by
2/(f[i+2]-f[i])).  ``librosa.load(path, sr=16000)`` on the bundled 16 kHz int16 files
reduces to ``int16 / 32768`` as float32.

Pinning: the bundled ``wavs/<spk>/*.wav`` -> ``spmel/<spk>/*.npy`` pairs are bit-level
goldens of this path.  ``oracle/gen_golden.py`` copies a subset (with the dither-stream
offset of each file) into ``tests/golden/frontend_*.npz``; in this container the full
71-pair check also runs (tests/test_oracle_frontend.py).
"""
from __future__ import annotations

import numpy as np
from scipy import signal
from scipy.signal import get_window

FS = 16000            # make_spect.py:22
CUTOFF = 30           # :21
ORDER = 5             # :23
N_FFT = 1024          # :24,:26
HOP = 256             # :25
N_MELS = 80           # :51
FMIN, FMAX = 90, 7600  # :51
MIN_LEVEL = np.exp(-100 / 20 * np.log(10))   # :52  (= 1e-5)


def butter_highpass():
    """make_spect.py:30-34."""
    nyq = 0.5 * FS
    return signal.butter(ORDER, CUTOFF / nyq, btype="high", analog=False)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=FS, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """librosa.filters.mel (0.9.1 defaults) -> (n_mels, 1 + n_fft//2) float32."""
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def py_stft(x):
    """make_spect.py:36-48 -> (513, F) float64 magnitudes, F = 1 + len(x)//256."""
    x = np.pad(x, N_FFT // 2, mode="reflect")
    noverlap = N_FFT - HOP
    n_frames = (x.shape[-1] - noverlap) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = x[idx]
    win = get_window("hann", N_FFT, fftbins=True)
    return np.abs(np.fft.rfft(win * frames, n=N_FFT).T)


def logmel_from_wav(x_f32: np.ndarray, dither_u01: np.ndarray, mel_basis=None, ba=None) -> np.ndarray:
    """make_spect.py:72-83,:94 for one utterance.

    ``x_f32``: waveform as ``librosa.load`` returns it (float32 in [-1, 1));
    ``dither_u01``: the ``prng.rand(N)`` uniform [0,1) draw the reference consumes for this
    utterance (:76).  Returns ``S`` (F, 80) float32 in [0, 1]."""
    if mel_basis is None:
        mel_basis = mel_filterbank().T                      # :51
    b, a = ba if ba is not None else butter_highpass()      # :53
    y = signal.filtfilt(b, a, x_f32)                        # :74
    wav = y * 0.96 + (dither_u01 - 0.5) * 1e-06             # :76
    D = py_stft(wav)                                        # :78
    D_mel = np.dot(D.T, mel_basis)                          # :81
    D_db = 20 * np.log10(np.maximum(MIN_LEVEL, D_mel)) - 16  # :82
    S = np.clip((D_db + 100) / 100, 0, 1)                   # :83
    return S.astype(np.float32)                             # :94


def logstft_from_wav(x_f32: np.ndarray, dither_u01: np.ndarray, ba=None) -> np.ndarray:
    """make_spect.py:72-78,:84-86,:94 (model_type 'stft') for one utterance: the log / clip of the 513 STFT magnitudes themselves.
    Returns ``S`` (513, F) float32 in [0, 1] -- the reference does not transpose D in this branch.  Shares every statement up to
    ``D`` with ``logmel_from_wav``, which is pinned to the reference's 71 bundled goldens."""
    b, a = ba if ba is not None else butter_highpass()      # :53
    y = signal.filtfilt(b, a, x_f32)                        # :74
    wav = y * 0.96 + (dither_u01 - 0.5) * 1e-06             # :76
    D = py_stft(wav)                                        # :78
    D_db = 20 * np.log10(np.maximum(MIN_LEVEL, D)) - 16     # :85
    S = np.clip((D_db + 100) / 100, 0, 1)                   # :86
    return S.astype(np.float32)                             # :94


def speaker_dither_streams(speaker_dir: str, lengths):
    """The per-speaker MT19937 stream of make_spect.py:68,:76: one ``RandomState(int(
    spk[1:]))`` consumed over the speaker's files in sorted order."""
    prng = np.random.RandomState(int(speaker_dir[1:]))
    return [prng.rand(n) for n in lengths]


def synthetic_waveforms(n_utt: int, n_samples: int, seed: int = 1234):
    """SURVEY §8(d) synthetic conversion inputs: noise shaped by a slow random envelope,
    clipped to [-1, 1), float32; and the matching uniform dither draws (float64 -> the
    kernels take them as float32... kept float64 here, callers cast)."""
    rs = np.random.RandomState(seed)
    t = np.arange(n_samples) / FS
    wav = np.empty((n_utt, n_samples), np.float32)
    for i in range(n_utt):
        env = 0.55 + 0.45 * np.sin(2 * np.pi * (0.3 + rs.rand()) * t + 2 * np.pi * rs.rand())
        wav[i] = np.clip(0.1 * rs.randn(n_samples) * env, -1.0, 1.0 - 2 ** -15).astype(np.float32)
    dither = rs.rand(n_utt, n_samples)
    return wav, dither

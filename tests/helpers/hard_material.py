"""PCM that pushes the HCA encoder's bit allocation to its corners: full-scale white noise (the noise-level search runs
to the top and, at low rates, gives bands up), full-scale square waves on and between transform bins (coefficients near
and on the quantiser's clamp), an impulse train and a full-scale step; every case as a mono and as a stereo WAV."""
import numpy as np

from pycricodecs_b200 import synth


def wavs(seed: int, n: int = 7000):
    rng = np.random.default_rng(seed)
    idx = np.arange(n)
    mats = [rng.integers(-32768, 32768, n), rng.choice([-32768, 32767], n)]
    for period in (2, 3, 8, 11, 51, 256 / 12.5, 256 / 40.0):
        mats.append(np.where(np.floor(idx * 2.0 / period).astype(np.int64) % 2 == 0, 32767, -32768))
    mats.append(np.where(idx % 97 == 0, 32767, 0))
    mats.append(np.where(idx > 3000, 32767, -32768))
    cases = []
    for k, m in enumerate(mats):
        mono = np.asarray(m, np.int64).clip(-32768, 32767).astype(np.int16)
        cases.append(synth.wav_header(1, n) + mono.tobytes())
        other = np.roll(mono, 37 * (k + 1))[::-1] if k % 2 else mono
        cases.append(synth.wav_header(2, n) + np.stack([mono, np.ascontiguousarray(other)], axis=1).tobytes())
    return cases

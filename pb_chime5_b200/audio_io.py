"""WAV reading / writing for the session driver.

Restates the two calls of the reference's I/O package that ``Enhancer.enhance_example`` /
``enhance_session`` make (``pb_chime5/core.py:389,454-462``):

* ``load_audio(path, start=, stop=)``  -- ``pb_chime5/io/audioread.py:34-225``: returns
  ``(channels, samples)`` (``(samples,)`` for mono) float64 in [-1, 1) (PCM / 2^(bits-1)),
  sample-indexed ``start`` / ``stop`` / ``frames``;
* ``dump_audio(x, path)``              -- ``pb_chime5/io/audiowrite.py:16-209``: 16-bit PCM,
  by default peak-normalised with ``(2^15 - 1) / 2^15 / max|x|``.

The reference goes through ``soundfile`` (libsndfile), which is not a dependency here: RIFF/WAVE
PCM (16 / 24 / 32 bit) and IEEE float (32 / 64 bit) files are parsed directly, so that a whole
utterance is one ``readinto`` of the slice that is needed.  Everything else (NIST SPHERE, FLAC,
...) raises ``RuntimeError`` like the reference does for files libsndfile cannot open.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_FMT_PCM, _FMT_FLOAT, _FMT_EXT = 1, 3, 0xFFFE


class WavInfo:
    __slots__ = ('channels', 'sample_rate', 'bits', 'is_float', 'data_offset', 'num_frames', 'frame_bytes')

    def __repr__(self):
        return (f'WavInfo(channels={self.channels}, sample_rate={self.sample_rate}, bits={self.bits}, '
                f'is_float={self.is_float}, num_frames={self.num_frames})')


def wav_info(path) -> WavInfo:
    """Parse the RIFF header; the data chunk is located, not read."""
    path = Path(path)
    with open(path, 'rb') as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b'RIFF' or head[8:12] != b'WAVE':
            raise RuntimeError(f'{path}: not a RIFF/WAVE file (header {head[:12]!r})')
        info = WavInfo()
        have_fmt = False
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise RuntimeError(f'{path}: no data chunk')
            cid, size = hdr[:4], struct.unpack('<I', hdr[4:])[0]
            if cid == b'fmt ':
                fmt = f.read(size + (size & 1))
                tag, ch, sr, _, block, bits = struct.unpack('<HHIIHH', fmt[:16])
                if tag == _FMT_EXT and size >= 26:
                    tag = struct.unpack('<H', fmt[24:26])[0]
                if tag not in (_FMT_PCM, _FMT_FLOAT):
                    raise RuntimeError(f'{path}: unsupported WAVE format tag {tag}')
                info.channels, info.sample_rate, info.bits = ch, sr, bits
                info.is_float = tag == _FMT_FLOAT
                info.frame_bytes = block
                have_fmt = True
            elif cid == b'data':
                if not have_fmt:
                    raise RuntimeError(f'{path}: data chunk before fmt chunk')
                info.data_offset = f.tell()
                f.seek(0, 2)
                avail = f.tell() - info.data_offset
                size = min(size, avail) if size not in (0, 0xFFFFFFFF) else avail   # streamed files
                info.num_frames = size // info.frame_bytes
                return info
            else:
                f.seek(size + (size & 1), 1)


def load_audio(path, *, frames=-1, start=0, stop=None, dtype=np.float64, expected_sample_rate=None,
               return_sample_rate=False, out=None):
    """audioread.py:34-225 (unit='samples').  ``out``: optional preallocated (channels, n) array
    (e.g. a view of a pinned buffer) that receives the samples."""
    info = wav_info(path)
    if expected_sample_rate is not None and expected_sample_rate != info.sample_rate:
        raise ValueError(f'Requested sampling rate is {expected_sample_rate} but the audiofile has {info.sample_rate}')
    n = info.num_frames
    start = int(start or 0)
    if start < 0:
        start += n
    stop = n if stop is None else (int(stop) + n if stop < 0 else int(stop))
    start, stop = max(0, min(start, n)), max(0, min(stop, n))
    count = max(0, stop - start)
    if frames is not None and frames >= 0:
        count = min(count, int(frames))
    width = info.bits // 8
    raw = np.empty(count * info.channels * width, dtype=np.uint8)
    with open(path, 'rb') as f:
        f.seek(info.data_offset + start * info.frame_bytes)
        got = f.readinto(memoryview(raw))
    if got != raw.size:
        raise RuntimeError(f'{path}: short read ({got} of {raw.size} bytes)')
    if info.is_float:
        x = raw.view('<f4' if info.bits == 32 else '<f8').astype(np.float64)
    elif info.bits == 16:
        x = raw.view('<i2').astype(np.float64) / 32768.0
    elif info.bits == 32:
        x = raw.view('<i4').astype(np.float64) / 2147483648.0
    elif info.bits == 24:
        b = raw.reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = v.astype(np.float64) / 8388608.0
    elif info.bits == 8:
        x = (raw.astype(np.float64) - 128.0) / 128.0
    else:
        raise RuntimeError(f'{path}: unsupported sample width {info.bits}')
    x = x.reshape(count, info.channels).T                       # soundfile gives (samples, channels); transposed
    if dtype is not None and x.dtype != np.dtype(dtype):
        if np.dtype(dtype).kind == 'i' and not info.is_float:
            x = np.rint(x * float(1 << (8 * np.dtype(dtype).itemsize - 1))).astype(dtype)
        else:
            x = x.astype(dtype)
    if info.channels == 1:
        x = x[0]
    if out is not None:
        out[...] = x
        x = out
    return (x, info.sample_rate) if return_sample_rate else x


def dump_audio(obj, path, *, sample_rate=16000, dtype=np.int16, normalize=True):
    """audiowrite.py:16-209 for the case the enhancer uses (new file, 16-bit PCM).  obj: (samples,)
    or (channels, samples).  normalize: scale so that the peak is (2^15 - 1) / 2^15
    (audiowrite.py:150-153); without it float input is expected in [-1, 1).  The float -> PCM
    conversion truncates ``x * 2^15`` toward zero and clips, which is what the reference's
    doctest pins (audiowrite.py:36-38: [1, 2, -4, 4] reads back as k / 2^15 with k = 8191,
    16383, -32767, 32767)."""
    obj = np.asarray(obj)
    if dtype not in (np.int16, np.dtype('int16')):
        raise TypeError(dtype)
    if normalize:
        if obj.dtype.kind not in 'fi':
            raise TypeError(f'Only float and int is currently supported with normalize. Got dtype {obj.dtype}')
        correction = (2 ** 15 - 1) / (2 ** 15)
        obj = obj * (correction / np.amax(np.abs(obj)))
    if obj.dtype.kind == 'f':
        pcm = np.clip(np.trunc(obj * 32768.0), -32768, 32767).astype('<i2')
    elif obj.dtype == np.int16:
        pcm = obj.astype('<i2')
    else:
        raise TypeError(f'cannot write dtype {obj.dtype} as 16-bit PCM')
    if pcm.ndim == 1:
        channels, data = 1, pcm
    elif pcm.ndim == 2:
        channels, data = pcm.shape[0], np.ascontiguousarray(pcm.T)
    else:
        raise ValueError(pcm.shape)
    payload = data.tobytes()
    block = 2 * channels
    header = (b'RIFF' + struct.pack('<I', 36 + len(payload)) + b'WAVE' + b'fmt ' +
              struct.pack('<IHHIIHH', 16, _FMT_PCM, channels, int(sample_rate), int(sample_rate) * block, block, 16) +
              b'data' + struct.pack('<I', len(payload)))
    path = Path(path)
    tmp = path.with_name(path.name + '.part')                   # a killed job never leaves a truncated wav behind
    with open(tmp, 'wb') as f:
        f.write(header)
        f.write(payload)
    tmp.replace(path)

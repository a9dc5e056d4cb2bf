"""CPU ORACLE, TEST INFRASTRUCTURE ONLY: numpy restatement of `WaveBox::load`
(reference src/synth/sample.rs:32-69), i.e. what a Sample module holds after "Load Sample...".

The byte-level WAV parsing lives in a third-party crate that is not under /root/reference:
**hound 3.5.1** (Cargo.lock).  Its published behaviour, as used by the reference's call site
(`WavReader::new(Cursor::new(data))`, `reader.spec()`, `reader.into_samples()`), is restated here:
RIFF/WAVE container, chunks walked until `data`, `fmt ` with tag 1 (PCM -> SampleFormat::Int),
3 (IEEE float, 32 bit -> SampleFormat::Float) or 0xFFFE (extensible; sub-format GUID decides),
8-bit PCM is unsigned in the file and read as `u8 - 128`, 16/24/32-bit PCM little-endian signed,
samples interleaved by channel.  The reference has no test that loads a WAV file => **parity
unpinned** by reference vectors; tests/test_sample.py pins this restatement against Python's own
`wave` module instead.

Reference conversions (sample.rs:49-54): 8 bit `x / 128`, 16 bit `x / 32768`, 24 bit
`I24::to_float_sample` (cpal 0.15.3 / dasp_sample: `x / 8388608`), anything else -> DecodeError
*after* `samples.clear()`; only channel 0 is kept (`idx % channels == 0`, :42,60).
"""
import struct

import numpy as np

_PCM_GUID_TAIL = bytes.fromhex("000000001000800000aa00389b71")


class WavError(ValueError):
    """hound's Err(..) before the reference touches the WaveBox: it stays as it was."""


class WavUnsupported(ValueError):
    """The reference's DecodeError (sample.rs:53): raised after `samples.clear()`, so the WaveBox
    is left EMPTY with its old sample rate and `new` not set."""


def parse(data):
    """-> (format 'int' | 'float', channels, sample_rate, bits, bytes_per_sample, payload bytes)"""
    data = bytes(data)
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise WavError("no RIFF/WAVE header")
    off, fmt = 12, None
    while True:
        if off + 8 > len(data):
            raise WavError("no data chunk")
        kind, size = data[off:off + 4], struct.unpack_from("<I", data, off + 4)[0]
        off += 8
        if kind == b"fmt ":
            if size < 16 or off + size > len(data):
                raise WavError("short fmt chunk")
            tag, ch, rate, _byte_rate, align, bits = struct.unpack_from("<HHIIHH", data, off)
            if ch == 0:
                raise WavError("zero channels")
            bps = align // ch
            if tag == 1:
                kind_ = "int"
            elif tag == 3:
                if bits != 32:
                    raise WavError("IEEE float must be 32 bit")
                kind_ = "float"
            elif tag == 0xFFFE:
                if size < 40:
                    raise WavError("short extensible fmt chunk")
                valid_bits = struct.unpack_from("<H", data, off + 18)[0]
                guid = data[off + 24:off + 40]
                if guid[2:] != _PCM_GUID_TAIL or guid[:2] not in (b"\x01\x00", b"\x03\x00"):
                    raise WavError("unknown sub-format")
                kind_ = "int" if guid[:2] == b"\x01\x00" else "float"
                if valid_bits != 8 * bps:
                    raise WavError("valid bits differ from the container size")
                bits = valid_bits
                if kind_ == "float" and bits != 32:
                    raise WavError("IEEE float must be 32 bit")
            else:
                raise WavError("unsupported format tag")
            if bits not in (8, 16, 24, 32) or bps * 8 != bits:
                raise WavError("unsupported sample size")
            fmt = (kind_, ch, rate, bits, bps)
        elif kind == b"data":
            if fmt is None:
                raise WavError("data before fmt")
            return fmt + (data[off:off + size], size)
        off += size + (size & 1)


def load(data):
    """WaveBox::load -> (samples f32 [frames], sample_rate f32)."""
    kind, ch, rate, bits, bps, payload, size = parse(data)
    if kind == "int" and bits == 32:
        raise WavUnsupported("32-bit integer PCM (sample.rs:53)")
    if len(payload) < size:
        raise WavUnsupported("data chunk is truncated")  # the reference's `s.unwrap()` panics here
    n = size // bps
    raw = np.frombuffer(payload, dtype=np.uint8, count=n * bps).reshape(n, bps)
    if kind == "float":
        x = raw.copy().view("<f4").reshape(n)
    elif bits == 8:
        x = (raw[:, 0].astype(np.int32) - 128).astype(np.float32) / np.float32(128.0)
    elif bits == 16:
        x = raw.copy().view("<i2").reshape(n).astype(np.float32) / np.float32(32768.0)
    else:
        v = raw[:, 0].astype(np.int32) | (raw[:, 1].astype(np.int32) << 8) | (raw[:, 2].astype(np.int32) << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = v.astype(np.float32) / np.float32(8388608.0)
    return np.ascontiguousarray(x[::ch], dtype=np.float32), np.float32(rate)

"""Sample module + WAV (src/synth/sample.rs:32-69 load, :192-240 calc).
CPU part: the oracle's Sample against an independent per-sample restatement written here from the
reference source; the numpy restatement of WaveBox::load (oracle/wav.py) against Python's own `wave`
module; the product's host-side WAV decode/encode and table calls through the C ABI against both; the
glibc exp2f restatement the device uses (dsp.cuh exp2f_glibc), replayed in numpy, against glibc.
The reference has no test for any of this => parity unpinned by reference vectors.
The GPU part lives in test_gpu_parity.py."""
import io
import struct
import wave

import numpy as np
import pytest

f32 = np.float32


def sample_reference(gate, cv, table, table_rate, sr, exp2f, new_at=()):
    """sample.rs:212-238 sample by sample.  `new_at`: sample indices (block starts) where a load
    happened just before (WaveBox.new)."""
    n = len(gate)
    out = np.zeros(n, f32)
    pos, playing, last = f32(0), False, True
    ratio = f32(table_rate) / f32(sr)
    for i in range(n):
        if i in new_at:
            pos, playing = f32(0), False
        above = gate[i] > 0
        trig = above and not last
        last = above
        if trig:
            pos, playing = f32(0), True
        idx = 0 if not (pos > 0) else int(pos)  # `as usize`: NaN / negative -> 0
        if idx >= len(table):
            pos, playing, idx = f32(0), False, 0
        out[i] = table[idx] if len(table) else 0.0
        if playing:
            e = exp2f(cv[i]) if cv is not None else f32(1)
            pos = f32(pos + f32(ratio * e))
    return out


def _player(orc, B, gate_hz, cv_depth, channels=3):
    """gate oscillator (square) -> Sample.gate; slow sine * depth -> Sample.cv; channels: player, gate, cv."""
    from srack_b200.patches import hz_to_val
    op = orc.OraclePatch(48000, B, channels)
    gate = op.module_create("OSCILLATOR")
    lfo = op.module_create("OSCILLATOR")
    depth = op.module_create("MULTIPLY")
    smp = op.module_create("SAMPLE")
    out = op.module_create("OUTPUT")
    op.set_param(gate, 0, hz_to_val(gate_hz))
    op.set_param(lfo, 0, hz_to_val(2.7))
    op.set_param(depth, 0, cv_depth or 0.0)
    op.connect(depth, 0, lfo, 0)
    op.connect(smp, 0, gate, 1)
    if cv_depth is not None:
        op.connect(smp, 1, depth, 0)
    op.connect(out, 0, smp, 0)
    op.connect(out, 1, gate, 1)
    op.connect(out, 2, depth, 0)
    return op, smp


@pytest.mark.parametrize("B", [64, 1024])
@pytest.mark.parametrize("cv_depth", [None, 0.0, 1.25])
def test_oracle_sample_matches_the_restatement(srk, orc, B, cv_depth):
    table, rate = srk.patches.sampler_wave(3000, 22050.0)
    op, smp = _player(orc, B, 9.0, cv_depth)
    op.set_sample(smp, table, rate)
    st, _ = op.render(1, 12000)
    exp2f = lambda x: f32(orc.lib().orc_exp2f(float(x)))
    ref = sample_reference(st[1, :, 0], st[2, :, 0] if cv_depth is not None else None, table, rate, 48000, exp2f,
                           new_at=(0,))
    assert (st[0, :, 0] == ref).all()
    assert np.abs(ref).max() > 0.3 and (ref == 0).sum() < len(ref) // 2
    # a load in the middle of playback rewinds at the next block and waits for the next gate edge
    table2 = (table[::-1] * f32(0.5)).copy()
    op.set_sample(smp, table2, 44100.0)
    st2, _ = op.render(1, 12000)
    # (state continues: compare against a restatement of both halves with the table swapped by hand)
    full_gate = np.concatenate([st[1, :, 0], st2[1, :, 0]])
    assert st2[0, 0, 0] == table2[0]
    first_edge = np.flatnonzero((full_gate[12000:] > 0) & ~(np.concatenate([[full_gate[11999]], full_gate[12000:-1]]) > 0))
    assert len(first_edge) and (st2[0, :first_edge[0], 0] == table2[0]).all()


def test_oracle_sample_edge_cases(srk, orc):
    # empty table (Default WaveBox): silence, and `pos as usize >= 0` keeps it parked
    op, smp = _player(orc, 256, 50.0, 0.5)
    st, _ = op.render(1, 3000)
    assert (st[0] == 0).all()
    # one-sample table, rate ratio > 1: every sample after the trigger re-parks at index 0
    op.set_sample(smp, np.array([0.25], f32), 96000.0)
    st, _ = op.render(1, 3000)
    assert (st[0] == 0.25).all()
    # no gate connected: never plays, sits on samples[0]
    op2 = orc.OraclePatch(48000, 128, 1)
    s2 = op2.module_create("SAMPLE")
    o2 = op2.module_create("OUTPUT")
    op2.connect(o2, 0, s2, 0)
    op2.set_sample(s2, np.array([0.5, -1.0, 1.0], f32), 48000.0)
    assert (op2.render(1, 500)[0] == 0.5).all()


# ---- glibc exp2f as the device restates it (dsp.cuh exp2f_glibc), replayed in numpy f64 ----------
def _exp2f_tab():
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    ln2 = Decimal(2).ln()
    t = []
    for i in range(32):
        bits = struct.unpack("<Q", struct.pack("<d", float((Decimal(i) / 32 * ln2).exp())))[0]
        t.append((bits - (i << 47)) & 0xFFFFFFFFFFFFFFFF)
    return np.array(t, dtype=np.uint64)


def exp2f_restated(x):
    x = np.asarray(x, dtype=f32)
    xd = x.astype(np.float64)
    shift = np.float64(float.fromhex("0x1.8p+52")) / 32.0
    kd = xd + shift
    ki = kd.view(np.uint64)
    kd = kd - shift
    with np.errstate(invalid="ignore"):
        r = xd - kd
    with np.errstate(over="ignore"):
        t = _exp2f_tab()[(ki & np.uint64(31)).astype(np.int64)] + (ki << np.uint64(47))
    s = t.view(np.float64)
    z = float.fromhex("0x1.c6af84b912394p-5") * r + float.fromhex("0x1.ebfce50fac4f3p-3")
    y = float.fromhex("0x1.62e42ff0c52d6p-1") * r + 1.0
    y = (z * (r * r) + y) * s
    with np.errstate(over="ignore", under="ignore"):
        out = y.astype(f32)
    out = np.where(x >= 128.0, f32(np.inf), out)
    out = np.where(x <= -150.0, f32(0), out)
    return np.where(np.isnan(x), x, out).astype(f32)


def test_exp2f_restatement_is_glibc_bit_for_bit(srk, orc):
    """The device table in dsp.cuh is these 32 words; the algorithm equals glibc's for every input tried
    (checked exhaustively over all 2^32 floats when it was written)."""
    import re, os
    src = open(os.path.join(os.path.dirname(srk.__file__), "csrc", "dsp.cuh")).read()
    body = src[src.index("kExp2fTab[32] = {"):]
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", body[:body.index("};")])]
    assert words == _exp2f_tab().tolist()
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-4, 4, 200000), rng.uniform(-150, 128, 50000), rng.normal(0, 1e-3, 20000),
                         [0.0, -0.0, 1.0, -1.0, 127.99999, 128.0, -126.0, -149.0, -149.5, -150.0, -151.0, np.inf, -np.inf,
                          1e30, -1e30, 1e-45]]).astype(f32)
    import ctypes
    ex = orc.lib().orc_exp2f
    want = np.array([ex(ctypes.c_float(float(v))) for v in xs], dtype=f32)
    got = exp2f_restated(xs)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert np.isnan(exp2f_restated(np.array([np.nan], f32))[0])


# ---- WAV -------------------------------------------------------------------------------------
def _pcm_wav(frames, channels, width, rate):
    """`frames`: int array [n][channels] already in the file's integer range -> WAV bytes via `wave`."""
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(rate)
        if width == 1:
            raw = (frames + 128).astype(np.uint8).tobytes()
        elif width == 2:
            raw = frames.astype("<i2").tobytes()
        else:
            u = frames.astype(np.int32) & 0xFFFFFF
            raw = np.stack([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=-1).astype(np.uint8).tobytes()
        w.writeframes(raw)
    return buf.getvalue()


def _float_wav(x, channels, rate, extensible=False, extra_chunk=True):
    data = np.asarray(x, "<f4").tobytes()
    if extensible:
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, channels, rate, rate * channels * 4, channels * 4, 32, 22, 32, 0)
        fmt += b"\x03\x00" + bytes.fromhex("000000001000800000aa00389b71")
    else:
        fmt = struct.pack("<HHIIHH", 3, channels, rate, rate * channels * 4, channels * 4, 32)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if extra_chunk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"  # odd size + pad byte
    chunks += b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def _cases():
    rng = np.random.default_rng(5)
    for width in (1, 2, 3):
        for ch in (1, 2, 3):
            lim = 1 << (8 * width - 1)
            frames = rng.integers(-lim, lim, size=(257, ch))
            frames[0, :] = -lim
            frames[1, :] = lim - 1
            yield f"pcm{8 * width}x{ch}", _pcm_wav(frames, ch, width, 22050 + width), frames[:, 0] / float(lim), 22050 + width
    x = rng.normal(0, 0.5, size=(300, 2)).astype(f32)
    yield "float", _float_wav(x, 2, 44100), x[:, 0], 44100
    yield "float-extensible", _float_wav(x, 2, 96000, extensible=True, extra_chunk=False), x[:, 0], 96000


def test_wav_restatement_against_pythons_wave_module(orc):
    from oracle import wav
    for name, data, want, rate in _cases():
        got, got_rate = wav.load(data)
        assert got.dtype == f32 and got_rate == f32(rate), name
        assert (got == want.astype(f32)).all(), name  # x / 2^(bits-1) is exact in f32 for <= 24 bits


def test_product_wav_decode_equals_the_restatement(srk, orc):
    from oracle import wav
    p = srk.Patch()
    smp = p.add_module("Sample")
    assert smp.get_kind() == "SAMPLE" and smp.get_num_inputs() == 2 and smp.get_num_outputs() == 1
    assert [smp.get_input_label(i) for i in range(2)] == ["Gate", "CV"] and smp.get_output_label(0) is None
    got, rate = smp.get_sample()
    assert got.size == 0 and rate == 0.0  # WaveBox::default()
    for name, data, _, _ in _cases():
        smp.load_wav(data)
        got, rate = smp.get_sample()
        want, want_rate = wav.load(data)
        assert rate == want_rate and (got.view(np.uint32) == want.view(np.uint32)).all(), name
    # malformed header: Err before the WaveBox is touched -> the table stays
    for bad in (b"", b"RIFF\x00\x00\x00\x00WAVX", data[:20], data.replace(b"fmt ", b"fmx ")):
        with pytest.raises(srk.SrackError) as e:
            smp.load_wav(bad)
        assert e.value.status == srk.STATUS["ERR_ARG"]
        with pytest.raises(wav.WavError):
            wav.load(bad)
        assert smp.get_sample()[0].size == want.size
    # 32-bit integer PCM: the reference's DecodeError comes after samples.clear() (sample.rs:36,53)
    int32 = _pcm_wav(np.zeros((4, 1), np.int64), 1, 2, 8000).replace(struct.pack("<HH", 2, 16), struct.pack("<HH", 4, 32))
    with pytest.raises(wav.WavUnsupported):
        wav.load(int32)
    with pytest.raises(srk.SrackError) as e:
        smp.load_wav(int32)
    assert e.value.status == srk.STATUS["ERR_UNSUPPORTED"]
    got, rate = smp.get_sample()
    assert got.size == 0 and rate == want_rate  # emptied, old rate kept
    # truncated data chunk
    smp.load_wav(data)
    with pytest.raises(srk.SrackError):
        smp.load_wav(data[:-10])
    assert smp.get_sample()[0].size == 0
    # other kinds have no table
    with pytest.raises(srk.SrackError) as e:
        p.add_module("Oscillator").set_sample(np.zeros(4, f32), 48000)
    assert e.value.status == srk.STATUS["ERR_KIND"]


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_wav_export_round_trip(srk, orc, tmp_path, bits):
    from oracle import wav
    rng = np.random.default_rng(bits)
    mix = np.clip(rng.normal(0, 0.4, size=(2, 5000)), -1.5, 1.5).astype(f32)
    mix[:, 0] = [1.0, -1.0]
    mix[:, 1] = [2.0, -2.0]  # clamps in PCM
    path = tmp_path / f"render{bits}.wav"
    srk.write_wav(path, mix, 48000, bits)
    data = path.read_bytes()
    got0, rate = wav.load(data)  # channel 0 through the WaveBox::load restatement
    assert rate == 48000.0
    if bits == 32:
        assert (got0 == mix[0]).all()
        kind, ch, r, b, bps, payload, size = wav.parse(data)
        assert (kind, ch, r, b) == ("float", 2, 48000, 32)
        assert (np.frombuffer(payload, "<f4").reshape(-1, 2).T == mix).all()
    else:
        with wave.open(io.BytesIO(data)) as w:  # and Python's own reader accepts the file
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, bits // 8, 48000, 5000)
        scale = float(1 << (bits - 1))
        want = np.clip(np.rint(mix[0].astype(np.float64) * scale), -scale, scale - 1) / scale
        assert (got0 == want.astype(f32)).all()
    # the exported file loads back into a Sample module
    smp = srk.Patch().add_module("Sample")
    smp.load_wav(data)
    assert (smp.get_sample()[0] == got0).all()


def test_sampler_program_shape(srk):
    """A load does not invalidate the plan; the table descriptor rides in the program image."""
    p = srk.Patch()
    h = srk.patches.sampler(p, 64)
    p.plan()
    ops = [i["op"] for i in p.program(64)[0]]
    assert ops.count("SAMPLE") == 1
    h["sample"].set_sample(np.zeros(10, f32), 8000.0)
    assert [i["op"] for i in p.program(64)[0]] == ops  # would raise NOT_PLANNED
    ops1 = [i["op"] for i in p.program(1 << 16)[0]]   # one-warp schedule
    assert ops1.count("SAMPLE") == 1

"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs and against the committed golden fixtures; plus the
size-independent properties the domain offers (determinism, chunk invariance, mix == sum of
stems, shard invariance) at BASELINE.json's full sizes.

Tolerance (BASELINE.json north_star): |gpu - ref| <= 1e-5 * max(|ref|, 1) per sample.
Held to more than that: bit-identical everywhere.  Every libm function on the path is glibc's on
the device too (s-rack_b200/csrc/libm_glibc.cuh: exp2, powf, exp2f operation by operation; the
sine port's `(float) sin(x)` from CUDA's sin except next to an f32 rounding tie, where it goes
through the restatement of glibc's)."""
import os
import random

import numpy as np
import pytest

from util import assert_mix_parity, assert_parity, build_both, parity_stats

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def render_pair(srk, orc, builder, V, N, B=1024, channels=2, **kw):
    gp, op, _, _ = build_both(srk, orc, builder, V, buffer_size=B, channels=channels, **kw)
    gp.plan()
    g_st, g_mix = gp.render(V, N, stems=True, mix=True)
    o_st, o_mix = op.render(V, N)
    return gp, op, g_st, g_mix, o_st, o_mix


def test_cfg1_single_sine(srk, orc, cuda_device):
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.cfg1, 1, 48000)
    assert_parity(g, o, exact=True, what="cfg1")  # SURVEY §8d asks for <= 1 f32 ulp on the sine path; the bits are equal
    assert_mix_parity(g_mix, o_mix, 1)
    # dco_tests::produces_440 through the GPU: sr = 1760, 17-sample buffers, phase carries over
    p = srk.Patch(srk.AudioConfig(440 * 4, 17, 2))
    srk.patches.cfg1(p, 1)
    p.plan()
    buf = p.execute(1, stems=True)[0][0, :, 0]
    assert buf[0] == 0.0 and abs(buf[1] - 1) < 1e-5 and abs(buf[2]) < 1e-5 and abs(buf[3] + 1) < 1e-5 and abs(buf[4]) < 1e-5
    assert abs(p.execute(1, stems=True)[0][0, 0, 0] - 1.0) < 1e-5


def test_cfg2_subtractive_is_bit_exact(srk, orc, cuda_device):
    # saw/square oscillators with host-computed delta, filter, ADSR, VCA: no libm on the path
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.cfg2, 100, 26000)
    assert np.abs(o).max() > 0.1
    assert_parity(g, o, exact=True, what="cfg2")
    assert_mix_parity(g_mix, o_mix, 100, what="cfg2 mix")


@pytest.mark.parametrize("name,V,N", [("cfg1", 40, 48000), ("cfg3", 99, 8192), ("cfg4", 33, 16000)])
def test_sine_port_through_the_restated_glibc_sin(srk, orc, cuda_device, monkeypatch, schedule, name, V, N):
    """With the tie band at its widest EVERY sine sample goes through the device's restatement of glibc's sin
    (libm_glibc.cuh: sin_glibc), the route a default build takes for 6e-8 of the samples: same bits as the oracle's
    libm call."""
    if schedule == "interp":
        pytest.skip("the band is part of a fused kernel's source; the interpreter kernels are built with the default")
    monkeypatch.setenv("SRK_FUSED", "1")
    monkeypatch.setenv("SRK_FUSED_DEFINE", f"SRK_SIN_TIE_BAND={0x10000000}")
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.CONFIGS[name][0], V, N)
    assert np.abs(o).max() > 0.1
    assert_parity(g, o, exact=True, what=f"{name}, every sine restated")


def test_cfg3_fm(srk, orc, cuda_device):
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.cfg3, 99, 8192)
    assert_parity(g, o, exact=True, what="cfg3")  # sin and exp2 are glibc's operation by operation (libm_glibc.cuh)
    assert_mix_parity(g_mix, o_mix, 99)


@pytest.mark.parametrize("B", [1, 7, 256, 1024])
def test_cfg3b_feedback_delay_equals_buffer_size(srk, orc, cuda_device, B):
    N = 4 * 1024 if B > 7 else 1022 // B * B
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.cfg3b, 70, N, B=B)
    assert len(gp.plan_cuts()) == 1
    assert_parity(g, o, exact=True, what=f"cfg3b B={B}")


def test_cfg4_full_subtractive(srk, orc, cuda_device):
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.cfg4, 70, 26000)
    assert np.abs(o).max() > 0.1
    assert_parity(g, o, exact=True, what="cfg4")
    assert_mix_parity(g_mix, o_mix, 70)


@pytest.mark.parametrize("idx", range(8))
def test_cfg5_graphs(srk, orc, cuda_device, idx):
    builder = srk.patches.CFG5_GRAPHS[idx]
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, builder, 37, 14336)
    assert np.abs(o).max() > 0.01
    assert_parity(g, o, what=f"cfg5[{idx}]")
    assert_mix_parity(g_mix, o_mix, 37)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg3b", "cfg4", "sequenced", "sampler"])
def test_against_committed_golden_vectors(srk, cuda_device, name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    V, N, B = int(g["n_voices"]), int(g["n_samples"]), int(g["buffer_size"])
    p = srk.Patch(srk.AudioConfig(48000, B, 2))
    (getattr(srk.patches, name) if name in ("sequenced", "sampler") else srk.patches.CONFIGS[name][0])(p, V)
    p.plan()
    st, _ = p.render(V, N, stems=True)
    assert_parity(st, g["stems"], exact=(name in ("cfg2", "sampler")), what=f"golden {name}")


def test_every_module_kind_and_port(srk, orc, cuda_device):
    """Sync input, antialiasing off, band/high-pass taps, VCA negative, mixer gains, Subtract,
    Non-Linear with constant and with a second input, unconnected-input defaults."""
    P = srk.PARAM

    def build(b, n_voices, seed=0):
        master = b.module_create("OSCILLATOR")
        slave = b.module_create("OSCILLATOR")
        raw = b.module_create("OSCILLATOR")
        filt = b.module_create("MOOG_FILTER")
        sub = b.module_create("SUBTRACT")
        nl = b.module_create("NON_LINEAR")
        nl2 = b.module_create("NON_LINEAR")
        vca = b.module_create("VCA")
        mix = b.module_create("MONO_MIXER")
        add = b.module_create("ADD")
        adsr = b.module_create("ADSR")
        out = b.module_create("OUTPUT")
        b.set_param(master, P["OSC_VAL"], -2.0)
        b.set_param_per_voice(slave, P["OSC_VAL"], srk.patches._u(7, 0, n_voices, -1.0, 1.5))
        b.connect(slave, 1, master, 1)            # hard sync from the master's square
        b.set_param(raw, P["OSC_ANTIALIASING"], 0.0)
        b.set_param(raw, P["OSC_VAL"], -0.5)
        b.connect(filt, 0, slave, 2)
        b.set_param(filt, P["MOOG_RES"], 0.8)
        b.connect(sub, 0, filt, 1)                # bandpass - highpass
        b.connect(sub, 1, filt, 2)
        b.connect(nl, 0, sub, 0)
        b.set_param(nl, P["MATH_CONSTANT"], 0.7)  # signed |x|^0.7
        b.connect(nl2, 0, raw, 1)                 # square ^ (something varying)
        b.connect(nl2, 1, add, 0)
        b.set_param(add, P["MATH_CONSTANT"], 1.25)  # (None, None) -> 0 + constant
        b.connect(vca, 0, nl, 0)
        b.connect(vca, 1, master, 0)              # sine CV, negative half multiplies too
        b.set_param(vca, P["VCA_NEGATIVE"], 1.0)
        b.connect(mix, 0, vca, 0)
        b.connect(mix, 2, nl2, 0)
        b.connect(mix, 3, adsr, 0)                # ADSR with no gate: stays at 0
        b.set_param(mix, P["MIXER_GAIN0"], 0.9)
        b.set_param(mix, P["MIXER_GAIN2"], 0.1)
        b.connect(out, 0, mix, 0)
        b.connect(out, 1, raw, 2)                 # non-antialiased saw
        return {}

    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, build, 45, 6000)
    assert np.abs(o[0]).max() > 0.05 and np.abs(o[1]).max() > 0.5
    assert_parity(g[1], o[1], exact=True, what="naive saw")
    assert_parity(g[0], o[0], what="everything else")


def test_more_than_four_and_single_channel_outputs(srk, orc, cuda_device):
    for channels in (1, 6):
        def build(b, n_voices, seed=0):
            osc = b.module_create("OSCILLATOR")
            out = b.module_create("OUTPUT")
            for c in range(channels):
                if c != 2:
                    b.connect(out, c, osc, c % 3)
            return {}
        gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, build, 33, 2000, channels=channels)
        assert g.shape == (channels, 2000, 33)
        assert_parity(g, o, what=f"{channels} channels")
        assert_mix_parity(g_mix, o_mix, 33)
        if channels > 2:
            assert (g[2] == 0).all() and (g_mix[2] == 0).all()  # unconnected channel -> zeros (output.rs:56)


def test_adsr_corner_cases(srk, orc, cuda_device):
    """a_sec == 0 (1/0 = +inf -> straight to Decay, adsr.rs:152-156), retrigger during attack/decay/
    release, gate high at sample 0 is not a transition (detector starts `last = true`, synth.rs:283)."""
    P = srk.PARAM
    for a_sec, gate_hz, d, s_, r in [(0.0, 40.0, 0.004, 0.25, 0.5), (0.02, 90.0, 0.5, 0.5, 0.5), (0.001, 25.0, 0.002, 0.7, 0.05)]:
        def build(b, n_voices, seed=0):
            gate = b.module_create("OSCILLATOR")
            const = b.module_create("ADD")
            gsum = b.module_create("ADD")
            adsr = b.module_create("ADSR")
            out = b.module_create("OUTPUT")
            b.set_param(gate, P["OSC_VAL"], srk.patches.hz_to_val(gate_hz))
            b.set_param_per_voice(const, P["MATH_CONSTANT"], np.linspace(-0.5, 1.5, n_voices).astype(np.float32))
            b.connect(gsum, 0, gate, 1)
            b.connect(gsum, 1, const, 0)          # per-voice gate offset: some voices start high / never fall
            b.connect(adsr, 0, gsum, 0)
            for pid, v in zip(("ADSR_A_SEC", "ADSR_D_SEC", "ADSR_S_VAL", "ADSR_R_SEC"), (a_sec, d, s_, r)):
                b.set_param(adsr, P[pid], v)
            b.connect(out, 0, adsr, 0)
            return {}
        gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, build, 41, 9000, channels=1)
        assert o.max() > 0.2
        assert_parity(g, o, exact=True, what=f"adsr a={a_sec}")


def _fuzz_patch(rng, n_modules, with_sample=False):
    from test_planner import KINDS, N_IN, N_OUT
    if with_sample:
        KINDS, N_IN, N_OUT = KINDS + ["SAMPLE", "SAMPLE"], dict(N_IN, SAMPLE=2), dict(N_OUT, SAMPLE=1)
    kinds = [rng.choice(KINDS) for _ in range(n_modules)] + ["OUTPUT"]
    # make sure something audible reaches the output
    kinds[0] = "OSCILLATOR"
    wires = [(len(kinds) - 1, 0, 0, rng.randrange(3))]
    for sink, k in enumerate(kinds[:-1]):
        for i in range(N_IN[k]):
            if rng.random() < 0.55:
                src = rng.randrange(len(kinds) - 1)
                if src != sink:
                    wires.append((sink, i, src, rng.randrange(N_OUT[kinds[src]])))
    src = rng.randrange(len(kinds) - 1)
    wires.append((len(kinds) - 1, 1, src, rng.randrange(N_OUT[kinds[src]])))
    return kinds, wires


@pytest.mark.parametrize("seed", range(36))
def test_random_patches(srk, orc, cuda_device, seed):
    """Random (often cyclic) graphs over every module kind with random parameters: exercises the
    patch compiler (wire liveness, rings, unconnected-input defaults) against the oracle."""
    rng = random.Random(1000 + seed)
    kinds, wires = _fuzz_patch(rng, rng.randrange(3, 12), with_sample=seed >= 24)  # seeds 24..: Sample players too
    B = rng.choice([1, 5, 64, 1024])
    V = rng.choice([1, 31, 33, 64])
    N = 64 * max(1, 1024 // 64) if B > 64 else 640 // B * B
    n_par = dict(OSCILLATOR=1, ADSR=4, MOOG_FILTER=3, MONO_MIXER=4, ADD=1, SUBTRACT=1, MULTIPLY=1, NON_LINEAR=1)
    ranges = dict(OSCILLATOR=(-3, 3), ADSR=(0.0, 0.01), MOOG_FILTER=(0.05, 0.9), MONO_MIXER=(0, 1), ADD=(-1, 1),
                  SUBTRACT=(-1, 1), MULTIPLY=(-1, 1), NON_LINEAR=(0.5, 2.0))  # (sequencers: tables below)
    pvals = {(m, pid): (rng.random() < 0.5, srk.patches._u(seed, 17 * m + pid, V, *ranges[k]))
             for m, k in enumerate(kinds) if k in n_par for pid in range(n_par[k])}

    tables = {}
    for m, k in enumerate(kinds):  # sequencers get a random table (and the grid a random scale)
        steps = rng.choice([1, 3, 16, 64])
        if k == "GRID_SEQUENCER":
            tables[m] = np.array([-1 if rng.random() < 0.3 else srk.grid_cell(rng.randrange(24), rng.random() < 0.5)
                                  for _ in range(steps)], dtype=np.int32)
        elif k == "PATTERN_SEQUENCER":
            tables[m] = np.array([[rng.choice([-1, 0, 1]) for _ in range(steps)] for _ in range(8)], dtype=np.int32)
    waves = {m: (np.random.default_rng(seed * 100 + m).uniform(-1, 1, rng.choice([0, 1, 97, 4000])).astype(np.float32),
                 rng.choice([8000.0, 44100.0, 192000.0])) for m, k in enumerate(kinds) if k == "SAMPLE"}

    def build(b, n_voices, seed=0):
        mods = [b.module_create(k) for k in kinds]
        for sink, i, src, port in wires:
            b.connect(mods[sink], i, mods[src], port)
        for m, cells in tables.items():
            b.set_sequence(mods[m], cells)
        for m, (wave, rate) in waves.items():
            b.set_sample(mods[m], wave, rate)
        for (m, pid), (per_voice, vals) in pvals.items():
            if per_voice:
                b.set_param_per_voice(mods[m], pid, vals)
            else:
                b.set_param(mods[m], pid, float(vals[0]))
        return {}

    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, build, V, N, B=B)
    assert [kinds.index(m.get_kind()) >= 0 for m in gp.plan()]
    # Every libm function on the path -- f64 sin (sine port), f64 exp2 (V/oct), powf (Non-Linear), exp2f (Sample) -- is
    # glibc's operation by operation on the device (libm_glibc.cuh): ANY random patch must match the oracle bit for bit,
    # feedback, NaN and infinities included.
    same = (g.view(np.uint32) == o.view(np.uint32)) | (np.isnan(g) & np.isnan(o))
    assert same.all(), f"fuzz {seed}: {int((~same).sum())} of {same.size} samples differ ({parity_stats(g, o)})"


def test_determinism_reset_and_chunk_invariance(srk, orc, cuda_device):
    V, N = 130, 9000
    p = srk.Patch()
    srk.patches.cfg4(p, V)
    p.plan()
    a_st, a_mix = p.render(V, N, stems=True)
    p.reset()
    b_st, b_mix = p.render(V, N, stems=True)
    assert (a_st.view(np.uint32) == b_st.view(np.uint32)).all()  # run to run bit-identical
    assert (a_mix.view(np.uint32) == b_mix.view(np.uint32)).all()
    p.reset()
    chunks, done = [], 0
    mixes = []
    for n in (1, 15, 16, 17, 1000, 4096, N - 5145):  # state persists across calls like across blocks
        st, mx = p.render(V, n, stems=True)
        chunks.append(st)
        mixes.append(mx)
        done += n
    assert done == N
    c_st = np.concatenate(chunks, axis=1)
    assert (a_st.view(np.uint32) == c_st.view(np.uint32)).all()
    # the mixdown adds a sample's voices in an order fixed by the absolute sample index: chunking does not move a bit
    assert (a_mix.view(np.uint32) == np.concatenate(mixes, axis=1).view(np.uint32)).all()


def test_chunked_feedback_patch_keeps_ring_phase(srk, orc, cuda_device):
    V, B = 40, 64
    p = srk.Patch(srk.AudioConfig(48000, B, 2))
    srk.patches.cfg3b(p, V)
    p.plan()
    whole = p.render(V, 1000, stems=True)[0]
    p.reset()
    parts = np.concatenate([p.render(V, n, stems=True)[0] for n in (3, 61, 64, 100, 772)], axis=1)
    assert (whole.view(np.uint32) == parts.view(np.uint32)).all()


def test_voice_offset_sharding_is_exact(srk, orc, cuda_device):
    """Rendering voices [0, V) in one go or as two shards with voice_offset gives the same stems
    bit for bit (incl. the per-voice noise key) and the shard mixes add up to the full mix."""
    V, N = 101, 13500
    full = srk.Patch()
    srk.patches.cfg4(full, V)
    full.plan()
    f_st, f_mix = full.render(V, N, stems=True)
    parts, mixes = [], []
    for rank in range(2):
        off, cnt = srk.shard.voice_range(V, rank, 2)
        p = srk.Patch()
        srk.patches.cfg4(p, V)
        p.plan()
        st, mx = p.render(cnt, N, voice_offset=off, stems=True)
        parts.append(st)
        mixes.append(mx)
    assert (np.concatenate(parts, axis=2).view(np.uint32) == f_st.view(np.uint32)).all()
    assert np.abs(mixes[0] + mixes[1] - f_mix).max() <= 1e-5 * np.sqrt(V)


def test_mix_equals_sum_of_stems_and_mix_only_mode(srk, orc, cuda_device):
    V, N = 300, 14000
    p = srk.Patch()
    srk.patches.cfg2(p, V)
    p.plan()
    st, mx = p.render(V, N, stems=True, mix=True)
    assert_mix_parity(mx, st.astype(np.float64).sum(axis=2), V)
    p.reset()
    _, mx2 = p.render(V, N, stems=False, mix=True)
    assert (mx.view(np.uint32) == mx2.view(np.uint32)).all()


def test_per_voice_array_too_short_is_an_error(srk, cuda_device):
    p = srk.Patch()
    srk.patches.cfg2(p, 8)
    p.plan()
    with pytest.raises(srk.SrackError) as e:
        p.render(16, 64)
    assert e.value.status == srk.STATUS["ERR_SIZE"]


def test_baseline_size_cfg2_properties(srk, orc, cuda_device):
    """BASELINE configs[1] at full size (4096 voices x 48000 samples), outputs kept in HBM:
    determinism, mix == sum(stems), boundedness, and per-voice parity for a sample of voices."""
    import torch

    V, N = 4096, 48000
    p = srk.Patch(device=0)
    srk.patches.cfg2(p, V)
    p.plan()
    stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0")
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    p.render_into(V, N, 0, stems.data_ptr(), mix.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    assert torch.isfinite(stems).all() and float(stems.abs().max()) <= 1.0  # VCA of a clamped filter
    assert float(stems.abs().max()) > 0.3
    ref_mix = stems.double().sum(dim=2)
    assert float((mix.double() - ref_mix).abs().max()) <= 1e-5 * max(np.sqrt(V), float(ref_mix.abs().max()))
    assert torch.equal(stems[0], stems[1])  # both channels read the same wire
    stems2 = torch.empty_like(stems)
    p.reset()
    p.render_into(V, N, 0, stems2.data_ptr(), None, device_out=True)
    torch.cuda.synchronize()
    assert torch.equal(stems, stems2)
    for v in (0, 1, 2047, 4095):
        op = orc.OraclePatch()
        srk.patches.cfg2(op, V)
        ref, _ = op.render(1, N, voice_offset=v)
        assert_parity(stems[:, :, v].cpu().numpy(), ref[:, :, 0], exact=True, what=f"cfg2 voice {v}")


def test_baseline_size_cfg3_sampled_parity(srk, orc, cuda_device):
    """BASELINE configs[2] (65536 voices x 48000): mix-only render at full size, stems for a voice
    slice rendered separately with voice_offset and compared with the oracle."""
    import torch

    V, N = 65536, 48000
    p = srk.Patch(device=0)
    srk.patches.cfg3(p, V)
    p.plan()
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    p.render_into(V, N, 0, None, mix.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    assert torch.isfinite(mix).all() and float(mix.abs().max()) <= V
    assert float(mix.abs().max()) > 10.0
    q = srk.Patch(device=0)
    srk.patches.cfg3(q, V)
    q.plan()
    off, cnt = 65536 - 48, 48
    st, _ = q.render(cnt, N, voice_offset=off, stems=True)
    op = orc.OraclePatch()
    srk.patches.cfg3(op, V)
    ref, _ = op.render(cnt, N, voice_offset=off)
    assert_parity(st, ref, exact=True, what="cfg3 voice slice")


def test_sequenced_patch(srk, orc, cuda_device):
    """Grid + Pattern sequencers driving a voice, per-voice tempo."""
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.sequenced, 67, 24000)
    assert np.abs(o[0]).max() > 0.05 and set(np.unique(o[1])) == {0.0, 1.0}
    assert_parity(g[1], o[1], exact=True, what="pattern sequencer sync output")
    assert_parity(g[0], o[0], exact=True, what="sequenced voice")  # the oscillator's CV goes through exp2: glibc's on the device too
    assert_mix_parity(g_mix, o_mix, 67)


def test_sequence_edit_keeps_voice_state(srk, orc, cuda_device):
    """The reference's ui() edits the table under the running module: counters, detectors, envelopes and
    filter state carry on."""
    V = 40
    gp = srk.Patch()
    gh = srk.patches.sequenced(gp, V)
    gp.plan()
    parts_g, parts_o = [], []
    for k in range(3):
        parts_g.append(gp.render(V, 7001, stems=True)[0])
        cells = np.array([srk.grid_cell((5 * k + i) % 19, i % 2 == 0) for i in range(8 + k)], dtype=np.int32)
        gp.set_sequence(gh["grid"], cells)
    # the oracle renders whole blocks only: one 7001-sample block per call, the same edits in between
    op2 = orc.OraclePatch(48000, 7001, 2)
    oh2 = srk.patches.sequenced(op2, V)
    for k in range(3):
        parts_o.append(op2.render(V, 7001)[0])
        cells = np.array([srk.grid_cell((5 * k + i) % 19, i % 2 == 0) for i in range(8 + k)], dtype=np.int32)
        op2.set_sequence(oh2["grid"], cells)
    g, o = np.concatenate(parts_g, axis=1), np.concatenate(parts_o, axis=1)
    assert_parity(g[1], o[1], exact=True, what="sync after edits")
    assert_parity(g[0], o[0], what="voice after edits")


def test_sampler_patch_is_bit_exact(srk, orc, cuda_device):
    """Sample module: the play position is an index path -- the rate goes through the
    device's restatement of glibc exp2f -- so anything short of equal bits would show as wrong table
    entries.  Channel 1 is the raw player, channel 0 the enveloped one."""
    for cv in (True, False):
        gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.sampler, 70, 26000, cv=cv)
        assert np.abs(o[1]).max() > 0.3
        assert_parity(g, o, exact=True, what=f"sampler cv={cv}")
        assert_mix_parity(g_mix, o_mix, 70)
    # many voices x a bent rate: ~1.7e7 exp2f results steering indices, one-warp schedule
    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, srk.patches.sampler, 2048, 8192)
    assert_parity(g, o, exact=True, what="sampler 2048 voices")


def test_loaded_srk_files_render_like_the_oracle(srk, orc, cuda_device):
    """a .srk file goes through the product's C++ loader on one side and through the
    schema-driven restatement onto the oracle on the other; the renders must agree.  The files carry
    mid-performance DSP state (oscillator phase, filter memory, an envelope in Decay, step counters, a
    playing sample): every voice starts from it; the port buffers in the file are ignored by both."""
    from oracle import srk_file as sf
    from test_srk_file import sequenced_file, subtractive_file
    for make, B, exact_ch in ((subtractive_file, 1024, ()), (sequenced_file, 512, (1,))):
        data = sf.dumps(make(B=B))
        gp = srk.Patch(srk.AudioConfig(48000, B, 2))
        assert gp.load_srk(data) == 0
        gp.plan()
        op = orc.OraclePatch(48000, B, 2)
        op.load_srk(data)
        V, N = 33, 16 * B  # (the 2 Hz gate of the first file opens at sample 12000)
        g, g_mix = gp.render(V, N, stems=True, mix=True)
        o, o_mix = op.render(V, N)
        assert np.abs(o).max() > 0.01
        assert_parity(g, o, exact=True, what=make.__name__)
        gp.reset()                                        # reset() returns to the loaded state
        assert (gp.render(V, N, stems=True, mix=False)[0].view(np.uint32) == g.view(np.uint32)).all()
        # save -> load -> save -> load (the list reversed twice = the same order) carries the state along
        gp2 = srk.Patch(srk.AudioConfig(48000, B, 2))
        gp2.load_srk(gp.save_srk())
        again = srk.Patch(srk.AudioConfig(48000, B, 2))
        again.load_srk(gp2.save_srk())
        assert [m.get_id() for m in again.modules] == [m.get_id() for m in gp.modules]
        again.plan()
        assert (again.render(V, N, stems=True, mix=False)[0].view(np.uint32) == g.view(np.uint32)).all()
        if make is subtractive_file:  # the same file saved from scratch sounds different: the state matters
            cold = srk.Patch(srk.AudioConfig(48000, B, 2))
            cold.load_srk(sf.dumps(make(B=B, state=False)))
            cold.plan()
            c0 = cold.render(V, 2 * B, stems=True, mix=False)[0]
            assert not (c0 == g[:, :2 * B]).all(), "the file's state was not applied"
            assert_parity(c0, orc_render_same_order(orc, sf, cold, B, V, 2 * B), what="subtractive_file from new() state")


def orc_render_same_order(orc, sf, patch, B, V, N):
    """The oracle on the file `patch` would load from: save twice so the module order survives the reversal."""
    tmp = type(patch)(patch.audio_config)
    tmp.load_srk(patch.save_srk())
    op = orc.OraclePatch(48000, B, 2)
    op.load_srk(tmp.save_srk())
    return op.render(V, N)[0]


def test_sample_reload_rewinds_and_keeps_the_detector(srk, orc, cuda_device):
    """WaveBox.new (sample.rs:66,212-216): a load rewinds every voice at the start of the next block and
    playback waits for the next gate edge; envelope and clock state carry on.  Also a WAV load through
    both decoders, an empty table, and reset()."""
    V, B = 40, 4096
    gp, op, gh, oh = build_both(srk, orc, srk.patches.sampler, V, buffer_size=B)
    gp.plan()
    wave2 = (np.linspace(-1, 1, 997) ** 3).astype(np.float32)
    import io, wave as wavmod
    buf = io.BytesIO()
    with wavmod.open(buf, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(32000)
        w.writeframes((np.stack([wave2, -wave2], 1) * 32767).astype("<i2").tobytes())
    parts_g, parts_o = [], []
    for k in range(4):
        parts_g.append(gp.render(V, 2 * B, stems=True)[0])
        parts_o.append(op.render(V, 2 * B)[0])
        if k == 0:
            gp.set_sample(gh["sample"], wave2, 96000.0); op.set_sample(oh["sample"], wave2, 96000.0)
        elif k == 1:
            gp.load_wav(gh["sample"], buf.getvalue()); op.load_wav(oh["sample"], buf.getvalue())
        elif k == 2:
            gp.set_sample(gh["sample"], np.zeros(0, np.float32), 48000.0)
            op.set_sample(oh["sample"], np.zeros(0, np.float32), 48000.0)
    g, o = np.concatenate(parts_g, axis=1), np.concatenate(parts_o, axis=1)
    assert np.abs(o[1, 2 * B:6 * B]).max() > 0.3 and (o[1, 6 * B:] == 0).all()
    assert_parity(g, o, exact=True, what="sampler across reloads")
    gp.reset(); op.reset()
    gp.set_sample(gh["sample"], wave2, 8000.0); op.set_sample(oh["sample"], wave2, 8000.0)
    assert_parity(gp.render(V, 3 * B, stems=True)[0], op.render(V, 3 * B)[0], exact=True, what="sampler after reset")


@pytest.mark.parametrize("name,B", [("cfg2", 1024), ("cfg4", 1024), ("cfg3b", 256), ("cfg5_two_osc", 1024), ("sequenced", 1024),
                                    ("sampler", 1024)])
def test_schedule_invariance(srk, orc, cuda_device, monkeypatch, name, B):
    """The same patch rendered as one warp per voice group in plan order (SRK_WARPS=1) and as a
    software pipeline over 4/16 warps, with different chunk sizes, gives the same bits: the
    schedule only moves work between warps, never changes a voice's arithmetic."""
    V, N = 77, 5003
    builder = getattr(srk.patches, name)
    results = []
    # (fused, warps, chunk | samples per straight-line group): the interpreter's schedules, then the fused kernel
    # (fused, warps | stages, chunk | samples per straight-line group): the interpreter's schedules, then the fused kernel
    for fused, warps, step in [(0, 1, 8), (0, 1, 32), (0, 1, 1), (0, 4, 16), (0, 16, 32), (0, 16, 8), (0, 2, 32),
                               (1, 1, 4), (1, 1, 1), (1, 1, 2), (1, 1, 8), (1, 2, 4), (1, 3, 4), (1, 5, 8), (1, 8, 1)]:
        monkeypatch.setenv("SRK_FUSED", str(fused))
        if fused:
            monkeypatch.delenv("SRK_WARPS", raising=False)
            monkeypatch.delenv("SRK_STEP", raising=False)
            monkeypatch.setenv("SRK_FUSED_GROUP", str(step))
            monkeypatch.setenv("SRK_FUSED_STAGES", str(warps))
        else:
            monkeypatch.setenv("SRK_WARPS", str(warps))
            monkeypatch.setenv("SRK_STEP", str(step))
        p = srk.Patch(srk.AudioConfig(48000, B, 2))
        builder(p, V)
        p.plan()
        info = p.program_info(V)
        assert info["fused"] == fused
        if fused:
            assert info["fused_group"] == (step if B >= step else 1) and 1 <= info["n_warps"] <= warps
        else:
            assert info["n_warps"] <= max(warps, 1) and info["step_samples"] <= step
        assert (info["n_warps"] > 1) == (info["n_stages"] > 1)
        st, mx = p.render(V, N, stems=True, mix=True)
        # a second call continues from the persisted state under the same schedule
        st2, _ = p.render(V, 997, stems=True, mix=True)
        results.append((info, np.concatenate([st, st2], axis=1), mx))
    for k in ("SRK_WARPS", "SRK_STEP", "SRK_FUSED_GROUP", "SRK_FUSED_STAGES"):
        monkeypatch.delenv(k, raising=False)
    assert any(r[0]["n_warps"] > 1 for r in results) and any(r[0]["n_warps"] == 1 for r in results)
    assert any(r[0]["fused"] and r[0]["n_warps"] > 1 for r in results) and any(r[0]["fused"] and r[0]["n_warps"] == 1 for r in results)
    for info, st, mx in results[1:]:
        assert (st.view(np.uint32) == results[0][1].view(np.uint32)).all(), info
        assert (mx.view(np.uint32) == results[0][2].view(np.uint32)).all(), info  # one summation order for every schedule
    gp, op, _, _ = build_both(srk, orc, builder, V, buffer_size=B)
    o_st, _ = op.render(V, N + 997)
    assert_parity(results[0][1], o_st, what=f"{name} schedule 0")


@pytest.mark.parametrize("name,B", [("cfg4", 1024), ("cfg3b", 64), ("sampler", 1024)])
def test_one_warp_schedule_groups_per_block(srk, orc, cuda_device, monkeypatch, name, B):
    """The one-warp schedule puts several voice groups into one thread block (each warp its own tables,
    one barrier per chunk, optionally one per instruction).  Grouping must not change a bit: 333 voices
    = 10 full groups + 13 voices, so the last block has a ragged group and, for most G, spare warps."""
    V, N = 333, 3001
    builder = getattr(srk.patches, name)
    monkeypatch.setenv("SRK_WARPS", "1")
    monkeypatch.setenv("SRK_FUSED", "0")  # the interpreter's one-warp schedule is what this test is about
    results = []
    for groups, op_barrier in [(1, 0), (3, 0), (4, 1), (16, 0), (16, 1), (5, 0)]:
        monkeypatch.setenv("SRK_SOLO_GROUPS", str(groups))
        monkeypatch.setenv("SRK_SOLO_OP_BARRIER", str(op_barrier))
        p = srk.Patch(srk.AudioConfig(48000, B, 2))
        builder(p, V)
        p.plan()
        st, mx = p.render(V, N, stems=True, mix=True)
        st2, mx2 = p.render(V, 515, stems=True, mix=True)  # state written back by every warp of every block
        results.append((np.concatenate([st, st2], axis=1), np.concatenate([mx, mx2], axis=1)))
    for st, mx in results[1:]:
        assert (st.view(np.uint32) == results[0][0].view(np.uint32)).all()
        # (the chunk length follows the groups per block, and the order in which a group's 32 voices are summed
        # follows the chunk length below 32 samples: the mix agrees to rounding, not to the bit)
        assert np.abs(mx - results[0][1]).max() <= 1e-5 * np.sqrt(V)
    for k in ("SRK_WARPS", "SRK_SOLO_GROUPS", "SRK_SOLO_OP_BARRIER"):
        monkeypatch.delenv(k)
    gp, op, _, _ = build_both(srk, orc, builder, V, buffer_size=B)
    o_st, o_mix = op.render(V, N + 515)
    assert_parity(results[0][0], o_st, exact=(name == "sampler"), what=f"{name} one-warp groups")
    assert_mix_parity(results[0][1], o_mix, V)


def test_baseline_size_cfg4_sampled_parity(srk, orc, cuda_device):
    """BASELINE configs[3] per GPU (32768 of the 262144 voices x 48000 samples, the one-warp
    schedule): mix-only render at full size, then the stems of two voice slices (a middle one and the
    ragged tail of the shard) rendered with voice_offset and compared with the oracle -- the per-voice
    noise key is the global voice index, so the slices must reproduce the shard's voices exactly."""
    import torch

    V_total, V, N = 262144, 32768, 48000
    off = srk.shard.voice_range(V_total, 5, 8)[0]  # rank 5's shard
    p = srk.Patch(device=0)
    srk.patches.cfg4(p, V_total)
    p.plan()
    assert p.program_info(V)["n_warps"] == 1
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    p.render_into(V, N, off, None, mix.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    assert torch.isfinite(mix).all() and float(mix.abs().max()) > 10.0
    total = np.zeros((2, N))
    for first, cnt in ((off + 12345, 40), (off + V - 37, 37)):
        q = srk.Patch(device=0)
        srk.patches.cfg4(q, V_total)
        q.plan()
        st, _ = q.render(cnt, N, voice_offset=first, stems=True)
        op = orc.OraclePatch()
        srk.patches.cfg4(op, V_total)
        ref, _ = op.render(cnt, N, voice_offset=first)
        assert_parity(st, ref, exact=True, what=f"cfg4 voices {first}..")
        total += st.astype(np.float64).sum(axis=2)
    assert np.abs(total).max() > 0.1


def test_baseline_size_cfg3b_feedback(srk, orc, cuda_device):
    """BASELINE configs[2] with the in-graph feedback wire (one cut edge, buffer_size 1024 of delay),
    65536 voices x 48000: mix-only at full size (rings: 268 MB of HBM), a voice slice against the oracle."""
    import torch

    V, N = 65536, 48000
    p = srk.Patch(device=0)
    srk.patches.cfg3b(p, V)
    p.plan()
    assert p.program_info(V)["n_rings"] == 1
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    p.render_into(V, N, 0, None, mix.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    assert torch.isfinite(mix).all() and float(mix.abs().max()) > 10.0
    q = srk.Patch(device=0)
    srk.patches.cfg3b(q, V)
    q.plan()
    st, _ = q.render(33, N, voice_offset=40000, stems=True)
    op = orc.OraclePatch()
    srk.patches.cfg3b(op, V)
    ref, _ = op.render(33, N, voice_offset=40000)
    assert_parity(st, ref, exact=True, what="cfg3b voice slice")


@pytest.mark.parametrize("name,B", [("cfg2", 1024), ("cfg3b", 256), ("cfg4", 1024), ("sequenced", 1024), ("sampler", 1024)])
def test_state_export_import_resumes_bit_for_bit(srk, orc, cuda_device, name, B):
    """Checkpoint / resume (the reference serializes every module's DSP state, ui.rs:98-134):
    render 20000 samples, export the per-voice state, import it into a NEW patch built from the same description,
    render 28000 more == one 48000-sample render, bit for bit, stems and mix -- incl. cfg3b's feedback ring, the
    noise counter (cfg4), step counters (sequenced) and play positions (sampler)."""
    builders = dict(srk.patches.CONFIGS)
    builder = builders[name][0] if name in builders else getattr(srk.patches, name)
    V, N1, N2 = 100, 20000, 28000

    def fresh():
        p = srk.Patch(srk.AudioConfig(48000, B, 2))
        builder(p, V)
        p.plan()
        return p

    whole = fresh()
    w_st, w_mix = whole.render(V, N1 + N2, stems=True)
    a = fresh()
    a_st, a_mix = a.render(V, N1, stems=True)
    blob = a.state_export()
    assert blob[:8] == b"SRKSTATE"
    b = fresh()
    b.state_import(blob)
    b_st, b_mix = b.render(V, N2, stems=True)
    got = np.concatenate([a_st, b_st], axis=1)
    assert (got.view(np.uint32) == w_st.view(np.uint32)).all()
    assert (np.concatenate([a_mix, b_mix], axis=1).view(np.uint32) == w_mix.view(np.uint32)).all()
    # the exporting patch carries on unchanged, and an export straight after an import returns the same blob
    a_st2, _ = a.render(V, N2, stems=True)
    assert (a_st2.view(np.uint32) == b_st.view(np.uint32)).all()
    c = fresh()
    c.state_import(blob)
    assert c.state_export() == blob
    # a blob from another graph, or a truncated one, is refused
    other = srk.Patch(srk.AudioConfig(48000, B, 2))
    srk.patches.cfg1(other, V)
    other.plan()
    with pytest.raises(srk.SrackError):
        other.state_import(blob)
    with pytest.raises(srk.SrackError):
        fresh().state_import(blob[:-4])


def test_state_export_needs_a_render_and_shards_keep_their_offset(srk, orc, cuda_device):
    p = srk.Patch()
    srk.patches.cfg4(p, 64)
    p.plan()
    with pytest.raises(srk.SrackError):
        p.state_export()
    # a shard's blob resumes the shard (voice_offset is part of it: the noise key and the per-voice parameters follow)
    st1, _ = p.render(24, 3000, voice_offset=40, stems=True)
    blob = p.state_export()
    q = srk.Patch()
    srk.patches.cfg4(q, 64)
    q.plan()
    q.state_import(blob)
    a, _ = p.render(24, 2000, voice_offset=40, stems=True)
    b, _ = q.render(24, 2000, voice_offset=40, stems=True)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


def test_state_epoch_counts_every_reset_including_the_implicit_ones(srk, cuda_device):
    p = srk.Patch()
    srk.patches.cfg2(p, 64)
    p.plan()
    assert p.state_epoch() == 0
    p.render(64, 100)
    p.render(64, 100)
    assert p.state_epoch() == 1          # streaming the same range: the state carries over
    p.render(32, 100, voice_offset=32)   # another voice range: implicit reset (documented in srack_b200.h)
    assert p.state_epoch() == 2
    p.reset()
    assert p.state_epoch() == 3
    blob = p.state_export()
    p.state_import(blob)
    assert p.state_epoch() == 4


def test_no_voices_gives_a_silent_mix(srk, cuda_device):
    """A rank that got no voices (world size > voices) must contribute zeros to the NCCL sum."""
    p = srk.Patch()
    srk.patches.cfg2(p, 8)
    p.plan()
    mix = np.full((2, 512), 7.0, np.float32)
    p.render_into(0, 512, 0, None, mix.ctypes.data)
    assert not mix.any()


def test_co_resident_hint_changes_the_launch_shape_not_the_bits(srk, orc, cuda_device, schedule):
    """Several patches rendering at once on one device are scheduled for the sum of their voices
    (srk_set_co_resident_voices): another launch shape, the same samples."""
    V, N = 256, 6000
    a = srk.Patch()
    srk.patches.cfg4(a, V)
    a.plan()
    b = srk.Patch()
    srk.patches.cfg4(b, V)
    b.set_co_resident_voices(65536)
    b.plan()
    ia, ib = a.program_info(V), b.program_info(V)
    if schedule == "auto":
        assert ia["n_warps"] > 1 and ib["n_warps"] == 1  # staged when alone on the chip, one warp per group when it is full
    a_st, a_mix = a.render(V, N, stems=True)
    b_st, b_mix = b.render(V, N, stems=True)
    assert (a_st.view(np.uint32) == b_st.view(np.uint32)).all()


def test_cv_driven_oscillators_and_non_linear_are_bit_exact(srk, orc, cuda_device):
    """glibc's f64 exp2 (the V/oct conversion of a CV-driven oscillator, oscillator.rs:43-48) and powf (Non-Linear,
    math.rs:203-205) restated operation by operation on the device: a saw-modulates-saw FM patch with hard sync and a
    waveshaper -- no sine tap anywhere -- equals the oracle bit for bit, feedback ring included."""
    P = srk.PARAM

    def build(b, n_voices, seed=0):
        mod, car, depth, fb, shaper, out = (b.module_create(k) for k in ("OSCILLATOR", "OSCILLATOR", "MULTIPLY", "MULTIPLY", "NON_LINEAR", "OUTPUT"))
        b.set_param_per_voice(mod, P["OSC_VAL"], srk.patches._u(7, 1, n_voices, -3.0, 0.5))
        b.set_param_per_voice(car, P["OSC_VAL"], srk.patches._u(7, 2, n_voices, -2.0, 1.0))
        b.set_param_per_voice(depth, P["MATH_CONSTANT"], srk.patches._u(7, 3, n_voices, 0.0, 1.5))
        b.set_param(fb, P["MATH_CONSTANT"], 0.35)
        b.set_param_per_voice(shaper, P["MATH_CONSTANT"], srk.patches._u(7, 4, n_voices, 0.5, 2.0))
        b.connect(depth, 0, mod, 2)      # saw
        b.connect(car, 0, depth, 0)      # -> carrier CV: exp2 per sample
        b.connect(car, 1, mod, 1)        # square -> hard sync
        b.connect(fb, 0, car, 1)         # carrier square -> modulator CV: a cycle, cut one block late
        b.connect(mod, 0, fb, 0)
        b.connect(shaper, 0, car, 2)     # |saw|^c with the sign kept: powf per sample
        b.connect(out, 0, car, 2)
        b.connect(out, 1, shaper, 0)
        return {}

    gp, op, g, g_mix, o, o_mix = render_pair(srk, orc, build, 96, 20000, B=256)
    assert len(gp.plan_cuts()) == 1
    assert np.abs(o[0]).max() > 0.5 and np.abs(o[1]).max() > 0.5 and np.isfinite(o).all()
    assert_parity(g, o, exact=True, what="CV-driven saw + Non-Linear")


@pytest.mark.parametrize("name,B,V", [("cfg2", 1024, 200), ("cfg4", 1024, 96), ("cfg3b", 256, 64), ("sampler", 1024, 70), ("cfg2", 1024, 20000)])
def test_measured_schedule_choice_keeps_the_bits(srk, orc, cuda_device, monkeypatch, schedule, name, B, V):
    """The first long render of a schedule measures the alternative launch shapes and keeps the fastest
    (engine.cu tune_schedule).  Whatever it keeps -- every candidate is forced in turn here -- the samples are the ones
    the cost model's choice gives, the voice state carries over into the next call, and srk_get_program_info /
    srk_kernel_id describe the shape in use."""
    if schedule != "auto":
        pytest.skip("a forced schedule knob turns the measurement off")
    builders = dict(srk.patches.CONFIGS)
    builder = builders[name][0] if name in builders else getattr(srk.patches, name)
    N1, N2 = 16384, 3000

    def render(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        p = srk.Patch(srk.AudioConfig(48000, B, 2))
        builder(p, V)
        p.plan()
        z = p.render(V, 333, stems=True)  # too short to measure on: the choice is made MID-STREAM, on the state this left
        a = p.render(V, N1, stems=True)
        b = p.render(V, N2, stems=True)   # (sampler: the table epoch moved after the first call: the kept shape is rebuilt)
        assert p.state_epoch() == 1       # no candidate run touched the voice state
        out = (np.concatenate([z[0], a[0], b[0]], axis=1), np.concatenate([z[1], a[1], b[1]], axis=1), p.schedule_report(), p.kernel_id(V),
               p.program_info(V))
        for k in env:
            monkeypatch.delenv(k)
        return out

    base_st, base_mix, rep0, _, _ = render({"SRK_TUNE": "0"})
    assert rep0 == ""
    ids = set()
    for pick in (0, 1, 2, 3, 4, 5, 7, 10, 13):  # (with a second NVRTC version at hand the list is twice as long)
        st, mix, rep, kid, info = render({"SRK_TUNE_PICK": str(pick)})
        assert "picked" in rep or "no alternative" in rep
        assert (st.view(np.uint32) == base_st.view(np.uint32)).all(), (pick, rep)
        assert (mix.view(np.uint32) == base_mix.view(np.uint32)).all(), (pick, rep)
        assert (kid.startswith("fused:") and info["fused"] == 1) or (kid.startswith("interpreter:") and info["fused"] == 0)
        if "picked" in rep:
            assert kid.split(":")[1][:12] in rep or kid.startswith("interpreter:")
        ids.add(kid)
    if V <= 200:
        assert len(ids) >= 3  # several launch shapes were really exercised
    # and the real thing: measure (no cached decision), then the same decision from the cache
    st, mix, rep, kid, _ = render({"SRK_KERNEL_CACHE_OFF": "0", "SRK_KERNEL_CACHE": str(os.path.join(GOLDEN, "..", "..", "gpurun_out", "tune_test_cache_%s_%d" % (name, V)))})
    assert (st.view(np.uint32) == base_st.view(np.uint32)).all() and ("measured" in rep or "decision from" in rep)
    st2, _, rep2, kid2, _ = render({"SRK_KERNEL_CACHE": str(os.path.join(GOLDEN, "..", "..", "gpurun_out", "tune_test_cache_%s_%d" % (name, V)))})
    assert "decision from" in rep2 and kid2 == kid and (st2.view(np.uint32) == base_st.view(np.uint32)).all()

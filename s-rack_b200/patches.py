"""The BASELINE.json benchmark / parity patches, described once and
applied to any backend that offers the five verbs
    module_create(kind) -> handle, connect(sink, in_idx, src, src_port),
    set_param(handle, pid, value), set_param_per_voice(handle, pid, values), set_seed(seed)
(the product's `Patch` and the test oracle's `OraclePatch` both do).  Only reference
modules are used; per-voice parameters come from a stateless generator keyed by
(seed, stream, global voice index) so every backend and every GPU rank derives the
same values for the same voice.
"""
import math

import numpy as np

try:
    from .constants import PARAM as P
    from .constants import SEQ_NONE, grid_cell
except ImportError:  # loaded as a plain file next to constants.py (bench.py's CPU reference arm: no product library)
    from constants import PARAM as P
    from constants import SEQ_NONE, grid_cell

SEED = 0x5EED5EED
SAMPLE_RATE = 48000
CHANNELS = 2
BUFFER_SIZE = 1024

# port numbers (oscillator.rs:90-97, filter.rs:166-173)
SINE, SQUARE, SAW = 0, 1, 2
LOWPASS, BANDPASS, HIGHPASS = 0, 1, 2


def uniform01(seed, stream, n_voices, first_voice=0):
    """splitmix64 of (seed, stream, voice) -> float64 in [0, 1); one value per voice."""
    v = np.arange(first_voice, first_voice + n_voices, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (v + np.uint64(1))
             + np.uint64(0xD1B54A32D192ED03) * np.uint64(stream + 1))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _u(seed, stream, n, lo, hi):
    return (lo + (hi - lo) * uniform01(seed, stream, n)).astype(np.float32)


def hz_to_val(hz):
    """Oscillator `val` (V/oct above 440 Hz, oscillator.rs:46) for a frequency."""
    return np.float32(math.log2(hz / 440.0))


def cfg1(b, n_voices, seed=SEED):
    """single sine Oscillator -> Output (both channels)."""
    osc = b.module_create("OSCILLATOR")
    out = b.module_create("OUTPUT")
    b.connect(out, 0, osc, SINE)
    b.connect(out, 1, osc, SINE)
    return dict(osc=osc, out=out)


def _subtractive_core(b, n_voices, seed, tap=LOWPASS, two_osc=False):
    lfo = b.module_create("OSCILLATOR")
    adsr = b.module_create("ADSR")
    osc = b.module_create("OSCILLATOR")
    h = dict(lfo=lfo, adsr=adsr, osc=osc)
    if two_osc:
        h["osc2"] = b.module_create("OSCILLATOR")
        h["mix"] = b.module_create("MONO_MIXER")
    filt = b.module_create("MOOG_FILTER")
    vca = b.module_create("VCA")
    out = b.module_create("OUTPUT")
    h.update(filt=filt, vca=vca, out=out)
    b.set_param(lfo, P["OSC_VAL"], hz_to_val(2.0))
    b.connect(adsr, 0, lfo, SQUARE)
    for pid, v in (("ADSR_A_SEC", 0.01), ("ADSR_D_SEC", 0.1), ("ADSR_S_VAL", 0.5), ("ADSR_R_SEC", 0.2)):
        b.set_param(adsr, P[pid], v)
    detune = _u(seed, 0, n_voices, -50.0, 50.0) / np.float32(1200.0)
    b.set_param_per_voice(osc, P["OSC_VAL"], (np.float32(-1.0) + detune).astype(np.float32))
    if two_osc:
        spread = _u(seed, 1, n_voices, -15.0, 15.0) / np.float32(1200.0)
        b.set_param_per_voice(h["osc2"], P["OSC_VAL"], (np.float32(-1.0) + detune + spread).astype(np.float32))
        b.connect(h["mix"], 0, osc, SAW)
        b.connect(h["mix"], 1, h["osc2"], SAW)
        b.set_param(h["mix"], P["MIXER_GAIN0"], 0.5)
        b.set_param(h["mix"], P["MIXER_GAIN1"], 0.5)
        b.connect(filt, 0, h["mix"], 0)
    else:
        b.connect(filt, 0, osc, SAW)
    b.connect(filt, 1, adsr, 0)
    b.connect(vca, 0, filt, tap)
    b.connect(vca, 1, adsr, 0)
    b.connect(out, 0, vca, 0)
    b.connect(out, 1, vca, 0)
    return h


def cfg2(b, n_voices, seed=SEED):
    """saw Oscillator -> Moog Filter -> VCA, ADSR (gated by a 2 Hz square LFO) on filter CV and VCA CV;
    per-voice detune of +-50 cents around A3."""
    return _subtractive_core(b, n_voices, seed)


def cfg3(b, n_voices, seed=SEED, feedback=None):
    """2-osc FM: osc1.sine -> Multiply(index) -> osc2.cv; osc2.sine -> Output.
    `feedback`: also osc2.sine -> Multiply(fb) -> osc1.cv (a cycle the planner must cut)."""
    osc1 = b.module_create("OSCILLATOR")
    mul = b.module_create("MULTIPLY")
    osc2 = b.module_create("OSCILLATOR")
    h = dict(osc1=osc1, mul=mul, osc2=osc2)
    if feedback is not None:
        h["mul2"] = b.module_create("MULTIPLY")
    out = b.module_create("OUTPUT")
    h["out"] = out
    v1 = _u(seed, 0, n_voices, -2.0, 1.0)
    ratio = np.array([0.0, 7.0 / 12.0, 1.0], dtype=np.float32)[np.arange(n_voices) % 3]
    b.set_param_per_voice(osc1, P["OSC_VAL"], v1)
    b.set_param_per_voice(osc2, P["OSC_VAL"], (v1 + ratio).astype(np.float32))
    b.set_param_per_voice(mul, P["MATH_CONSTANT"], _u(seed, 1, n_voices, 0.0, 1.0))
    b.connect(mul, 0, osc1, SINE)
    b.connect(osc2, 0, mul, 0)
    b.connect(out, 0, osc2, SINE)
    b.connect(out, 1, osc2, SINE)
    if feedback is not None:
        b.set_param(h["mul2"], P["MATH_CONSTANT"], feedback)
        b.connect(h["mul2"], 0, osc2, SINE)
        b.connect(osc1, 0, h["mul2"], 0)
    return h


def cfg3b(b, n_voices, seed=SEED):
    return cfg3(b, n_voices, seed, feedback=0.25)


def cfg4(b, n_voices, seed=SEED, noise=True):
    """full subtractive: 2 osc + noise -> mixer -> filter -> VCA; 2 Hz gate -> 2 ADSRs; 5 Hz LFO +
    ADSR2 -> filter CV; ADSR1 -> VCA CV."""
    osc1 = b.module_create("OSCILLATOR")
    osc2 = b.module_create("OSCILLATOR")
    nz = b.module_create("NOISE") if noise else None
    mix = b.module_create("MONO_MIXER")
    gate = b.module_create("OSCILLATOR")
    adsr1 = b.module_create("ADSR")
    adsr2 = b.module_create("ADSR")
    lfo = b.module_create("OSCILLATOR")
    mull = b.module_create("MULTIPLY")
    add = b.module_create("ADD")
    filt = b.module_create("MOOG_FILTER")
    vca = b.module_create("VCA")
    out = b.module_create("OUTPUT")
    b.set_seed(seed)
    v1 = (np.float32(-1.0) + _u(seed, 0, n_voices, -1.0, 1.0)).astype(np.float32)
    b.set_param_per_voice(osc1, P["OSC_VAL"], v1)
    b.set_param_per_voice(osc2, P["OSC_VAL"], (v1 + _u(seed, 1, n_voices, -10.0, 10.0) / np.float32(1200.0)).astype(np.float32))
    b.set_param_per_voice(filt, P["MOOG_FREQ"], _u(seed, 2, n_voices, 0.1, 0.6))
    b.set_param_per_voice(filt, P["MOOG_RES"], _u(seed, 3, n_voices, 0.0, 0.9))
    b.set_param(gate, P["OSC_VAL"], hz_to_val(2.0))
    b.set_param(lfo, P["OSC_VAL"], hz_to_val(5.0))
    b.set_param(mull, P["MATH_CONSTANT"], 0.1)
    for a, vals in ((adsr1, (0.01, 0.1, 0.5, 0.2)), (adsr2, (0.05, 0.2, 0.3, 0.3))):
        for pid, v in zip(("ADSR_A_SEC", "ADSR_D_SEC", "ADSR_S_VAL", "ADSR_R_SEC"), vals):
            b.set_param(a, P[pid], v)
    b.connect(mix, 0, osc1, SAW)
    b.connect(mix, 1, osc2, SQUARE)
    if noise:
        b.connect(mix, 2, nz, 0)
        b.set_param(mix, P["MIXER_GAIN2"], 0.25)
    b.connect(filt, 0, mix, 0)
    b.connect(adsr1, 0, gate, SQUARE)
    b.connect(adsr2, 0, gate, SQUARE)
    b.connect(mull, 0, lfo, SINE)
    b.connect(add, 0, adsr2, 0)
    b.connect(add, 1, mull, 0)
    b.connect(filt, 1, add, 0)
    b.connect(vca, 0, filt, LOWPASS)
    b.connect(vca, 1, adsr1, 0)
    b.connect(out, 0, vca, 0)
    b.connect(out, 1, vca, 0)
    return dict(osc1=osc1, osc2=osc2, noise=nz, mix=mix, gate=gate, adsr1=adsr1, adsr2=adsr2, lfo=lfo, mull=mull,
                add=add, filt=filt, vca=vca, out=out)


def cfg5_bandpass(b, n_voices, seed=SEED):
    return _subtractive_core(b, n_voices, seed, tap=BANDPASS)


def cfg5_two_osc(b, n_voices, seed=SEED):
    return _subtractive_core(b, n_voices, seed, two_osc=True)


def cfg5_no_noise(b, n_voices, seed=SEED):
    return cfg4(b, n_voices, seed, noise=False)


def cfg5_gated_sine(b, n_voices, seed=SEED):
    """cfg1 through an ADSR-gated VCA."""
    lfo = b.module_create("OSCILLATOR")
    adsr = b.module_create("ADSR")
    osc = b.module_create("OSCILLATOR")
    vca = b.module_create("VCA")
    out = b.module_create("OUTPUT")
    b.set_param(lfo, P["OSC_VAL"], hz_to_val(3.0))
    b.set_param(adsr, P["ADSR_A_SEC"], 0.02)
    b.set_param_per_voice(osc, P["OSC_VAL"], _u(seed, 0, n_voices, -2.0, 2.0))
    b.connect(adsr, 0, lfo, SQUARE)
    b.connect(vca, 0, osc, SINE)
    b.connect(vca, 1, adsr, 0)
    b.connect(out, 0, vca, 0)
    b.connect(out, 1, vca, 0)
    return dict(lfo=lfo, adsr=adsr, osc=osc, vca=vca, out=out)


def sequenced(b, n_voices, seed=SEED):
    """the reference's own way of playing a patch.  A per-voice clock steps a Grid
    Sequencer (melody -> oscillator CV, gate -> ADSR) and a Pattern Sequencer (row 0 holds a gate for
    a noise burst through a second VCA, row 3 passes the clock through); the grid's sync output
    hard-syncs the oscillator at the top of the sequence."""
    clock = b.module_create("OSCILLATOR")
    grid = b.module_create("GRID_SEQUENCER")
    pat = b.module_create("PATTERN_SEQUENCER")
    osc = b.module_create("OSCILLATOR")
    adsr = b.module_create("ADSR")
    filt = b.module_create("MOOG_FILTER")
    vca = b.module_create("VCA")
    nz = b.module_create("NOISE")
    adsr2 = b.module_create("ADSR")
    vca2 = b.module_create("VCA")
    mix = b.module_create("MONO_MIXER")
    out = b.module_create("OUTPUT")
    b.set_seed(seed)
    tempo = 24.0 + 16.0 * uniform01(seed, 5, n_voices)  # steps per second
    b.set_param_per_voice(clock, P["OSC_VAL"], np.log2(tempo / 440.0).astype(np.float32))
    melody = [0, 3, 7, 12, SEQ_NONE, 10, 7, 3, 0, SEQ_NONE, 5, 8, 12, 15, 12, SEQ_NONE]
    cells = [SEQ_NONE if v == SEQ_NONE else grid_cell(v, hold=(i % 3 != 1)) for i, v in enumerate(melody)]
    b.set_sequence(grid, np.array(cells, dtype=np.int32))
    rows = np.full((8, 16), SEQ_NONE, dtype=np.int32)
    rows[0, ::4] = 1          # held gate on every fourth step
    rows[3, 1::2] = 0         # clock passed through on odd steps
    b.set_sequence(pat, rows)
    b.set_param(osc, P["OSC_VAL"], -1.0)
    for a, vals in ((adsr, (0.002, 0.02, 0.6, 0.01)), (adsr2, (0.0, 0.008, 0.0, 0.004))):
        for pid, v in zip(("ADSR_A_SEC", "ADSR_D_SEC", "ADSR_S_VAL", "ADSR_R_SEC"), vals):
            b.set_param(a, P[pid], v)
    b.connect(grid, 0, clock, SQUARE)
    b.connect(pat, 0, clock, SQUARE)
    b.connect(osc, 0, grid, 0)        # CV (V/oct)
    b.connect(osc, 1, grid, 2)        # sync at step 0
    b.connect(adsr, 0, grid, 1)       # gate
    b.connect(filt, 0, osc, SAW)
    b.connect(filt, 1, adsr, 0)
    b.connect(vca, 0, filt, LOWPASS)
    b.connect(vca, 1, adsr, 0)
    b.connect(adsr2, 0, pat, 0)
    b.connect(vca2, 0, nz, 0)
    b.connect(vca2, 1, adsr2, 0)
    b.connect(mix, 0, vca, 0)
    b.connect(mix, 1, vca2, 0)
    b.connect(mix, 2, pat, 3)
    b.set_param(mix, P["MIXER_GAIN2"], 0.05)
    b.connect(out, 0, mix, 0)
    b.connect(out, 1, pat, 8)          # the pattern sequencer's sync output
    return dict(clock=clock, grid=grid, pat=pat, osc=osc, adsr=adsr, filt=filt, vca=vca, noise=nz, adsr2=adsr2,
                vca2=vca2, mix=mix, out=out)


def sampler_wave(n=6000, rate=22050.0):
    """A deterministic test table for the Sample module: a decaying two-partial pluck, f32."""
    t = np.arange(n, dtype=np.float64) / rate
    x = np.exp(-6.0 * t) * (0.7 * np.sin(2 * np.pi * 220.0 * t) + 0.3 * np.sin(2 * np.pi * 663.0 * t + 0.5))
    return x.astype(np.float32), np.float32(rate)


def sampler(b, n_voices, seed=SEED, cv=True):
    """a Sample module (sample.rs) retriggered by a per-voice clock, its playback
    rate bent by a slow sine through the CV input (`ratio * 2^cv`, an index path: bit-exact or
    nothing), shaped by an ADSR-driven VCA; the raw player goes to channel 1."""
    clock = b.module_create("OSCILLATOR")
    lfo = b.module_create("OSCILLATOR")
    depth = b.module_create("MULTIPLY")
    smp = b.module_create("SAMPLE")
    adsr = b.module_create("ADSR")
    vca = b.module_create("VCA")
    out = b.module_create("OUTPUT")
    b.set_seed(seed)
    rate = 3.0 + 9.0 * uniform01(seed, 6, n_voices)  # retriggers per second
    b.set_param_per_voice(clock, P["OSC_VAL"], np.log2(rate / 440.0).astype(np.float32))
    b.set_param(lfo, P["OSC_VAL"], hz_to_val(1.3))
    b.set_param_per_voice(depth, P["MATH_CONSTANT"], _u(seed, 7, n_voices, 0.0, 1.5))
    wave, wave_rate = sampler_wave()
    b.set_sample(smp, wave, wave_rate)
    for pid, v in zip(("ADSR_A_SEC", "ADSR_D_SEC", "ADSR_S_VAL", "ADSR_R_SEC"), (0.001, 0.05, 0.4, 0.03)):
        b.set_param(adsr, P[pid], v)
    b.connect(depth, 0, lfo, SINE)
    b.connect(smp, 0, clock, SQUARE)
    if cv:
        b.connect(smp, 1, depth, 0)
    b.connect(adsr, 0, clock, SQUARE)
    b.connect(vca, 0, smp, 0)
    b.connect(vca, 1, adsr, 0)
    b.connect(out, 0, vca, 0)
    b.connect(out, 1, smp, 0)
    return dict(clock=clock, lfo=lfo, depth=depth, sample=smp, adsr=adsr, vca=vca, out=out)


# name -> (builder, BASELINE voice count, description)
CONFIGS = {
    "cfg1": (cfg1, 1, "single sine Oscillator->Output, 1 voice, 48000 samples"),
    "cfg2": (cfg2, 4096, "saw Oscillator->Filter->Envelope->VCA, 4096 detuned voices, 48 kHz x 1 s"),
    "cfg3": (cfg3, 65536, "2-osc FM feed-forward, 65536 voices, 48 kHz x 1 s"),
    "cfg3b": (cfg3b, 65536, "2-osc FM with in-graph feedback (one cut wire), 65536 voices, 48 kHz x 1 s"),
    "cfg4": (cfg4, 262144, "full subtractive (2 osc+noise+LFO+Filter+2xADSR+VCA), 262144 voices over 8 GPUs"),
}
# cfg5: 8 distinct graphs x 32768 voices, one graph per GPU
CFG5_GRAPHS = [cfg2, cfg3, cfg3b, cfg4, cfg5_bandpass, cfg5_two_osc, cfg5_no_noise, cfg5_gated_sine]
CFG5_VOICES = 32768

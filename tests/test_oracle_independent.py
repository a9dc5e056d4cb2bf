"""A second, independent restatement of the module kinds the reference has no test for, and the C++ oracle held to it
bit for bit.

Every function below is written straight from the reference source (file:line in its docstring) as a scalar,
per-sample loop over numpy f32 / Python f64 values -- not per block, no code shared with oracle/srack_oracle.cpp --
and is fed random inputs plus the corners: `a_sec = 0`, the filter's all-zero coefficient cache, NaN and +-inf
control voltages, resonance 1.0, unconnected ports.  The oracle's modules are driven one calc() at a time through
its debug hooks with the same input blocks; outputs must have equal bits (NaN compares equal to NaN).

libm: `sin`, `exp2`, `fmod` and `powf` are the platform's (glibc) in both, as in the reference (Rust std calls
the platform libm) -- what is pinned here is the algorithm around them.
"""
import ctypes
import ctypes.util
import math

import numpy as np
import pytest

from oracle import orc

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.exp2.restype = ctypes.c_double
_libm.exp2.argtypes = [ctypes.c_double]


def powf(a, b):
    return F(_libm.powf(float(a), float(b)))


def rmin(a, b):
    """f32::min: the other operand when one is NaN."""
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def rmax(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


class Detector:
    """TransitionDetector, synth.rs:276-298: `last` starts true."""

    def __init__(self):
        self.last = True

    def is_transition(self, val):
        above = bool(val > 0.0)
        t = above and not self.last
        self.last = above
        return t


# --------------------------------------------------------------------------- restatements
class MoogRef:
    """filter.rs:48-91 (InternalMoogFilterState) + :202-216 (the per-sample clamps and the output tuple)."""

    def __init__(self, freq=0.2, res=0.5, exp_amt=0.5):
        self.freq_p, self.res_p, self.exp_amt = F(freq), F(res), F(exp_amt)
        self.f = self.p = self.q = F(0)
        self.b = [F(0)] * 5
        self.freq = self.res = F(0)  # #[derive(Default)]: the cache starts at (0, 0) with all-zero coefficients

    def step(self, audio, cv):
        frequency = rmin(rmax(self.freq_p + cv * self.exp_amt, F(0.0)), F(0.9))  # :213
        res = rmin(rmax(self.res_p, F(0.0)), F(1.0))                              # :214
        if frequency != self.freq or res != self.res:                            # :61
            self.freq = frequency
            self.res = res
            self.q = F(1.0) - self.freq
            self.p = self.freq + F(0.8) * self.freq * self.q
            self.f = self.p * F(2.0) - F(1.0)
            self.q = self.res * (F(1.0) + F(0.5) * self.q * (F(1.0) - self.q + F(5.6) * self.q * self.q))
        b = self.b
        inp = audio - (self.q * b[4])
        t1 = b[1]
        b[1] = (inp + b[0]) * self.p - b[1] * self.f
        t2 = b[2]
        b[2] = (b[1] + t1) * self.p - b[2] * self.f
        t1 = b[3]
        b[3] = (b[2] + t2) * self.p - b[3] * self.f
        b[4] = (b[3] + t1) * self.p - b[4] * self.f
        b[4] = b[4] - (b[4] * b[4] * b[4]) * F(0.166667)  # powi(3)
        b[0] = inp
        for k in range(5):  # clamp_buffers :84-89
            b[k] = rmax(rmin(b[k], F(1.0)), F(-1.0))
        # calc() returns (b4, input - b4, 3 (b3 - b4)) and :211 assigns it to (lowpass, HIGHPASS, BANDPASS)
        return b[4], F(3.0) * (b[3] - b[4]), inp - b[4]  # in port order: lowpass, bandpass, highpass


class AdsrRef:
    """adsr.rs:134-217; new() :36-53."""

    def __init__(self, a=0.0, d=0.5, s=0.25, r=0.5, sample_rate=48000):
        self.a, self.d, self.s, self.r = F(a), F(d), F(s), F(r)
        self.phase, self.mode, self.r_val, self.from_a_val = F(0), "None", F(0), F(0)
        self.sr = F(sample_rate)
        self.det = Detector()

    def step(self, gate):
        """gate: f32, or None when the port is unconnected."""
        tr = self.det.is_transition(F(0.0) if gate is None else gate)
        high = gate is not None and bool(gate > 0.0)
        m = self.mode
        if m == "None":
            if high:
                self.phase, self.mode = F(0), "Attack"
        elif m == "Attack":
            self.phase = self.phase + F(1.0) / (self.sr * self.a)
            if self.phase >= 1.0:
                self.phase, self.mode = F(0), "Decay"
            elif tr:
                self.phase = F(0)
                self.r_val = self.from_a_val
        elif m == "Decay":
            self.phase = self.phase + F(1.0) / (self.sr * self.d)
            if self.phase >= 1.0:
                self.phase, self.mode = F(0), "Sustain"
            if tr:
                self.phase, self.mode = F(0), "Attack"
        elif m == "Sustain":
            if gate is None or bool(gate <= 0.0):  # (a NaN gate is neither > 0 nor <= 0: Sustain holds)
                self.phase, self.mode = F(0), "Release"
            if tr:
                self.phase, self.mode = F(0), "Attack"
        else:  # Release
            if high:
                self.phase, self.mode = F(0), "Attack"
            self.phase = self.phase + F(1.0) / (self.sr * self.r)  # added even after the retrigger (:195)
            if self.phase >= 1.0:
                self.phase, self.r_val, self.mode = F(0), F(0), "None"
        m = self.mode
        if m == "None":
            out = F(0)
        elif m == "Attack":
            out = self.r_val + (F(1.0) - self.r_val) * self.phase
        elif m == "Decay":
            out = self.s + (F(1.0) - self.s) * (F(1.0) - self.phase)
        elif m == "Sustain":
            out = self.s
        else:
            out = self.s * (F(1.0) - self.phase)
        if m != "Attack":
            self.r_val = out
        else:
            self.from_a_val = out
        return out


def vca_ref(audio, cv, negative, n):
    """vca.rs:117-148: zero unless both ports are connected."""
    out = np.zeros(n, F)
    if audio is None or cv is None:
        return out
    for i in range(n):
        out[i] = audio[i] * cv[i] if (negative or cv[i] > 0.0) else F(0.0)
    return out


def mixer_ref(inputs, gains):
    """mixer.rs:101-122: input by input, `*dst += src * gain`."""
    n = len(next(x for x in inputs if x is not None)) if any(x is not None for x in inputs) else 0
    out = np.zeros(n, F)
    for buf, g in zip(inputs, gains):
        if buf is None:
            continue
        for i in range(n):
            out[i] = out[i] + buf[i] * F(g)
    return out


def math_ref(op, i1, i2, constant, n):
    """math.rs:46-52,139-160 (Add / Subtract / Multiply) and :203-205,292-313 (Non-Linear)."""
    def f(a, b):
        if op == "ADD":
            return a + b
        if op == "SUBTRACT":
            return a - b
        if op == "MULTIPLY":
            return a * b
        return powf(a, b) if a > 0.0 else -powf(-a, b)
    out = np.zeros(n, F)
    c = F(constant)
    for i in range(n):
        if i1 is not None and i2 is not None:
            out[i] = f(i1[i], i2[i])
        elif i1 is not None:
            out[i] = f(i1[i], c)
        elif i2 is not None:
            out[i] = f(F(0.0), i2[i])
        else:
            out[i] = f(F(0.0), c)
    return out


def poly_blep(t, dt):
    """oscillator.rs:50-67, f64."""
    if dt == 0.0:
        return 0.0
    if t < dt:
        t = t / dt
        return t + t - t * t - 1.0
    elif t > 1.0 - dt:
        t = (t - 1.0) / dt
        return t * t + t + t + 1.0
    return 0.0


class OscRef:
    """oscillator.rs:43-48 (V/oct) and :124-153 (per-sample body); `2f64.powf(x)` is libm exp2."""

    def __init__(self, val=0.0, antialiasing=True, sample_rate=48000):
        self.val, self.aa, self.sr = F(val), antialiasing, sample_rate
        self.pos = 0.0
        self.det = Detector()

    def step(self, cv, sync):
        if self.det.is_transition(F(0.0) if sync is None else sync):
            self.pos = 0.0
        x = float(self.val) if cv is None else float(cv) + float(self.val)
        delta = 440.0 * _libm.exp2(x) / float(self.sr)
        pos = self.pos
        sine = F(math.sin(pos * math.pi * 2.0)) if math.isfinite(pos) else F(np.nan)
        sq = F(-1.0) if pos < 0.5 else F(1.0)
        sq = sq - (F(poly_blep(pos, delta) - poly_blep(math.fmod(pos + 0.5, 1.0) if math.isfinite(pos) else math.nan, delta))
                   if self.aa else F(0.0))
        saw = (F(pos) * F(2.0) - F(1.0)) - (F(poly_blep(pos, delta)) if self.aa else F(0.0))
        pos = pos + delta
        self.pos = math.fmod(pos, 1.0) if math.isfinite(pos) else math.nan
        return sine, sq, saw


# --------------------------------------------------------------------------- driving the oracle one calc() at a time
B = 64


def _signal(rng, n, kind):
    if kind == "audio":
        x = rng.uniform(-1.2, 1.2, n)
    elif kind == "cv":
        x = rng.uniform(-1.5, 2.5, n)
    elif kind == "gate":  # runs of high / low with random lengths, some exactly 0, some negative
        x = np.zeros(n)
        i, level = 0, 0.0
        while i < n:
            k = int(rng.integers(1, 400))
            x[i:i + k] = level
            i += k
            level = float(rng.choice([0.0, 1.0, -1.0, 0.3, 1.0, 0.0]))
    else:
        raise ValueError(kind)
    return x.astype(F)


def _with_specials(rng, x, frac=0.01):
    x = x.copy()
    idx = rng.choice(len(x), max(1, int(len(x) * frac)), replace=False)
    x[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-42, 3e38, -3e38], dtype=F), len(idx))
    return x


class Rig:
    """One module under test; every connected input fed from a private Add module whose output buffer is overwritten."""

    def __init__(self, kind, n_in, connected, params=(), sample_rate=48000):
        self.p = orc.OraclePatch(sample_rate, B, 2)
        self.m = self.p.module_create(kind)
        self.src = []
        for i in range(n_in):
            if connected[i]:
                s = self.p.module_create("ADD")
                self.p.connect(self.m, i, s, 0)
                self.src.append(s)
            else:
                self.src.append(None)
        out = self.p.module_create("OUTPUT")
        self.p.connect(out, 0, self.m, 0)
        for pid, v in params:
            self.p.set_param(self.m, pid, v)
        self.p.plan()
        self.p.debug_prepare(1)
        self.n_out = orc.lib().orc_num_outputs(self.p._h, self.m)

    def run(self, inputs, n):
        assert n % B == 0
        outs = [np.zeros(n, F) for _ in range(self.n_out)]
        for k in range(0, n, B):
            for s, x in zip(self.src, inputs):
                if s is not None:
                    self.p.debug_set_output(s, 0, x[k:k + B])
            self.p.debug_calc(self.m)
            for port in range(self.n_out):
                outs[port][k:k + B] = self.p.debug_output(self.m, port)
        return outs


def assert_same_bits(got, want, what):
    got, want = np.asarray(got, F), np.asarray(want, F)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    if not same.all():
        i = int(np.argmin(same))
        raise AssertionError(f"{what}: first difference at sample {i}: oracle {got[i]!r} vs restatement {want[i]!r} "
                             f"({int((~same).sum())} of {same.size} differ)")


N = 64 * 60


@pytest.mark.parametrize("case", [
    dict(freq=0.2, res=0.5, exp=0.5, conn=(1, 1), special=False),
    dict(freq=0.05, res=1.0, exp=1.0, conn=(1, 1), special=False),     # resonance 1.0
    dict(freq=0.0, res=0.0, exp=0.5, conn=(1, 1), special=False, virgin=True),  # all-zero cache until cv first goes > 0
    dict(freq=0.0, res=0.0, exp=0.5, conn=(1, 0), special=False),     # never leaves the all-zero cache
    dict(freq=0.6, res=0.9, exp=0.5, conn=(0, 1), special=False),     # no audio
    dict(freq=0.3, res=1.7, exp=-0.7, conn=(1, 1), special=True),     # res clamped, negative exp_amt, NaN / inf input
    dict(freq=float("nan"), res=float("nan"), exp=0.5, conn=(1, 1), special=True),
])
def test_moog_filter(case):
    rng = np.random.default_rng(11)
    with np.errstate(all="ignore"):
        audio = _signal(rng, N, "audio")
        cv = _signal(rng, N, "cv") * F(0.5)
        if case.get("virgin"):
            cv[:700] = -np.abs(cv[:700])      # clamped to fc = 0: cache hit on the Default state, coefficients stay 0
            cv[1500:2100] = -np.abs(cv[1500:2100])  # back at (0, 0) after a miss: now really computed (f = -1)
        if case["special"]:
            audio, cv = _with_specials(rng, audio), _with_specials(rng, cv)
        rig = Rig("MOOG_FILTER", 2, case["conn"], [(0, case["freq"]), (1, case["res"]), (2, case["exp"])])
        got = rig.run([audio, cv], N)
        ref = MoogRef(case["freq"], case["res"], case["exp"])
        want = np.zeros((3, N), F)
        for i in range(N):
            a = audio[i] if case["conn"][0] else F(0)
            c = cv[i] if case["conn"][1] else F(0)
            want[:, i] = ref.step(a, c)
    for port, name in enumerate(("lowpass", "bandpass", "highpass")):
        assert_same_bits(got[port], want[port], f"moog {name} {case}")
    if case.get("virgin"):
        assert not got[0][:700].any() and got[0][1500:2100].any()


@pytest.mark.parametrize("case", [
    dict(a=0.01, d=0.1, s=0.5, r=0.2, conn=True),
    dict(a=0.0, d=0.5, s=0.25, r=0.5, conn=True),          # new() defaults: a_sec = 0 -> +inf increment
    dict(a=0.002, d=0.001, s=1.0, r=0.0005, conn=True),
    dict(a=0.004, d=0.003, s=0.0, r=0.0, conn=True),       # r_sec = 0
    dict(a=0.001, d=0.002, s=0.7, r=0.004, conn=True, special=True),
    dict(a=0.01, d=0.1, s=0.5, r=0.2, conn=False),         # unconnected gate: stays idle
    dict(a=0.003, d=0.002, s=0.6, r=0.01, conn=True, sr=44100),
])
def test_adsr(case):
    rng = np.random.default_rng(5)
    sr = case.get("sr", 48000)
    gate = _signal(rng, N, "gate")
    if case.get("special"):
        gate = _with_specials(rng, gate, 0.02)
    rig = Rig("ADSR", 1, (case["conn"],), [(0, case["a"]), (1, case["d"]), (2, case["s"]), (3, case["r"])], sample_rate=sr)
    got = rig.run([gate], N)[0]
    ref = AdsrRef(case["a"], case["d"], case["s"], case["r"], sr)
    with np.errstate(all="ignore"):
        want = np.array([ref.step(gate[i] if case["conn"] else None) for i in range(N)], F)
    assert_same_bits(got, want, f"adsr {case}")
    if case["conn"]:
        assert want.max() > 0.5


@pytest.mark.parametrize("negative", [False, True])
@pytest.mark.parametrize("conn", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_vca(negative, conn):
    rng = np.random.default_rng(3)
    audio = _with_specials(rng, _signal(rng, 512, "audio"))
    cv = _with_specials(rng, _signal(rng, 512, "cv"))
    rig = Rig("VCA", 2, conn, [(0, 1.0 if negative else 0.0)])
    got = rig.run([audio, cv], 512)[0]
    with np.errstate(all="ignore"):
        want = vca_ref(audio if conn[0] else None, cv if conn[1] else None, negative, 512)
    assert_same_bits(got, want, f"vca negative={negative} conn={conn}")


@pytest.mark.parametrize("conn", [(1, 1, 1, 1), (1, 0, 1, 0), (0, 0, 0, 1), (0, 0, 0, 0), (1, 1, 0, 0)])
def test_mono_mixer(conn):
    rng = np.random.default_rng(9)
    xs = [_with_specials(rng, _signal(rng, 512, "audio"), 0.004) for _ in range(4)]
    gains = [1.0, 0.25, -1.5, 3.0e-3]
    rig = Rig("MONO_MIXER", 4, conn, list(enumerate(gains)))
    got = rig.run(xs, 512)[0]
    with np.errstate(all="ignore"):
        want = mixer_ref([x if c else None for x, c in zip(xs, conn)], gains)
    if not any(conn):
        want = np.zeros(512, F)
    assert_same_bits(got, want, f"mixer conn={conn}")


@pytest.mark.parametrize("op", ["ADD", "SUBTRACT", "MULTIPLY", "NON_LINEAR"])
@pytest.mark.parametrize("conn", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_math_and_non_linear(op, conn):
    rng = np.random.default_rng(21)
    a = _with_specials(rng, _signal(rng, 512, "audio") * F(2))
    b = _with_specials(rng, _signal(rng, 512, "cv"))
    constant = 1.37
    rig = Rig(op, 2, conn, [(0, constant)])
    got = rig.run([a, b], 512)[0]
    with np.errstate(all="ignore"):
        want = math_ref(op, a if conn[0] else None, b if conn[1] else None, constant, 512)
    assert_same_bits(got, want, f"{op} conn={conn}")


@pytest.mark.parametrize("case", [
    dict(val=0.0, aa=True, conn=(0, 0)),
    dict(val=-1.03, aa=True, conn=(0, 0)),
    dict(val=2.9, aa=False, conn=(0, 0)),
    dict(val=-7.78135971352466, aa=True, conn=(0, 0)),     # the 2 Hz gate oscillator of cfg2
    dict(val=-0.4, aa=True, conn=(1, 0)),                  # CV-driven (FM)
    dict(val=0.3, aa=True, conn=(1, 1)),                   # CV + hard sync
    dict(val=1.1, aa=False, conn=(0, 1)),
    dict(val=5.5, aa=True, conn=(1, 0), hot=True),         # delta beyond 0.5: polyBLEP's overlapping arms
    dict(val=0.0, aa=True, conn=(1, 1), special=True),     # NaN / inf CV poisons the phase for good
])
def test_oscillator(case):
    rng = np.random.default_rng(17)
    n = 64 * 40
    cv = _signal(rng, n, "cv") * F(0.4 if not case.get("hot") else 1.0)
    sync = _signal(rng, n, "gate")
    if case.get("special"):
        cv = cv.copy()
        cv[1800] = np.nan
        cv[900] = np.inf
        cv[300] = -np.inf
        sync = _with_specials(rng, sync, 0.01)
    rig = Rig("OSCILLATOR", 2, case["conn"], [(0, case["val"]), (1, 1.0 if case["aa"] else 0.0)])
    got = rig.run([cv, sync], n)
    ref = OscRef(case["val"], case["aa"])
    want = np.zeros((3, n), F)
    with np.errstate(all="ignore"):
        for i in range(n):
            want[:, i] = ref.step(cv[i] if case["conn"][0] else None, sync[i] if case["conn"][1] else None)
    for port, name in enumerate(("sine", "square", "saw")):
        assert_same_bits(got[port], want[port], f"oscillator {name} {case}")


def test_output_copies_or_zeroes():
    """output.rs:46-60: channel c is a copy of its input block, zeros when unconnected; rendered through execute()."""
    p = orc.OraclePatch(48000, B, 3)
    o1 = p.module_create("OSCILLATOR")
    o2 = p.module_create("OSCILLATOR")
    out = p.module_create("OUTPUT")
    p.set_param(o2, 0, 0.5)
    p.connect(out, 0, o1, 2)
    p.connect(out, 2, o2, 1)  # channel 1 left unconnected
    p.plan()
    stems, _ = p.render(1, 3 * B)
    a, b = OscRef(0.0), OscRef(0.5)
    saw = np.array([a.step(None, None)[2] for _ in range(3 * B)], F)
    sq = np.array([b.step(None, None)[1] for _ in range(3 * B)], F)
    assert_same_bits(stems[0, :, 0], saw, "output channel 0")
    assert not stems[1].any()
    assert_same_bits(stems[2, :, 0], sq, "output channel 2")


def test_chain_through_execute():
    """cfg2's chain (saw -> Moog lowpass -> VCA, 2 Hz-style square gate -> ADSR -> both CVs) run by the oracle's
    block-based execute() over its planner, against the restatements composed per sample: pins the wiring and the
    one-block-late rule's absence on an acyclic patch."""
    p = orc.OraclePatch(48000, B, 2)
    lfo, osc, adsr, filt, vca, out = (p.module_create(k) for k in ("OSCILLATOR", "OSCILLATOR", "ADSR", "MOOG_FILTER", "VCA", "OUTPUT"))
    p.set_param(lfo, 0, math.log2(40.0 / 440.0))  # 40 Hz gate: several envelopes inside the render
    p.set_param(osc, 0, -0.97)
    for pid, v in enumerate((0.002, 0.004, 0.5, 0.003)):
        p.set_param(adsr, pid, v)
    p.connect(adsr, 0, lfo, 1)
    p.connect(filt, 0, osc, 2)
    p.connect(filt, 1, adsr, 0)
    p.connect(vca, 0, filt, 0)
    p.connect(vca, 1, adsr, 0)
    p.connect(out, 0, vca, 0)
    p.connect(out, 1, filt, 2)
    p.plan()
    n = 64 * 50
    stems, _ = p.render(1, n)
    r_lfo, r_osc = OscRef(F(math.log2(40.0 / 440.0))), OscRef(-0.97)
    r_adsr, r_f = AdsrRef(0.002, 0.004, 0.5, 0.003), MoogRef()
    want = np.zeros((2, n), F)
    with np.errstate(all="ignore"):
        for i in range(n):
            gate = r_lfo.step(None, None)[1]
            saw = r_osc.step(None, None)[2]
            env = r_adsr.step(gate)
            lp, _, hp = r_f.step(saw, env)
            want[0, i] = lp * env if env > 0.0 else F(0)
            want[1, i] = hp
    assert_same_bits(stems[0, :, 0], want[0], "chain: VCA out")
    assert_same_bits(stems[1, :, 0], want[1], "chain: filter highpass")
    assert np.abs(want[0]).max() > 0.05

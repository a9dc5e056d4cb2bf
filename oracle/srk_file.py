"""CPU ORACLE, TEST INFRASTRUCTURE ONLY: the reference's .srk patch file, restated from its serde
derives as data.

`FileFormat` (reference src/ui.rs:578-586) is written with
`container.serialize(&mut rmp_serde::Serializer::new(&mut buf))` (ui.rs:112).  The byte layout comes
from a third-party crate that is not under /root/reference: **rmp-serde 1.3.0** (Cargo.toml:31) over
**rmp 0.8**.  Its published rules, restated here: a struct is an array of its non-skipped fields in
declaration order; newtype structs, Arc, RwLock, Mutex and Box are transparent; Option is nil or the
value; an enum variant with data is a 1-entry map {variant name: data}; a unit variant is its name as
a str; tuples and fixed arrays are arrays; f32 -> 0xca, f64 -> 0xcb, bool -> 0xc2/0xc3; unsigned
integers take the shortest of fixint / 0xcc / 0xcd / 0xce / 0xcf; str -> fixstr / 0xd9 / 0xda / 0xdb;
arrays -> fixarray / 0xdc / 0xdd.  The reference ships no .srk file and has no test that saves or
loads one => **parity unpinned** by reference vectors; this module and the product's C++ codec
(s-rack_b200/csrc/srkfile.cpp) are two independent restatements checked against each other byte for
byte (tests/test_srk_file.py).

Schemas below mirror the struct definitions field by field (`#[serde(skip)]` fields left out):
src/synth/output.rs:7, oscillator.rs:10,309, sequencer.rs:13,337,628, adsr.rs:8,27, vca.rs:7,
filter.rs:12,49,252, mixer.rs:7, sample.rs:16,73, math.rs:7,14,177, freeverb.rs:8, synth.rs:277,300.
"""
import struct

# ---- type descriptors -------------------------------------------------------------------------
STR, F32, F64, BOOL, UINT = "str", "f32", "f64", "bool", "uint"


def Opt(t):
    return ("opt", t)


def Seq(t):
    return ("seq", t)


def Tup(*ts):
    return ("tup", ts)


def Struct(*fields):
    return ("struct", fields)


def UnitEnum(*names):
    return ("unit_enum", names)


BUF = Opt(Seq(F32))                                   # AudioBuffer(Option<Arc<RwLock<Box<[f32]>>>>), synth.rs:28
DET = Struct(("last", BOOL))                          # TransitionDetector, synth.rs:277
MOOG_STATE = Struct(("f", F32), ("p", F32), ("q", F32), ("b", Tup(F32, F32, F32, F32, F32)), ("freq", F32), ("res", F32))
WAVEBOX = Struct(("samples", Seq(F32)), ("sample_rate", F32), ("new", BOOL))

_GRID_TAIL = (("octaves", UINT), ("steps_per_octave", UINT), ("current_step", UINT), ("transition_detector", DET),
              ("sync_transition_detector", DET), ("last", F32), ("ui_dirty", BOOL))
VARIANTS = [  # SynthModuleType, synth.rs:300-317, declaration order
    ("OutputModuleV0", Struct(("id", STR), ("bufs", Seq(BUF)))),
    ("OscillatorModuleV0", Struct(("id", STR), ("val", F32), ("sample_rate", UINT), ("sine", BUF), ("square", BUF),
                                  ("saw", BUF), ("pos", F64), ("antialiasing", BOOL), ("sync_detector", DET))),
    ("NoiseModuleV0", Struct(("id", STR), ("out", BUF))),
    ("GridSequencerModuleV0", Struct(("id", STR), ("cv_out", BUF), ("gate_out", BUF), ("sync_out", BUF),
                                     ("sequence", Seq(Opt(UINT))), *_GRID_TAIL)),
    ("GridSequencerModuleV1", Struct(("id", STR), ("cv_out", BUF), ("gate_out", BUF), ("sync_out", BUF),
                                     ("sequence", Seq(Opt(Tup(UINT, BOOL)))), *_GRID_TAIL)),
    ("PatternSequencerModuleV0", Struct(("id", STR), ("gate_outs", Seq(BUF)), ("sync_out", BUF),
                                        ("sequence", Seq(Seq(Opt(BOOL)))), ("current_step", UINT),
                                        ("transition_detector", DET), ("sync_transition_detector", DET), ("ui_dirty", BOOL))),
    ("ADSRModuleV0", Struct(("id", STR), ("a_sec", F32), ("d_sec", F32), ("s_val", F32), ("r_sec", F32), ("phase", F32),
                            ("mode", UnitEnum("Attack", "Decay", "Sustain", "Release", "None")), ("r_val", F32),
                            ("from_a_val", F32), ("sample_rate", F32), ("transition_detector", DET),
                            ("output_buffer", BUF), ("ui_dirty", BOOL))),
    ("VCAModuleV0", Struct(("id", STR), ("buf", BUF), ("negative", BOOL))),
    ("MoogFilterModuleV0", Struct(("id", STR), ("buf", BUF), ("freq", F32), ("res", F32), ("exp_amt", F32),
                                  ("state", MOOG_STATE))),
    ("MoogFilterModuleV1", Struct(("id", STR), ("lowpass", BUF), ("bandpass", BUF), ("highpass", BUF), ("freq", F32),
                                  ("res", F32), ("exp_amt", F32), ("state", MOOG_STATE))),
    ("MonoMixerModuleV0", Struct(("id", STR), ("gain", Seq(F32)), ("buf", BUF))),
    ("SampleModuleV0", Struct(("id", STR), ("transition_detector", DET), ("pos", F32), ("buf", BUF), ("wavebox", WAVEBOX),
                              ("playing", BOOL), ("sample_rate", F32))),
    ("MathModuleV0", Struct(("id", STR), ("buf", BUF), ("constant", F32),
                            ("operation", UnitEnum("Add", "Subtract", "Multiply")))),
    ("NonLinearModuleV0", Struct(("id", STR), ("buf", BUF), ("constant", F32))),
    ("FreeverbModuleV0", Struct(("id", STR), ("left_out", BUF), ("right_out", BUF), ("sample_rate", UINT),
                                ("dampening", F64), ("dampening_ctl", F64), ("freeze", BOOL), ("freeze_ctl", BOOL),
                                ("wet", F64), ("wet_ctl", F64), ("width", F64), ("width_ctl", F64), ("room_size", F64),
                                ("room_size_ctl", F64), ("dry", F64), ("dry_ctl", F64))),
]
SCHEMA = dict(VARIANTS)
MODULE = ("enum", VARIANTS)
FILE_FORMAT = Struct(("modules", Seq(MODULE)), ("connections", Seq(Tup(STR, UINT, STR, UINT))),
                     ("positions", Seq(Tup(STR, Tup(F32, F32)))))


# ---- encoder ----------------------------------------------------------------------------------
def _uint(v):
    if v < 128:
        return bytes([v])
    if v < 1 << 8:
        return b"\xcc" + struct.pack(">B", v)
    if v < 1 << 16:
        return b"\xcd" + struct.pack(">H", v)
    if v < 1 << 32:
        return b"\xce" + struct.pack(">I", v)
    return b"\xcf" + struct.pack(">Q", v)


def _str(s):
    b = s.encode()
    n = len(b)
    head = bytes([0xa0 | n]) if n < 32 else b"\xd9" + bytes([n]) if n < 256 else \
        b"\xda" + struct.pack(">H", n) if n < 65536 else b"\xdb" + struct.pack(">I", n)
    return head + b


def _arr(n):
    return bytes([0x90 | n]) if n < 16 else b"\xdc" + struct.pack(">H", n) if n < 65536 else b"\xdd" + struct.pack(">I", n)


def encode(t, v):
    if t == STR:
        return _str(v)
    if t == F32:
        return b"\xca" + struct.pack(">f", v)
    if t == F64:
        return b"\xcb" + struct.pack(">d", v)
    if t == BOOL:
        return b"\xc3" if v else b"\xc2"
    if t == UINT:
        return _uint(int(v))
    kind = t[0]
    if kind == "opt":
        return b"\xc0" if v is None else encode(t[1], v)
    if kind == "seq":
        return _arr(len(v)) + b"".join(encode(t[1], x) for x in v)
    if kind == "tup":
        assert len(v) == len(t[1])
        return _arr(len(v)) + b"".join(encode(tt, x) for tt, x in zip(t[1], v))
    if kind == "struct":
        return _arr(len(t[1])) + b"".join(encode(ft, v[name]) for name, ft in t[1])
    if kind == "unit_enum":
        assert v in t[1]
        return _str(v)
    if kind == "enum":
        name, payload = v
        return b"\x81" + _str(name) + encode(dict(t[1])[name], payload)
    raise TypeError(t)


# ---- strict decoder (accepts exactly what the encoder above emits) ------------------------------
class _R:
    def __init__(self, b):
        self.b, self.i = bytes(b), 0

    def take(self, n):
        if self.i + n > len(self.b):
            raise ValueError("truncated")
        s = self.b[self.i:self.i + n]
        self.i += n
        return s

    def byte(self):
        return self.take(1)[0]

    def arr_len(self):
        c = self.byte()
        if 0x90 <= c <= 0x9f:
            return c & 15
        if c == 0xdc:
            return struct.unpack(">H", self.take(2))[0]
        if c == 0xdd:
            return struct.unpack(">I", self.take(4))[0]
        raise ValueError(f"expected array, got {c:#x}")

    def str_(self):
        c = self.byte()
        n = c & 31 if 0xa0 <= c <= 0xbf else self.byte() if c == 0xd9 else \
            struct.unpack(">H", self.take(2))[0] if c == 0xda else struct.unpack(">I", self.take(4))[0] if c == 0xdb else None
        if n is None:
            raise ValueError(f"expected str, got {c:#x}")
        return self.take(n).decode()


def decode(t, r):
    if t == STR:
        return r.str_()
    if t == F32:
        if r.byte() != 0xca:
            raise ValueError("expected f32")
        return struct.unpack(">f", r.take(4))[0]
    if t == F64:
        if r.byte() != 0xcb:
            raise ValueError("expected f64")
        return struct.unpack(">d", r.take(8))[0]
    if t == BOOL:
        c = r.byte()
        if c not in (0xc2, 0xc3):
            raise ValueError("expected bool")
        return c == 0xc3
    if t == UINT:
        c = r.byte()
        if c < 128:
            return c
        n = {0xcc: 1, 0xcd: 2, 0xce: 4, 0xcf: 8}.get(c)
        if n is None:
            raise ValueError(f"expected uint, got {c:#x}")
        return int.from_bytes(r.take(n), "big")
    kind = t[0]
    if kind == "opt":
        if r.b[r.i] == 0xc0:
            r.i += 1
            return None
        return decode(t[1], r)
    if kind == "seq":
        return [decode(t[1], r) for _ in range(r.arr_len())]
    if kind == "tup":
        if r.arr_len() != len(t[1]):
            raise ValueError("tuple arity")
        return tuple(decode(tt, r) for tt in t[1])
    if kind == "struct":
        if r.arr_len() != len(t[1]):
            raise ValueError("struct arity")
        return {name: decode(ft, r) for name, ft in t[1]}
    if kind == "unit_enum":
        s = r.str_()
        if s not in t[1]:
            raise ValueError("unknown unit variant")
        return s
    if kind == "enum":
        if r.byte() != 0x81:
            raise ValueError("expected a 1-entry map")
        name = r.str_()
        return name, decode(dict(t[1])[name], r)
    raise TypeError(t)


def dumps(file_format):
    return encode(FILE_FORMAT, file_format)


def loads(data):
    r = _R(data)
    v = decode(FILE_FORMAT, r)
    if r.i != len(r.b):
        raise ValueError("trailing bytes")
    return v


# ---- X::new(&AudioConfig) as the serializer sees it ---------------------------------------------
def new_module(variant, id_, sample_rate=48000, buffer_size=1024, channels=2, **over):
    z = lambda: [0.0] * buffer_size  # AudioBuffer::new(Some(buffer_size)): zeros (synth.rs:32)
    det = {"last": True}             # TransitionDetector::new() (synth.rs:283)
    grid_tail = dict(octaves=2, steps_per_octave=12, current_step=0, transition_detector=dict(det),
                     sync_transition_detector=dict(det), last=0.0, ui_dirty=True)
    moog_state = dict(f=0.0, p=0.0, q=0.0, b=(0.0,) * 5, freq=0.0, res=0.0)
    m = {
        "OutputModuleV0": lambda: dict(bufs=[z() for _ in range(channels)]),
        "OscillatorModuleV0": lambda: dict(val=0.0, sample_rate=sample_rate, sine=z(), square=z(), saw=z(), pos=0.0,
                                           antialiasing=True, sync_detector=dict(det)),
        "NoiseModuleV0": lambda: dict(out=z()),
        "GridSequencerModuleV0": lambda: dict(cv_out=z(), gate_out=z(), sync_out=z(), sequence=[None] * 64, **grid_tail),
        "GridSequencerModuleV1": lambda: dict(cv_out=z(), gate_out=z(), sync_out=z(), sequence=[None] * 64, **grid_tail),
        "PatternSequencerModuleV0": lambda: dict(gate_outs=[z() for _ in range(8)], sync_out=z(),
                                                 sequence=[[None] * 64 for _ in range(8)], current_step=0,
                                                 transition_detector=dict(det), sync_transition_detector=dict(det),
                                                 ui_dirty=True),
        "ADSRModuleV0": lambda: dict(a_sec=0.0, d_sec=0.5, s_val=0.25, r_sec=0.5, phase=0.0, mode="None", r_val=0.0,
                                     from_a_val=0.0, sample_rate=float(sample_rate), transition_detector=dict(det),
                                     output_buffer=z(), ui_dirty=True),
        "VCAModuleV0": lambda: dict(buf=z(), negative=False),
        "MoogFilterModuleV0": lambda: dict(buf=z(), freq=0.2, res=0.5, exp_amt=0.5, state=dict(moog_state)),
        "MoogFilterModuleV1": lambda: dict(lowpass=z(), bandpass=z(), highpass=z(), freq=0.2, res=0.5, exp_amt=0.5,
                                           state=dict(moog_state)),
        "MonoMixerModuleV0": lambda: dict(gain=[1.0] * 4, buf=z()),
        "SampleModuleV0": lambda: dict(transition_detector=dict(det), pos=0.0, buf=z(),
                                       wavebox=dict(samples=[], sample_rate=0.0, new=False), playing=False,
                                       sample_rate=float(sample_rate)),
        "MathModuleV0": lambda: dict(buf=z(), constant=0.0, operation="Add"),
        "NonLinearModuleV0": lambda: dict(buf=z(), constant=1.0),
        "FreeverbModuleV0": lambda: dict(left_out=z(), right_out=z(), sample_rate=sample_rate, dampening=0.5,
                                         dampening_ctl=0.5, freeze=False, freeze_ctl=False, wet=0.33, wet_ctl=0.33,
                                         width=0.5, width_ctl=0.5, room_size=0.5, room_size_ctl=0.5, dry=0.0, dry_ctl=0.0),
    }[variant]()
    m["id"] = id_
    m.update(over)
    return variant, m


# ---- what SynthModuleWorkspaceImpl::deserialize (ui.rs:115-134) builds from a file, on the oracle ----
_KIND_OF = {"OutputModuleV0": "OUTPUT", "OscillatorModuleV0": "OSCILLATOR", "NoiseModuleV0": "NOISE",
            "GridSequencerModuleV0": "GRID_SEQUENCER", "GridSequencerModuleV1": "GRID_SEQUENCER",
            "PatternSequencerModuleV0": "PATTERN_SEQUENCER", "ADSRModuleV0": "ADSR", "VCAModuleV0": "VCA",
            "MoogFilterModuleV0": "MOOG_FILTER", "MoogFilterModuleV1": "MOOG_FILTER", "MonoMixerModuleV0": "MONO_MIXER",
            "SampleModuleV0": "SAMPLE", "NonLinearModuleV0": "NON_LINEAR"}
_N_INPUTS = {"OUTPUT": None, "OSCILLATOR": 2, "NOISE": 0, "GRID_SEQUENCER": 2, "PATTERN_SEQUENCER": 2, "ADSR": 1,
             "VCA": 2, "MOOG_FILTER": 2, "MONO_MIXER": 4, "SAMPLE": 2, "ADD": 2, "SUBTRACT": 2, "MULTIPLY": 2,
             "NON_LINEAR": 2}
_N_OUTPUTS = {"OUTPUT": 0, "OSCILLATOR": 3, "MOOG_FILTER": 3, "GRID_SEQUENCER": 3, "PATTERN_SEQUENCER": 9}


def _f32(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def state_words(variant, m):
    """The DSP state a serialized module carries, in the device's per-voice state-word layout
    (s-rack_b200/csrc/program.hpp): what a voice starts from after a load.  None = the kind has none."""
    det = lambda d: 1 if d["last"] else 0
    if variant == "OscillatorModuleV0":
        b = struct.unpack("<Q", struct.pack("<d", m["pos"]))[0]
        return [b & 0xFFFFFFFF, b >> 32, det(m["sync_detector"])]
    if variant in ("MoogFilterModuleV0", "MoogFilterModuleV1"):
        st = m["state"]
        return [_f32(st["f"]), _f32(st["p"]), _f32(st["q"])] + [_f32(x) for x in st["b"]] + [_f32(st["freq"]), _f32(st["res"])]
    if variant == "ADSRModuleV0":
        mode = ["Attack", "Decay", "Sustain", "Release", "None"].index(m["mode"])
        return [_f32(m["phase"]), _f32(m["r_val"]), _f32(m["from_a_val"]), mode | (det(m["transition_detector"]) << 8)]
    if variant in ("GridSequencerModuleV0", "GridSequencerModuleV1"):
        return [m["current_step"] | (det(m["transition_detector"]) << 16) | (det(m["sync_transition_detector"]) << 17),
                _f32(m["last"])]
    if variant == "PatternSequencerModuleV0":
        return [m["current_step"] | (det(m["transition_detector"]) << 16) | (det(m["sync_transition_detector"]) << 17)]
    if variant == "SampleModuleV0":
        pos, playing = (0.0, False) if m["wavebox"]["new"] else (m["pos"], m["playing"])  # sample.rs:212-216
        return [_f32(pos), (1 if playing else 0) | (det(m["transition_detector"]) << 1)]
    return None


def build(backend, file_format, channels=2):
    """Apply a decoded file to any backend with the patch verbs (module_create, connect, set_param,
    set_sequence, set_sample) -> {id: handle}.  Module list = the file's reversed (unpack_modules pops
    from the back, ui.rs:652-660); connections back to front, unknown ids / bad ports skipped
    (ui.rs:662-681).  The DSP state in the file is applied by the caller (state_words); port buffers are not."""
    import numpy as np
    handles, kinds = {}, {}
    for variant, m in reversed(file_format["modules"]):
        kind = _KIND_OF.get(variant) or m["operation"].upper()
        h = backend.module_create(kind)
        handles[m["id"]], kinds[m["id"]] = h, kind
        if kind == "OSCILLATOR":
            backend.set_param(h, 0, m["val"])
            backend.set_param(h, 1, 1.0 if m["antialiasing"] else 0.0)
        elif kind == "ADSR":
            for pid, name in enumerate(("a_sec", "d_sec", "s_val", "r_sec")):
                backend.set_param(h, pid, m[name])
        elif kind == "VCA":
            backend.set_param(h, 0, 1.0 if m["negative"] else 0.0)
        elif kind == "MOOG_FILTER":
            for pid, name in enumerate(("freq", "res", "exp_amt")):
                backend.set_param(h, pid, m[name])
        elif kind == "MONO_MIXER":
            for pid, g in enumerate(m["gain"]):
                backend.set_param(h, pid, g)
        elif kind in ("ADD", "SUBTRACT", "MULTIPLY", "NON_LINEAR"):
            backend.set_param(h, 0, m["constant"])
        elif kind == "GRID_SEQUENCER":
            cells = []
            for c in m["sequence"]:
                if c is None:
                    cells.append(-1)
                elif variant == "GridSequencerModuleV0":
                    cells.append(c & 0xFFFF)  # (v, false), sequencer.rs:651-655
                else:
                    cells.append((c[0] & 0xFFFF) | (0x10000 if c[1] else 0))
            backend.set_sequence(h, np.array(cells, dtype=np.int32))
            backend.set_param(h, 0, m["steps_per_octave"])
        elif kind == "PATTERN_SEQUENCER":
            backend.set_sequence(h, np.array([[-1 if c is None else int(c) for c in row] for row in m["sequence"]],
                                             dtype=np.int32))
        elif kind == "SAMPLE":
            backend.set_sample(h, np.array(m["wavebox"]["samples"], dtype=np.float32), m["wavebox"]["sample_rate"])
    for src_id, src_port, sink_id, sink_port in reversed(file_format["connections"]):
        if src_id not in handles or sink_id not in handles or src_id == sink_id:
            continue
        n_in = _N_INPUTS[kinds[sink_id]]
        if sink_port >= (channels if n_in is None else n_in) or src_port >= _N_OUTPUTS.get(kinds[src_id], 1):
            continue
        backend.connect(handles[sink_id], sink_port, handles[src_id], src_port)
    return handles

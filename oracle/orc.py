"""ctypes wrapper over the CPU oracle (oracle/srack_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(s-rack_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsrack_oracle.so")

# numeric ids, equal to include/srack_b200.h (tests/test_abi.py checks that)
KIND = dict(OUTPUT=0, OSCILLATOR=1, NOISE=2, ADSR=3, VCA=4, MOOG_FILTER=5, MONO_MIXER=6,
            ADD=7, SUBTRACT=8, MULTIPLY=9, NON_LINEAR=10, GRID_SEQUENCER=11, PATTERN_SEQUENCER=12, SAMPLE=13)


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "srack_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_patch_create.restype = C.c_void_p
        L.orc_patch_create.argtypes = [C.c_uint16, C.c_size_t, C.c_uint8]
        L.orc_patch_destroy.argtypes = [C.c_void_p]
        L.orc_set_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_module_create.argtypes = [C.c_void_p, C.c_int]
        L.orc_num_inputs.argtypes = [C.c_void_p, C.c_int]
        L.orc_num_outputs.argtypes = [C.c_void_p, C.c_int]
        L.orc_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_disconnect.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.orc_set_param_per_voice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_set_module_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_set_sequence.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_set_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_float]
        L.orc_exp2f.restype = C.c_float
        L.orc_exp2f.argtypes = [C.c_float]
        L.orc_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_debug_prepare.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_debug_calc.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.orc_debug_execute.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_debug_output.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.orc_debug_set_output.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_version.restype = C.c_char_p
        _lib = L
    return _lib


class OraclePatch:
    """One patch graph + a bank of per-voice instances, rendered on the CPU."""

    def __init__(self, sample_rate=48000, buffer_size=1024, channels=2):
        self.sample_rate, self.buffer_size, self.channels = sample_rate, buffer_size, channels
        self._h = lib().orc_patch_create(sample_rate, buffer_size, channels)
        self.n_modules = 0

    def close(self):
        if self._h:
            lib().orc_patch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- graph building (same verbs as the product's Patch) --
    def set_seed(self, seed):
        lib().orc_set_seed(self._h, seed)

    def module_create(self, kind):
        k = KIND[kind] if isinstance(kind, str) else int(kind)
        m = lib().orc_module_create(self._h, k)
        if m < 0:
            raise ValueError(f"bad kind {kind}")
        self.n_modules += 1
        return m

    def connect(self, sink, in_idx, src, src_port):
        rc = lib().orc_connect(self._h, sink, in_idx, src, src_port)
        if rc:
            raise ValueError(f"connect failed rc={rc}")

    def disconnect(self, sink, in_idx):
        rc = lib().orc_disconnect(self._h, sink, in_idx)
        if rc:
            raise ValueError(f"disconnect failed rc={rc}")

    def set_param(self, module, pid, value):
        rc = lib().orc_set_param(self._h, module, pid, float(value))
        if rc:
            raise ValueError(f"set_param failed rc={rc}")

    def set_param_per_voice(self, module, pid, values):
        v = np.ascontiguousarray(values, dtype=np.float32)
        rc = lib().orc_set_param_per_voice(self._h, module, pid, v.ctypes.data, v.size)
        if rc:
            raise ValueError(f"set_param_per_voice failed rc={rc}")

    def set_sequence(self, module, cells):
        """Grid sequencer: cells[n_steps]; pattern sequencer: cells[8][n_steps] (int32, see srack_b200.h)."""
        c = np.ascontiguousarray(cells, dtype=np.int32)
        rc = lib().orc_set_sequence(self._h, module, c.ctypes.data, c.shape[-1])
        if rc:
            raise ValueError(f"set_sequence failed rc={rc}")

    def set_sample(self, module, samples, sample_rate):
        """The WaveBox of a Sample module after a load (sample.rs:32-69): channel-0 f32 samples + file rate."""
        a = np.ascontiguousarray(samples, dtype=np.float32)
        rc = lib().orc_set_sample(self._h, module, a.ctypes.data, a.size, float(sample_rate))
        if rc:
            raise ValueError(f"set_sample failed rc={rc}")

    def load_wav(self, module, data):
        """WaveBox::load (sample.rs:32-69) through the numpy restatement in oracle/wav.py."""
        from . import wav
        samples, rate = wav.load(data)
        self.set_sample(module, samples, rate)

    def set_state(self, module, words):
        """Deserialized DSP state of one module as device state words (uint32), see srk_file.state_words."""
        w = np.ascontiguousarray(words, dtype=np.uint32)
        rc = lib().orc_set_state(self._h, module, w.ctypes.data, w.size)
        if rc:
            raise ValueError(f"set_state failed rc={rc}")

    def set_adsr_sample_rate(self, module, sample_rate):
        """The `sample_rate` field an ADSR carries through a .srk file (set_audio_config never updates it)."""
        self.set_param(module, 100, sample_rate)

    def load_srk(self, data):
        """SynthModuleWorkspaceImpl::deserialize (ui.rs:115-134) via oracle/srk_file.py -> {id: module}."""
        from . import srk_file
        ff = srk_file.loads(data)
        handles = srk_file.build(self, ff, channels=self.channels)
        for variant, m in ff["modules"]:
            if variant == "ADSRModuleV0":
                self.set_adsr_sample_rate(handles[m["id"]], m["sample_rate"])
            words = srk_file.state_words(variant, m)
            if words is not None:
                self.set_state(handles[m["id"]], words)
        return handles

    def set_module_order(self, order):
        o = np.ascontiguousarray(order, dtype=np.int32)
        lib().orc_set_module_order(self._h, o.ctypes.data, o.size)

    def plan(self):
        """-> (plan order as module indices, cut wires as (reader, writer) pairs)"""
        n = self.n_modules
        out = np.zeros(n + 1, dtype=np.int32)
        cuts = np.zeros(2 * (4 * n + 4), dtype=np.int32)
        n_out, n_cuts = C.c_int(0), C.c_int(0)
        rc = lib().orc_plan(self._h, out.ctypes.data, C.byref(n_out), cuts.ctypes.data, C.byref(n_cuts))
        if rc:
            return [], []
        return out[: n_out.value].tolist(), [tuple(cuts[2 * i: 2 * i + 2].tolist()) for i in range(n_cuts.value)]

    def reset(self):
        lib().orc_reset(self._h)

    def render(self, n_voices, n_samples, voice_offset=0, stems=True, mix=True, n_threads=1):
        """-> (stems f32 [C][T][V] or None, mix f64 [C][T] or None)"""
        st = np.zeros((self.channels, n_samples, n_voices), dtype=np.float32) if stems else None
        mx = np.zeros((self.channels, n_samples), dtype=np.float64) if mix else None
        rc = lib().orc_render(self._h, n_voices, voice_offset, n_samples,
                              st.ctypes.data if stems else None, mx.ctypes.data if mix else None, n_threads)
        if rc:
            raise RuntimeError(f"oracle render failed rc={rc}")
        return st, mx

    # -- KAT hooks --
    def debug_prepare(self, n_voices=1):
        rc = lib().orc_debug_prepare(self._h, n_voices)
        if rc:
            raise RuntimeError(f"prepare failed rc={rc}")

    def debug_calc(self, module, voice=0):
        assert lib().orc_debug_calc(self._h, voice, module) == 0

    def debug_execute(self, voice=0):
        assert lib().orc_debug_execute(self._h, voice) == 0

    def debug_output(self, module, port, voice=0):
        out = np.zeros(self.buffer_size, dtype=np.float32)
        assert lib().orc_debug_output(self._h, voice, module, port, out.ctypes.data) == 0
        return out

    def debug_set_output(self, module, port, data, voice=0):
        """Overwrite an output buffer (one block): the modules reading that port see `data` at their next calc()."""
        d = np.ascontiguousarray(data, dtype=np.float32)
        assert d.size == self.buffer_size
        assert lib().orc_debug_set_output(self._h, voice, module, port, d.ctypes.data) == 0


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o

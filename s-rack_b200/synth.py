"""Host-side mirror of the reference's module-graph API (src/synth.rs) over the C ABI.

Names follow the reference: `AudioConfig` (synth.rs:20-25), `SynthModule` with
`get_id/get_name/get_num_inputs/get_num_outputs/get_input/get_input_label/
get_output_label/set_input/disconnect_input/disconnect_inputs` (synth.rs:222-246),
`plan_execution` (synth.rs:128), `execute` (synth.rs:97), `get_inputs` (synth.rs:214),
`get_catalog` (synth.rs:421).  The reference's `Err(())` becomes `PortError`.
The new axis is voices: a `Patch` renders `n_voices` independent instances of its
graph in lockstep on one B200.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi
from ._ffi import KIND, PARAM, STATUS, lib

_STATUS_NAME = {v: k for k, v in STATUS.items()}
_KIND_NAME = {v: k for k, v in KIND.items()}


class SrackError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        name = _STATUS_NAME.get(status, str(status))
        super().__init__(f"SRK_{name}: {detail or lib.srk_status_string(status).decode()}")


class PortError(SrackError):
    """The reference's Err(()) for an out-of-range port index."""


@dataclass
class AudioConfig:
    sample_rate: int = 48000
    buffer_size: int = 1024
    channels: int = 2

    def _c(self):
        return _ffi.srk_audio_config(self.sample_rate, self.buffer_size, self.channels)


class SynthModule:
    """One module of a patch (the reference's SharedSynthModule). Identity == handle."""

    def __init__(self, patch, handle):
        self._patch = patch
        self._h = handle

    def __eq__(self, other):  # shared_are_eq, synth.rs:272-274
        return isinstance(other, SynthModule) and other._h == self._h

    def __hash__(self):
        return hash(self._h)

    def __repr__(self):
        return f"<{self.get_name()} {self.get_id()[:8]}>"

    def _check(self, rc):
        if rc == STATUS["ERR_PORT"]:
            raise PortError(rc, self._patch._err())
        if rc:
            raise SrackError(rc, self._patch._err())

    def get_id(self):
        return lib.srk_get_id(self._h).decode()

    def get_name(self):
        return lib.srk_get_name(self._h).decode()

    def get_kind(self):
        return _KIND_NAME[lib.srk_get_kind(self._h)]

    def get_num_inputs(self):
        return lib.srk_get_num_inputs(self._h)

    def get_num_outputs(self):
        return lib.srk_get_num_outputs(self._h)

    def get_input(self, input_idx):
        src, port = C.c_void_p(), C.c_uint8()
        self._check(lib.srk_get_input(self._h, input_idx, C.byref(src), C.byref(port)))
        if not src.value:
            return None
        return self._patch._wrap(src.value), port.value

    def get_input_label(self, input_idx):
        s = C.c_char_p()
        self._check(lib.srk_get_input_label(self._h, input_idx, C.byref(s)))
        return s.value.decode() if s.value is not None else None

    def get_output_label(self, output_idx):
        s = C.c_char_p()
        self._check(lib.srk_get_output_label(self._h, output_idx, C.byref(s)))
        return s.value.decode() if s.value is not None else None

    def set_input(self, input_idx, src_module, src_port):
        self._check(lib.srk_connect(self._h, input_idx, src_module._h, src_port))

    def disconnect_input(self, input_idx):
        self._check(lib.srk_disconnect(self._h, input_idx))

    def disconnect_inputs(self):
        self._check(lib.srk_disconnect_inputs(self._h))

    # parameters: struct fields in the reference, ids from include/srack_b200.h
    def set_param(self, param, value):
        self._check(lib.srk_set_param_f32(self._h, _pid(param), float(value)))
        self._patch.per_voice.pop((self, _pid(param)), None)

    def get_param(self, param):
        v = C.c_float()
        self._check(lib.srk_get_param_f32(self._h, _pid(param), C.byref(v)))
        return v.value

    def set_param_per_voice(self, param, values):
        a = np.ascontiguousarray(values, dtype=np.float32)
        self._check(lib.srk_set_param_f32_per_voice(self._h, _pid(param), a.ctypes.data, a.size))
        self._patch.per_voice[(self, _pid(param))] = a  # host copy, e.g. to re-send or re-shard


    def set_sequence(self, cells):
        """The step table a sequencer's ui() edits (sequencer.rs:98-188, :388-470): int32 cells,
        [n_steps] for the grid sequencer (SEQ_NONE or grid_cell(val, hold)), [8][n_steps] for the
        pattern sequencer (SEQ_NONE, 0 = pass the clock, 1 = hold)."""
        a = np.ascontiguousarray(cells, dtype=np.int32)
        self._check(lib.srk_set_sequence(self._h, a.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[-1]))

    def get_sequence(self):
        n = C.c_size_t()
        self._check(lib.srk_get_sequence(self._h, None, 0, C.byref(n)))
        rows = 8 if self.get_kind() == "PATTERN_SEQUENCER" else 1
        buf = (C.c_int32 * (rows * n.value))()
        self._check(lib.srk_get_sequence(self._h, buf, rows * n.value, C.byref(n)))
        a = np.array(buf, dtype=np.int32)
        return a.reshape(rows, n.value) if rows > 1 else a

    # Sample module: the WaveBox behind "Load Sample..." (sample.rs:14-20, 242-257)
    def load_wav(self, data):
        """WaveBox::load (sample.rs:32-69): WAV file bytes -> channel-0 f32 table + the file's rate."""
        data = bytes(data)
        self._check(lib.srk_load_wav(self._h, data, len(data)))

    def set_sample(self, samples, sample_rate):
        a = np.ascontiguousarray(samples, dtype=np.float32)
        self._check(lib.srk_set_sample(self._h, a.ctypes.data, a.size, float(sample_rate)))

    def get_sample(self):
        """-> (samples f32, sample_rate)"""
        n, rate = C.c_size_t(), C.c_float()
        self._check(lib.srk_get_sample(self._h, None, 0, C.byref(n), C.byref(rate)))
        a = np.empty(n.value, dtype=np.float32)
        self._check(lib.srk_get_sample(self._h, a.ctypes.data, a.size, C.byref(n), C.byref(rate)))
        return a, rate.value


def _pid(param):
    return PARAM[param] if isinstance(param, str) else int(param)


def get_inputs(module):
    """synth.rs:214-218"""
    return [module.get_input(i) for i in range(module.get_num_inputs())]


class Patch:
    """The module list + plan of one patch (ui.rs:51-60), rendered for many voices."""

    def __init__(self, audio_config=None, device=None):
        self.audio_config = audio_config or AudioConfig()
        h = C.c_void_p()
        cfg = self.audio_config._c()
        rc = lib.srk_patch_create(C.byref(cfg), C.byref(h))
        if rc:
            raise SrackError(rc)
        self._h = h
        self._wrappers = {}
        self.per_voice = {}  # (module, param id) -> f32 array indexed by global voice
        if device is not None:
            self._check(lib.srk_set_device(self._h, int(device)))

    def close(self):
        if getattr(self, "_h", None):
            lib.srk_patch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return lib.srk_last_error(self._h).decode()

    def _check(self, rc):
        if rc == STATUS["ERR_PORT"]:
            raise PortError(rc, self._err())
        if rc:
            raise SrackError(rc, self._err())

    def _wrap(self, handle):
        w = self._wrappers.get(handle)
        if w is None:
            w = self._wrappers[handle] = SynthModule(self, handle)
        return w

    # -- module list ---------------------------------------------------------
    def add_module(self, kind):
        """`kind`: catalog name ("Moog Filter"), enum name ("MOOG_FILTER") or srk_kind int."""
        h = C.c_void_p()
        if isinstance(kind, str) and kind in KIND:
            rc = lib.srk_module_create(self._h, KIND[kind], C.byref(h))
        elif isinstance(kind, str):
            rc = lib.srk_module_create_by_name(self._h, kind.encode(), C.byref(h))
        else:
            rc = lib.srk_module_create(self._h, int(kind), C.byref(h))
        self._check(rc)
        return self._wrap(h.value)

    def delete_module(self, module):
        self._check(lib.srk_module_remove(self._h, module._h))
        self._wrappers.pop(module._h, None)

    @property
    def modules(self):
        return [self._wrap(lib.srk_module_at(self._h, i)) for i in range(lib.srk_module_count(self._h))]

    def set_module_order(self, modules):
        arr = (C.c_void_p * len(modules))(*[m._h for m in modules])
        self._check(lib.srk_set_module_order(self._h, arr, len(modules)))

    def set_audio_config(self, audio_config):
        cfg = audio_config._c()
        self._check(lib.srk_set_audio_config(self._h, C.byref(cfg)))
        self.audio_config = audio_config

    def set_seed(self, seed):
        self._check(lib.srk_set_seed(self._h, int(seed)))

    # -- same verbs as oracle.orc.OraclePatch so one patch description drives both ----
    def module_create(self, kind):
        return self.add_module(kind)

    def connect(self, sink, in_idx, src, src_port):
        sink.set_input(in_idx, src, src_port)

    def disconnect(self, sink, in_idx):
        sink.disconnect_input(in_idx)

    def set_param(self, module, pid, value):
        module.set_param(pid, value)

    def set_param_per_voice(self, module, pid, values):
        module.set_param_per_voice(pid, values)

    def set_sequence(self, module, cells):
        module.set_sequence(cells)

    def set_sample(self, module, samples, sample_rate):
        module.set_sample(samples, sample_rate)

    def load_wav(self, module, data):
        module.load_wav(data)

    # -- .srk patch files (FileFormat, ui.rs:578-586) ------------------------
    def load_srk(self, data):
        """SynthModuleWorkspaceImpl::deserialize (ui.rs:115-134): replace this patch by the file's.
        -> number of connections skipped (unknown ids / bad ports, like the reference's `let _ =`)."""
        data = bytes(data)
        skipped = C.c_size_t()
        self._check(lib.srk_patch_load_srk(self._h, data, len(data), C.byref(skipped)))
        self._wrappers.clear()
        self.per_voice.clear()
        return skipped.value

    def save_srk(self):
        """SynthModuleWorkspaceImpl::serialize (ui.rs:98-114) -> bytes."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(lib.srk_patch_save_srk(self._h, C.byref(ptr), C.byref(n)))
        return C.string_at(ptr.value, n.value)

    def module_by_id(self, module_id):
        for m in self.modules:
            if m.get_id() == module_id:
                return m
        raise KeyError(module_id)

    # -- planning ------------------------------------------------------------
    def plan(self):
        """plan_execution (synth.rs:128-212) -> modules in execution order."""
        self._check(lib.srk_plan(self._h))
        n = C.c_size_t()
        cap = lib.srk_module_count(self._h)
        arr = (C.c_void_p * max(cap, 1))()
        self._check(lib.srk_plan_get(self._h, arr, cap, C.byref(n)))
        return [self._wrap(arr[i]) for i in range(n.value)]

    def plan_cuts(self):
        """Wires removed by the cycle breaker, as (reader, writer) pairs."""
        n = C.c_size_t()
        self._check(lib.srk_plan_cuts(self._h, None, None, 0, C.byref(n)))
        r = (C.c_void_p * max(n.value, 1))()
        w = (C.c_void_p * max(n.value, 1))()
        self._check(lib.srk_plan_cuts(self._h, r, w, n.value, C.byref(n)))
        return [(self._wrap(r[i]), self._wrap(w[i])) for i in range(n.value)]

    def program_info(self, n_voices=0):
        info = _ffi.srk_program_info()
        self._check(lib.srk_get_program_info(self._h, n_voices, C.byref(info)))
        return {f: getattr(info, f) for f, _ in info._fields_}

    def fused_source(self, n_voices=0):
        """CUDA C++ generated for this patch when a render of n_voices would use a fused kernel, else ''."""
        src, n = C.c_char_p(), C.c_size_t()
        self._check(lib.srk_fused_source(self._h, n_voices, C.byref(src), C.byref(n)))
        return (src.value or b"").decode()

    def precompile(self, n_voices=0):
        """Compile the fused kernels a render of n_voices can launch -- the cost model's choice and the alternatives a long
        render measures against it -- into the on-disk cubin cache (NVRTC, no GPU needed) -> how many were compiled now."""
        done = C.c_int()
        self._check(lib.srk_precompile(self._h, n_voices, C.byref(done)))
        return int(done.value)

    def kernel_id(self, n_voices=0):
        """Identity of the kernel image a render of n_voices would launch (profiles are stamped with it)."""
        s = C.c_char_p()
        self._check(lib.srk_kernel_id(self._h, n_voices, C.byref(s)))
        return (s.value or b"").decode()

    def schedule_report(self):
        """How the launch shape in use was chosen: '' or the measured candidates and the winner."""
        s = C.c_char_p()
        self._check(lib.srk_schedule_report(self._h, C.byref(s)))
        return (s.value or b"").decode()

    def set_co_resident_voices(self, n_voices):
        """Voices other patches render on this device concurrently: the launch is scheduled for the sum."""
        self._check(lib.srk_set_co_resident_voices(self._h, int(n_voices)))

    # -- per-voice DSP state: checkpoint / resume of a render (the reference serializes it with the patch, ui.rs:98-134)
    def state_export(self):
        """-> bytes: state words, feedback history and sample index of the voice range last rendered."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(lib.srk_state_export(self._h, C.byref(ptr), C.byref(n)))
        return C.string_at(ptr.value, n.value)

    def state_import(self, blob):
        blob = bytes(blob)
        self._check(lib.srk_state_import(self._h, blob, len(blob)))

    def program(self, n_voices=0):
        """The compiled, scheduled device program for n_voices voices -> (instrs, wires): lists of
        dicts (op name, flags, warp, stage, in/out wire slots, ...) and (first_tile, n_tiles)."""
        ni, nw = C.c_size_t(0), C.c_size_t(0)
        self._check(lib.srk_get_program(self._h, n_voices, None, 0, C.byref(ni), None, 0, C.byref(nw)))
        instrs = (_ffi.srk_instr_info * max(ni.value, 1))()
        wires = (_ffi.srk_wire_info * max(nw.value, 1))()
        self._check(lib.srk_get_program(self._h, n_voices, instrs, ni.value, C.byref(ni), wires, nw.value, C.byref(nw)))
        out = [dict(op=_ffi.OPS[i.op], flags=i.flags, warp=i.warp, stage=i.stage, ins=list(i.in_), outs=list(i.out),
                    n_ch=i.n_ch, state=i.state, param=i.param, aux=i.aux) for i in instrs[:ni.value]]
        return out, [(w.first_tile, w.n_tiles) for w in wires[:nw.value]]

    # -- rendering -----------------------------------------------------------
    def render(self, n_voices, n_samples, voice_offset=0, stems=False, mix=True):
        """Render to freshly allocated host arrays -> (stems [C][N][V] or None, mix [C][N] or None)."""
        ch = self.audio_config.channels
        st = np.empty((ch, n_samples, n_voices), dtype=np.float32) if stems else None
        mx = np.empty((ch, n_samples), dtype=np.float32) if mix else None
        self.render_into(n_voices, n_samples, voice_offset, st.ctypes.data if stems else None,
                         mx.ctypes.data if mix else None)
        return st, mx

    def render_into(self, n_voices, n_samples, voice_offset=0, stems_ptr=None, mix_ptr=None, device_out=False,
                    async_=False, stream=None):
        """Raw-pointer render (host pointers, or device pointers with device_out=True)."""
        flags = (_ffi.RENDER_DEVICE_OUT if device_out else 0) | (_ffi.RENDER_ASYNC if async_ else 0)
        if stream is None:
            rc = lib.srk_render(self._h, n_voices, voice_offset, n_samples, flags, stems_ptr, mix_ptr)
        else:
            rc = lib.srk_render_on_stream(self._h, n_voices, voice_offset, n_samples, flags, stems_ptr, mix_ptr,
                                          C.c_void_p(stream))
        self._check(rc)

    def execute(self, n_voices, voice_offset=0, stems=False, mix=True):
        """One `execute(&plan)` of the reference (synth.rs:97-101): one block of buffer_size samples."""
        return self.render(n_voices, self.audio_config.buffer_size, voice_offset, stems, mix)

    def sync(self):
        self._check(lib.srk_sync(self._h))

    def reset(self):
        self._check(lib.srk_reset(self._h))

    def last_render_ms(self):
        k, t = C.c_float(), C.c_float()
        self._check(lib.srk_last_render_ms(self._h, C.byref(k), C.byref(t)))
        return k.value, t.value

    def launch_count(self):
        return lib.srk_launch_count(self._h)

    def state_epoch(self):
        """How often the voice state was (re)initialised (implicitly, too: a render of another voice range resets it)."""
        return lib.srk_state_epoch(self._h)


def plan_execution(patch):
    """synth.rs:128: output = first Output in the list (ui.rs:84-96), all_modules = the patch's list."""
    return patch.plan()


def execute(patch, n_voices, **kw):
    return patch.execute(n_voices, **kw)


def write_wav(path, planar, sample_rate=48000, bits=32):
    """WAV export of a render (`mix` [C][N] as Patch.render returns it): 32-bit float or 16/24-bit PCM."""
    a = np.ascontiguousarray(planar, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    rc = lib.srk_write_wav(str(path).encode(), a.ctypes.data, a.shape[0], a.shape[1], int(sample_rate), int(bits))
    if rc:
        raise SrackError(rc, f"cannot write {path}")


def get_catalog():
    """synth.rs:421-515 -> [(name, constructor(patch) -> SynthModule)]; entries outside the hot
    path raise SrackError(SRK_ERR_UNSUPPORTED) when constructed."""
    out = []
    for i in range(lib.srk_catalog_size()):
        name = lib.srk_catalog_name(i).decode()
        if name == "Output":
            continue  # created by the app (main.rs:130), not in the reference catalog
        out.append((name, (lambda n: (lambda patch: patch.add_module(n)))(name)))
    return out

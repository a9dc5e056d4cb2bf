"""ctypes binding of libsrack_b200.so (the C ABI in include/srack_b200.h).

There is no fallback of any kind: if the CUDA library has not been built this
module raises at import time, and rendering without a CUDA device fails with
SRK_ERR_NO_DEVICE.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrack_b200.so")


class srk_audio_config(C.Structure):
    # src/synth.rs:20-25
    _fields_ = [("sample_rate", C.c_uint16), ("buffer_size", C.c_size_t), ("channels", C.c_uint8)]


class srk_program_info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_instr", "step_samples", "block_threads", "smem_bytes",
                                          "n_wires", "state_words", "param_words", "n_rings",
                                          "n_warps", "n_stages", "n_tiles", "groups_per_block",
                                          "fused", "fused_group", "fused_regs", "fused_local_bytes")]


class srk_instr_info(C.Structure):
    _fields_ = [("op", C.c_uint8), ("flags", C.c_uint8), ("warp", C.c_uint8), ("stage", C.c_uint8),
                ("in_", C.c_int16 * 4), ("out", C.c_int16 * 3), ("n_ch", C.c_uint8), ("reserved", C.c_uint8),
                ("state", C.c_uint16), ("param", C.c_uint16), ("aux", C.c_uint16)]


class srk_wire_info(C.Structure):
    _fields_ = [("first_tile", C.c_uint16), ("n_tiles", C.c_uint16)]


from .constants import (KIND, OPS, PARAM, RENDER_ASYNC, RENDER_DEVICE_OUT, SEQ_NONE, STATUS, grid_cell)  # noqa: F401,E402

_P = C.c_void_p
_SIGNATURES = {
    "srk_version": (C.c_char_p, []),
    "srk_status_string": (C.c_char_p, [C.c_int]),
    "srk_catalog_size": (C.c_int, []),
    "srk_catalog_name": (C.c_char_p, [C.c_int]),
    "srk_catalog_kind": (C.c_int, [C.c_int]),
    "srk_patch_create": (C.c_int, [C.POINTER(srk_audio_config), C.POINTER(_P)]),
    "srk_patch_destroy": (None, [_P]),
    "srk_set_audio_config": (C.c_int, [_P, C.POINTER(srk_audio_config)]),
    "srk_get_audio_config": (C.c_int, [_P, C.POINTER(srk_audio_config)]),
    "srk_set_seed": (C.c_int, [_P, C.c_uint64]),
    "srk_set_device": (C.c_int, [_P, C.c_int]),
    "srk_last_error": (C.c_char_p, [_P]),
    "srk_module_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "srk_module_create_by_name": (C.c_int, [_P, C.c_char_p, C.POINTER(_P)]),
    "srk_module_remove": (C.c_int, [_P, _P]),
    "srk_module_count": (C.c_size_t, [_P]),
    "srk_module_at": (_P, [_P, C.c_size_t]),
    "srk_get_id": (C.c_char_p, [_P]),
    "srk_get_name": (C.c_char_p, [_P]),
    "srk_get_kind": (C.c_int, [_P]),
    "srk_get_num_inputs": (C.c_int, [_P]),
    "srk_get_num_outputs": (C.c_int, [_P]),
    "srk_get_input_label": (C.c_int, [_P, C.c_uint8, C.POINTER(C.c_char_p)]),
    "srk_get_output_label": (C.c_int, [_P, C.c_uint8, C.POINTER(C.c_char_p)]),
    "srk_connect": (C.c_int, [_P, C.c_uint8, _P, C.c_uint8]),
    "srk_disconnect": (C.c_int, [_P, C.c_uint8]),
    "srk_disconnect_inputs": (C.c_int, [_P]),
    "srk_get_input": (C.c_int, [_P, C.c_uint8, C.POINTER(_P), C.POINTER(C.c_uint8)]),
    "srk_set_param_f32": (C.c_int, [_P, C.c_int, C.c_float]),
    "srk_get_param_f32": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "srk_set_param_f32_per_voice": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "srk_set_sequence": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_size_t]),
    "srk_get_sequence": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_size_t, C.POINTER(C.c_size_t)]),
    "srk_load_wav": (C.c_int, [_P, _P, C.c_size_t]),
    "srk_set_sample": (C.c_int, [_P, _P, C.c_size_t, C.c_float]),
    "srk_get_sample": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_float)]),
    "srk_write_wav": (C.c_int, [C.c_char_p, _P, C.c_uint, C.c_size_t, C.c_uint32, C.c_int]),
    "srk_patch_load_srk": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "srk_patch_save_srk": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "srk_plan": (C.c_int, [_P]),
    "srk_plan_get": (C.c_int, [_P, C.POINTER(_P), C.c_size_t, C.POINTER(C.c_size_t)]),
    "srk_plan_cuts": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.c_size_t, C.POINTER(C.c_size_t)]),
    "srk_set_module_order": (C.c_int, [_P, C.POINTER(_P), C.c_size_t]),
    "srk_render": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, _P, _P]),
    "srk_render_on_stream": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, _P, _P, _P]),
    "srk_sync": (C.c_int, [_P]),
    "srk_reset": (C.c_int, [_P]),
    "srk_last_render_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "srk_launch_count": (C.c_uint64, [_P]),
    "srk_state_epoch": (C.c_uint64, [_P]),
    "srk_get_program_info": (C.c_int, [_P, C.c_size_t, C.POINTER(srk_program_info)]),
    "srk_fused_source": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]),
    "srk_precompile": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int)]),
    "srk_kernel_id": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_char_p)]),
    "srk_schedule_report": (C.c_int, [_P, C.POINTER(C.c_char_p)]),
    "srk_set_co_resident_voices": (C.c_int, [_P, C.c_size_t]),
    "srk_state_export": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "srk_state_import": (C.c_int, [_P, _P, C.c_size_t]),
    "srk_get_program": (C.c_int, [_P, C.c_size_t, C.POINTER(srk_instr_info), C.c_size_t, C.POINTER(C.c_size_t),
                                  C.POINTER(srk_wire_info), C.c_size_t, C.POINTER(C.c_size_t)]),
}


def _point_at_second_nvrtc():
    """The library builds its patch-specialised kernels with the CUDA toolkit's NVRTC and, when ANOTHER version is at
    hand, measures that one's code as well (csrc/fused_rt.cpp: neither version is faster everywhere).  A Python
    environment with the nvidia-cuda-nvrtc wheel (torch's dependency) has one: name it, so that the choice does not
    depend on whether torch happened to be imported first.  SRK_NVRTC_ALT=0 turns the second compiler off."""
    if "SRK_NVRTC_ALT" in os.environ:
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.cuda_nvrtc")
        for d in (spec.submodule_search_locations if spec else ()):
            cand = os.path.join(d, "lib", "libnvrtc.so.12")
            if os.path.exists(cand):
                os.environ["SRK_NVRTC_ALT"] = cand
                return
    except Exception:
        pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C s-rack_b200/csrc). "
            "srack_b200 has no CPU or pure-Python path.")
    _point_at_second_nvrtc()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()

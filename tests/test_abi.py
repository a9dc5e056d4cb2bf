"""The C ABI library loads without a GPU, exports every symbol include/srack_b200.h declares,
agrees with the Python constants, and reproduces the reference's error behaviour
(Err(()) on bad port indices, labels, port counts) -- no compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "srack_b200.h")


def declared_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SRK_API\s+[^;(]*?\b(srk_\w+)\s*\(", src)))


def header_enum(name):
    src = open(HEADER).read()
    body = re.search(r"enum\s+%s\s*\{(.*?)\};" % name, src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = {}
    for k, v in re.findall(r"(SRK_\w+)\s*=\s*([^,\s]+)", body):
        out[k] = int(eval(v.replace("u", ""))) if "<<" in v else int(v)
    return out


def test_every_declared_symbol_is_exported(srk):
    fns = declared_functions()
    assert len(fns) >= 40
    raw = ctypes.CDLL(srk.LIB_PATH)
    missing = [f for f in fns if not hasattr(raw, f)]
    assert not missing, missing
    # and the Python binding covers the whole header
    from srack_b200 import _ffi
    assert sorted(_ffi._SIGNATURES) == fns


def test_abi_from_plain_c(srk, tmp_path):
    """include/srack_b200.h is valid C99 and the library links and behaves from a C translation unit."""
    import subprocess
    lib_dir = os.path.dirname(srk.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER])
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", lib_dir, "-lsrack_b200",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "ok" in out.stdout


def test_python_constants_match_the_header(srk, orc):
    st = header_enum("srk_status")
    assert {"SRK_" + k: v for k, v in srk.STATUS.items()} == st
    kinds = header_enum("srk_kind")
    kinds.pop("SRK_KIND_COUNT")
    assert {"SRK_KIND_" + k: v for k, v in srk.KIND.items()} == kinds
    assert orc.KIND == srk.KIND  # the oracle uses the same numeric ids
    params = header_enum("srk_param")
    assert {"SRK_" + k: v for k, v in srk.PARAM.items()} == params


def test_catalog_matches_the_reference(srk):
    # get_catalog(), src/synth.rs:421-515 -- names in the reference's order
    names = [n for n, _ in srk.get_catalog()]
    assert names == ["Oscillator", "Noise", "Grid Sequencer", "Pattern Sequencer", "ADSR", "VCA", "Moog Filter",
                     "Mono Mixer", "Sample", "Add", "Subtract", "Multiply", "Non-Linear", "Freeverb"]
    p = srk.Patch()
    for name, make in srk.get_catalog():
        if name == "Freeverb":  # outside the hot path
            with pytest.raises(srk.SrackError) as e:
                make(p)
            assert e.value.status == srk.STATUS["ERR_UNSUPPORTED"]
        else:
            assert make(p).get_name() == name
    assert p.add_module("Output").get_name() == "Output"


PORTS = {  # name: (inputs, outputs, input labels, output labels) from the reference modules
    "Oscillator": (2, 3, ["CV", "Sync"], ["Sine", "Square", "Sawtooth"]),   # oscillator.rs:99-106,160-182
    "Grid Sequencer": (2, 3, ["Step", "Sync"], ["CV", "Gate", "Sync"]),     # sequencer.rs:262-307
    "Pattern Sequencer": (2, 9, ["Step", "Sync"], [str(i) for i in range(8)] + ["Sync"]),  # sequencer.rs:551-596
    "Sample": (2, 1, ["Gate", "CV"], [None]),                              # sample.rs:112-190
    "Noise": (0, 1, [], [None]),                                           # oscillator.rs:344-375
    "ADSR": (1, 1, ["Gate"], [None]),                                      # adsr.rs:73-132
    "VCA": (2, 1, ["Audio", "CV"], [None]),                                # vca.rs
    "Moog Filter": (2, 3, ["Audio", "CV"], [None, None, None]),            # filter.rs:109-180
    "Mono Mixer": (4, 1, [None] * 4, [None]),                              # mixer.rs
    "Add": (2, 1, ["In1", "In2"], [None]),                                 # math.rs:68-137
    "Subtract": (2, 1, ["In1", "In2"], [None]),
    "Multiply": (2, 1, ["In1", "In2"], [None]),
    "Non-Linear": (2, 1, ["In1", "In2"], [None]),                          # math.rs:221-290
    "Output": (2, 0, [None, None], []),                                    # output.rs (channels = 2)
}


@pytest.mark.parametrize("name", sorted(PORTS))
def test_ports_labels_and_err_behaviour(srk, name):
    n_in, n_out, in_labels, out_labels = PORTS[name]
    p = srk.Patch()
    m = p.add_module(name)
    src = p.add_module("Oscillator")
    assert m.get_num_inputs() == n_in and m.get_num_outputs() == n_out
    assert [m.get_input_label(i) for i in range(n_in)] == in_labels
    assert [m.get_output_label(i) for i in range(n_out)] == out_labels
    assert srk.get_inputs(m) == [None] * n_in
    for bad in (n_in, n_in + 1, 255):  # the reference's Err(())
        with pytest.raises(srk.PortError):
            m.get_input(bad)
        with pytest.raises(srk.PortError):
            m.get_input_label(bad)
        with pytest.raises(srk.PortError):
            m.set_input(bad, src, 0)
        with pytest.raises(srk.PortError):
            m.disconnect_input(bad)
    with pytest.raises(srk.PortError):
        m.get_output_label(n_out)
    if n_in:
        m.set_input(0, src, 2)
        assert m.get_input(0) == (src, 2)
        with pytest.raises(srk.PortError):  # source port validated at connect (header note)
            m.set_input(0, src, 3)
        m.disconnect_inputs()
        assert srk.get_inputs(m) == [None] * n_in
        with pytest.raises(srk.SrackError) as e:
            m.set_input(0, m, 0) if n_out else (_ for _ in ()).throw(srk.SrackError(srk.STATUS["ERR_SELF_LOOP"]))
        assert e.value.status == srk.STATUS["ERR_SELF_LOOP"]
    assert len(m.get_id()) == 36 and m.get_id() != src.get_id()  # uuid v4 text


def test_parameter_defaults_follow_the_reference(srk):
    p = srk.Patch()
    assert p.add_module("Oscillator").get_param("OSC_VAL") == 0.0                 # oscillator.rs:32
    assert p.add_module("Oscillator").get_param("OSC_ANTIALIASING") == 1.0        # oscillator.rs:38
    a = p.add_module("ADSR")                                                       # adsr.rs:39-42
    assert [a.get_param(i) for i in range(4)] == [0.0, 0.5, 0.25, 0.5]
    f = p.add_module("Moog Filter")                                                # filter.rs:36-38
    assert [round(f.get_param(i), 6) for i in range(3)] == [0.2, 0.5, 0.5]
    assert [p.add_module("Mono Mixer").get_param(i) for i in range(4)] == [1.0] * 4  # mixer.rs:20
    assert p.add_module("Add").get_param(0) == 0.0                                 # math.rs:32
    assert p.add_module("Non-Linear").get_param(0) == 1.0                          # math.rs:194
    assert p.add_module("VCA").get_param("VCA_NEGATIVE") == 0.0                   # vca.rs:24
    with pytest.raises(srk.SrackError) as e:
        a.set_param(4, 1.0)
    assert e.value.status == srk.STATUS["ERR_PARAM"]
    with pytest.raises(srk.SrackError):
        p.add_module("Oscillator").set_param_per_voice("OSC_ANTIALIASING", [1.0, 0.0])  # uniform only


def test_set_audio_config_semantics(srk):
    p = srk.Patch(srk.AudioConfig(48000, 1024, 2))
    osc = p.add_module("Oscillator")
    out = p.add_module("Output")
    out.set_input(1, osc, 0)
    p.set_audio_config(srk.AudioConfig(44100, 256, 3))
    assert out.get_num_inputs() == 3            # output.rs:40-45: inputs rebuilt for `channels`...
    assert srk.get_inputs(out) == [None] * 3    # ...and dropped


def test_delete_module_disconnects_readers(srk):
    p = srk.Patch()
    osc = p.add_module("Oscillator")
    out = p.add_module("Output")
    out.set_input(0, osc, 0)
    p.delete_module(osc)
    assert srk.get_inputs(out) == [None, None]
    assert [m.get_name() for m in p.modules] == ["Output"]


def test_render_without_plan_or_device_fails_loudly(srk):
    import torch

    p = srk.Patch()
    srk.patches.cfg1(p, 1)
    with pytest.raises(srk.SrackError) as e:
        p.render(1, 16)
    assert e.value.status == srk.STATUS["ERR_NOT_PLANNED"]
    p.plan()
    if not torch.cuda.is_available():
        with pytest.raises(srk.SrackError) as e:
            p.render(1, 16)
        assert e.value.status == srk.STATUS["ERR_NO_DEVICE"]  # no CPU path


def test_kernel_id_names_the_image_a_render_would_use(srk):
    """srk_kernel_id (no device needed): equal patches -> equal ids, another graph or another launch shape -> another
    id; the interpreter kernels are named by the hash of their sources at build time."""
    def patch(builder, V):
        p = srk.Patch()
        builder(p, V)
        p.plan()
        return p
    a, b = patch(srk.patches.cfg2, 4096), patch(srk.patches.cfg2, 4096)
    ida = a.kernel_id(65536)
    assert ida == b.kernel_id(65536) and ida.startswith("fused:") and len(ida) > 12
    assert ida != patch(srk.patches.cfg4, 4096).kernel_id(65536)
    assert ida != a.kernel_id(4096)  # few voice groups per SM: the staged variant
    import os
    os.environ["SRK_FUSED"] = "0"
    try:
        idi = a.kernel_id(65536)
    finally:
        del os.environ["SRK_FUSED"]
    assert idi.startswith("interpreter:") and idi.endswith(":solo") and "unknown" not in idi


def test_state_blob_calls_fail_cleanly_without_a_render(srk):
    import ctypes as C
    p = srk.Patch()
    srk.patches.cfg2(p, 8)
    with pytest.raises(srk.SrackError):   # not planned
        p.state_import(b"SRKSTATE" + bytes(64))
    p.plan()
    with pytest.raises(srk.SrackError):   # nothing rendered
        p.state_export()
    with pytest.raises(srk.SrackError):   # not a blob
        p.state_import(b"nonsense")


def test_hostile_srk_file_is_refused_without_a_memory_blowup(srk):
    """nested array32 headers declaring huge counts used to allocate count x sizeof(Val) per level."""
    import resource
    import time
    p = srk.Patch()
    evil = (b"\xdd\x7f\xff\xff\xff" * 60) + b"\xc0" * (256 * 1024)
    before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    t0 = time.time()
    with pytest.raises(srk.SrackError):
        p.load_srk(evil)
    wide = b"\xdd\x00\x03\xff\xff" + b"\xc0" * (0x3ffff)  # one honest, wide array of nils: still not a patch
    with pytest.raises(srk.SrackError):
        p.load_srk(wide)
    assert time.time() - t0 < 5.0
    assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 200 * 1024  # KiB

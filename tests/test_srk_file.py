""".srk patch files (FileFormat src/ui.rs:578-586, serialize :98-114, deserialize
:115-134, the per-module serde derives).  The reference ships no .srk file and cannot be run here, so
the format is **unpinned by reference vectors**: what is checked is that two independent restatements
of rmp-serde's layout -- oracle/srk_file.py (schema-driven, from the struct definitions) and the
product's C++ codec behind srk_patch_load_srk / srk_patch_save_srk -- agree byte for byte, plus the
reference's load semantics (reversed module list, back-to-front connections, skipped bad entries).
The GPU render of a loaded file is in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import srk_file as sf

IDS = [f"{i:08x}-0000-4000-8000-{i:012x}" for i in range(1, 40)]


def subtractive_file(B=1024, state=True):
    """A cfg2-like patch as the reference GUI would have saved it mid-performance: non-trivial DSP
    state and port buffers in the file (which a load must ignore), V0 + V1 variants."""
    rng = np.random.default_rng(1)
    buf = lambda: [float(x) for x in rng.normal(0, 0.3, B).astype(np.float32)] if state else [0.0] * B
    lfo, osc, adsr, filt, vca, out, mul, noise, mix = IDS[:9]
    mods = [
        sf.new_module("OutputModuleV0", out, buffer_size=B),
        sf.new_module("OscillatorModuleV0", lfo, buffer_size=B, val=-7.78125, pos=0.3125 if state else 0.0, sine=buf()),
        sf.new_module("OscillatorModuleV0", osc, buffer_size=B, val=-1.25, antialiasing=False, saw=buf(),
                      sync_detector={"last": not state}),
        sf.new_module("ADSRModuleV0", adsr, buffer_size=B, a_sec=0.01, d_sec=0.1, s_val=0.5, r_sec=0.2,
                      phase=0.4 if state else 0.0, mode="Decay" if state else "None", r_val=0.1 if state else 0.0),
        sf.new_module("MoogFilterModuleV0", filt, buffer_size=B, freq=0.3, res=0.7, exp_amt=0.25,
                      state=dict(f=0.1, p=0.2, q=0.3, b=(0.1, 0.2, 0.3, 0.4, 0.5), freq=0.3, res=0.7)),
        sf.new_module("VCAModuleV0", vca, buffer_size=B, negative=True),
        sf.new_module("MathModuleV0", mul, buffer_size=B, constant=0.5, operation="Multiply"),
        sf.new_module("NoiseModuleV0", noise, buffer_size=B),
        sf.new_module("MonoMixerModuleV0", mix, buffer_size=B, gain=[1.0, 0.25, 0.0, 2.0]),
    ]
    conns = [(lfo, 1, adsr, 0), (osc, 2, mix, 0), (noise, 0, mul, 0), (mul, 0, mix, 1), (mix, 0, filt, 0), (adsr, 0, filt, 1),
             (filt, 0, vca, 0), (adsr, 0, vca, 1), (vca, 0, out, 0), (vca, 0, out, 1)]
    pos = [(m[1]["id"], (10.0 * i, -5.5 * i)) for i, m in enumerate(mods)]
    return dict(modules=mods, connections=conns, positions=pos)


def sequenced_file(B=512):
    clock, grid0, grid1, pat, osc, smp, nl, out, sub, add = IDS[10:20]
    cells1 = [None if i % 5 == 4 else (i * 3 % 25, i % 2 == 0) for i in range(16)]
    cells0 = [None if i % 3 == 2 else i * 2 for i in range(8)]
    rows = [[None if (r + s) % 3 == 0 else bool((r * s) % 2) for s in range(12)] for r in range(8)]
    wave = [float(x) for x in np.sin(np.arange(500) * 0.05).astype(np.float32)]
    mods = [
        sf.new_module("OscillatorModuleV0", clock, buffer_size=B, val=-4.0),
        sf.new_module("GridSequencerModuleV1", grid1, buffer_size=B, sequence=cells1, steps_per_octave=7, current_step=3, last=0.5),
        sf.new_module("GridSequencerModuleV0", grid0, buffer_size=B, sequence=cells0, octaves=3),
        sf.new_module("PatternSequencerModuleV0", pat, buffer_size=B, sequence=rows),
        sf.new_module("OscillatorModuleV0", osc, buffer_size=B, val=-1.0),
        sf.new_module("SampleModuleV0", smp, buffer_size=B, pos=17.5, playing=True,
                      wavebox=dict(samples=wave, sample_rate=22050.0, new=False)),
        sf.new_module("NonLinearModuleV0", nl, buffer_size=B, constant=0.75),
        sf.new_module("MathModuleV0", sub, buffer_size=B, constant=0.125, operation="Subtract"),
        sf.new_module("MathModuleV0", add, buffer_size=B, constant=-0.5, operation="Add"),
        sf.new_module("OutputModuleV0", out, buffer_size=B),
    ]
    conns = [(clock, 1, grid1, 0), (clock, 1, grid0, 0), (clock, 1, pat, 0), (grid1, 0, osc, 0), (grid0, 0, add, 0),
             (add, 0, smp, 1), (pat, 2, smp, 0), (osc, 2, nl, 0), (nl, 0, sub, 0), (smp, 0, sub, 1), (sub, 0, out, 0),
             (grid1, 1, out, 1)]
    return dict(modules=mods, connections=conns, positions=[])


def test_schema_round_trip_and_known_bytes():
    ff = subtractive_file(B=8)
    data = sf.dumps(ff)
    back = sf.loads(data)
    assert sf.dumps(back) == data
    assert [v for v, _ in back["modules"]] == [v for v, _ in ff["modules"]]
    # rmp-serde's layout, spelled out once by hand for a one-module file
    tiny = dict(modules=[sf.new_module("VCAModuleV0", "ab", buffer_size=2, negative=True)],
                connections=[("ab", 0, "cd", 200)], positions=[("ab", (1.0, -2.0))])
    want = (b"\x93"                                  # FileFormat: 3 fields
            b"\x91" b"\x81" b"\xabVCAModuleV0"        # modules: [ {"VCAModuleV0":
            b"\x93" b"\xa2ab"                         #   [id,
            b"\x92\xca\x00\x00\x00\x00\xca\x00\x00\x00\x00"  # buf: Some([0.0, 0.0])
            b"\xc3"                                   #   negative]
            b"\x91" b"\x94\xa2ab\x00\xa2cd\xcc\xc8"   # connections: [(src, 0, sink, 200)]
            b"\x91" b"\x92\xa2ab\x92\xca\x3f\x80\x00\x00\xca\xc0\x00\x00\x00")  # positions
    assert sf.dumps(tiny) == want


def test_load_follows_the_reference_semantics(srk):
    ff = subtractive_file()
    p = srk.Patch()
    p.add_module("Oscillator")  # replaced by the load
    assert p.load_srk(sf.dumps(ff)) == 0
    mods = p.modules
    # unpack_modules pops from the back: the list is the file's reversed (ui.rs:652-660)
    assert [m.get_id() for m in mods] == [m["id"] for _, m in reversed(ff["modules"])]
    assert [m.get_name() for m in mods] == ["Mono Mixer", "Noise", "Multiply", "VCA", "Moog Filter", "ADSR", "Oscillator",
                                            "Oscillator", "Output"]
    by = {m.get_id(): m for m in mods}
    lfo, osc, adsr, filt, vca, out, mul, noise, mix = (by[i] for i in IDS[:9])
    assert lfo.get_param("OSC_VAL") == -7.78125 and osc.get_param("OSC_ANTIALIASING") == 0.0
    assert [adsr.get_param(i) for i in range(4)] == [np.float32(x) for x in (0.01, 0.1, 0.5, 0.2)]
    assert [filt.get_param(i) for i in range(3)] == [np.float32(x) for x in (0.3, 0.7, 0.25)]
    assert vca.get_param("VCA_NEGATIVE") == 1.0 and mul.get_param(0) == 0.5
    assert [mix.get_param(i) for i in range(4)] == [1.0, 0.25, 0.0, 2.0]
    assert srk.get_inputs(filt) == [(mix, 0), (adsr, 0)] and srk.get_inputs(out) == [(vca, 0), (vca, 0)]
    assert srk.get_inputs(mix) == [(osc, 2), (mul, 0), None, None] and srk.get_inputs(adsr) == [(lfo, 1)]
    plan = p.plan()
    assert plan[-1] == out and len(plan) == 9
    # a second load replaces everything again; unknown ids, bad ports and duplicates
    ff2 = sequenced_file()
    ff2["connections"] += [("nobody", 0, IDS[17], 0), (IDS[14], 0, "nobody", 0), (IDS[14], 9, IDS[17], 1), (IDS[14], 0, IDS[17], 7),
                           (IDS[14], 0, IDS[14], 0)]
    ff2["connections"].insert(0, (IDS[14], 1, IDS[17], 1))   # earlier entry for out.1: applied LAST, wins (ui.rs:673)
    assert p.load_srk(sf.dumps(ff2)) == 5
    by = {m.get_id(): m for m in p.modules}
    assert len(by) == 10 and by[IDS[17]].get_input(1) == (by[IDS[14]], 1)
    g0, g1, pat, smp = by[IDS[11]], by[IDS[12]], by[IDS[13]], by[IDS[15]]
    assert g1.get_param("GRIDSEQ_STEPS_PER_OCTAVE") == 7.0
    assert g1.get_sequence().tolist() == [-1 if c is None else (c[0] | (0x10000 if c[1] else 0)) for c in ff2["modules"][1][1]["sequence"]]
    assert g0.get_sequence().tolist() == [-1 if c is None else c for c in ff2["modules"][2][1]["sequence"]]  # V0: hold = false
    assert pat.get_sequence().shape == (8, 12) and pat.get_sequence()[1, 1] == 1 and pat.get_sequence()[0, 0] == -1
    wave, rate = smp.get_sample()
    assert rate == 22050.0 and (wave == np.array(ff2["modules"][5][1]["wavebox"]["samples"], np.float32)).all()
    assert by[IDS[18]].get_name() == "Subtract" and by[IDS[19]].get_name() == "Add" and by[IDS[16]].get_param(0) == 0.75


def test_save_matches_the_schema_encoder_byte_for_byte(srk):
    """Product writer vs the schema-driven restatement: same bytes for the same freshly built patch."""
    for make, B in ((subtractive_file, 64), (sequenced_file, 32)):
        ff = make(B=B)
        p = srk.Patch(srk.AudioConfig(48000, B, 2))
        p.load_srk(sf.dumps(ff))
        saved = p.save_srk()
        got = sf.loads(saved)  # strict decode: layout is exactly rmp-serde's
        # expected: the loaded modules with zeroed port buffers and the state they came with, list order = reversed file order
        want_mods = []
        for variant, m in reversed(ff["modules"]):
            over = {k: v for k, v in m.items() if k in ("val", "antialiasing", "a_sec", "d_sec", "s_val", "r_sec", "negative",
                                                        "freq", "res", "exp_amt", "gain", "constant", "operation",
                                                        "steps_per_octave",
                                                        # the DSP state a loaded module carries is written back
                                                        "pos", "sync_detector", "phase", "mode", "r_val", "from_a_val",
                                                        "transition_detector", "sync_transition_detector", "state",
                                                        "current_step", "last", "playing")}
            v_out = {"GridSequencerModuleV0": "GridSequencerModuleV1", "MoogFilterModuleV0": "MoogFilterModuleV1"}.get(variant, variant)
            if "Sequencer" in variant:
                seq = m["sequence"]
                over["sequence"] = [None if c is None else (c, False) for c in seq] if variant == "GridSequencerModuleV0" else seq
            if variant == "SampleModuleV0":
                over["wavebox"] = dict(samples=m["wavebox"]["samples"], sample_rate=m["wavebox"]["sample_rate"], new=False)
            want_mods.append(sf.new_module(v_out, m["id"], buffer_size=B, **over))
        ids = [m["id"] for _, m in want_mods]
        n_in = lambda mid: {v: k for k, v in enumerate(ids)}[mid]
        want_conns = sorted(ff["connections"], key=lambda c: (n_in(c[2]), c[3]))  # capture_connections: per module, per input
        want = dict(modules=want_mods, connections=want_conns, positions=ff["positions"])
        assert got["connections"] == want["connections"]
        assert got["positions"] == [(i, (np.float32(x), np.float32(y))) for i, (x, y) in want["positions"]]
        assert saved == sf.dumps(want)
        # and a save -> load -> save cycle is a fixed point up to the list reversal
        q = srk.Patch(srk.AudioConfig(48000, B, 2))
        q.load_srk(saved)
        r = srk.Patch(srk.AudioConfig(48000, B, 2))
        r.load_srk(q.save_srk())
        assert r.save_srk() == saved


def test_lenient_reader_and_errors(srk):
    p = srk.Patch()
    keep = p.add_module("VCA")
    good = sf.dumps(subtractive_file(B=4))
    for bad in (b"", b"\x93", good[:-3], good[:200], b"\x93\x90\x90", b"\xc1", good.replace(b"VCAModuleV0", b"VCAModuleV9")):
        with pytest.raises(srk.SrackError) as e:
            p.load_srk(bad)
        assert e.value.status == srk.STATUS["ERR_ARG"], bad[:8]
        assert p.modules == [keep]  # a refused file leaves the patch alone
    fv = dict(modules=[sf.new_module("FreeverbModuleV0", "x", buffer_size=4)], connections=[], positions=[])
    with pytest.raises(srk.SrackError) as e:
        p.load_srk(sf.dumps(fv))
    assert e.value.status == srk.STATUS["ERR_UNSUPPORTED"] and p.modules == [keep]
    # struct-as-map and variant-by-index encodings (other rmp-serde configurations) are accepted
    idx_variant = b"\x93\x91\x81\x07\x93\xa2ab\xc0\xc3\x90\x90"  # {7: [id, buf = None, negative = true]}
    assert p.load_srk(idx_variant) == 0 and p.modules[0].get_name() == "VCA" and p.modules[0].get_param(0) == 1.0
    as_map = (b"\x83\xa7modules\x91\x81\xabVCAModuleV0\x83\xa2id\xa2zz\xa3buf\xc0\xa8negative\xc2"
              b"\xabconnections\x90\xa9positions\x90")
    assert p.load_srk(as_map) == 0 and p.modules[0].get_id() == "zz" and p.modules[0].get_param(0) == 0.0


def test_oracle_loads_the_same_graph(srk, orc):
    """oracle/srk_file.build drives an OraclePatch to the same plan as the product's load."""
    for make in (subtractive_file, sequenced_file):
        ff = make(B=256)
        data = sf.dumps(ff)
        p = srk.Patch(srk.AudioConfig(48000, 256, 2))
        p.load_srk(data)
        plan_ids = [m.get_id() for m in p.plan()]
        op = orc.OraclePatch(48000, 256, 2)
        handles = op.load_srk(data)
        rev = {h: i for i, h in handles.items()}
        oplan, _ = op.plan()
        assert [rev[h] for h in oplan] == plan_ids


def test_committed_srk_fixture(srk):
    """tests/golden/subtractive_b16.srk (tests/golden/make_srk_golden.py): both codecs still read the bytes the
    way they did when the fixture was written, state included."""
    import os
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "subtractive_b16.srk"), "rb").read()
    assert data == sf.dumps(subtractive_file(B=16))          # the encoder has not drifted
    ff = sf.loads(data)
    p = srk.Patch(srk.AudioConfig(48000, 16, 2))
    assert p.load_srk(data) == 0
    back = sf.loads(p.save_srk())
    by_id = {m["id"]: (v, m) for v, m in back["modules"]}
    for variant, m in ff["modules"]:
        v2, m2 = by_id[m["id"]]
        assert v2 == {"MoogFilterModuleV0": "MoogFilterModuleV1"}.get(variant, variant)
        for key in ("val", "pos", "sync_detector", "antialiasing", "a_sec", "d_sec", "s_val", "r_sec", "phase", "mode", "r_val",
                    "from_a_val", "transition_detector", "freq", "res", "exp_amt", "state", "negative", "constant",
                    "operation", "gain"):
            if key in m:
                assert m2[key] == m[key], (variant, key)
    adsr = by_id[IDS[2]][1]
    assert adsr["mode"] == "Decay" and abs(adsr["phase"] - 0.4) < 1e-7 and by_id[IDS[0]][1]["pos"] == 0.3125

"""Grid / Pattern sequencers (src/synth/sequencer.rs:190-246, :482-533).
CPU part: the oracle against an independent per-sample restatement written here from the
reference source, and the C ABI's table calls.  The GPU part lives in test_gpu_parity.py."""
import numpy as np
import pytest

SEQ_NONE = -1


def cell(val, hold):
    return (val & 0xFFFF) | (0x10000 if hold else 0)


def grid_reference(step_in, sync_in, cells, steps_per_octave):
    """sequencer.rs:213-241, sample by sample; detectors start with last = true (synth.rs:283)."""
    n = len(step_in)
    cv, gate, sync = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    cur, last_step, last_sync, last = 0, True, True, np.float32(0)
    inv = np.float32(1.0) / np.float32(steps_per_octave)
    for i in range(n):
        a = step_in[i] > 0
        if a and not last_step:
            cur += 1
        last_step = a
        b = sync_in[i] > 0
        if b and not last_sync:
            cur = 0
        last_sync = b
        if cur >= len(cells):
            cur = 0
        c = cells[cur]
        if c >= 0:
            cv[i] = np.float32(c & 0xFFFF) * inv
            gate[i] = 1.0 if (c >> 16) & 1 else step_in[i]
        else:
            cv[i], gate[i] = last, 0.0
        sync[i] = 1.0 if cur == 0 else 0.0
        last = cv[i]
    return cv, gate, sync


def pattern_reference(step_in, sync_in, rows):
    n, steps = len(step_in), rows.shape[1]
    outs = np.zeros((9, n), np.float32)
    cur, last_step, last_sync = 0, True, True
    for i in range(n):
        a = step_in[i] > 0
        if a and not last_step:
            cur += 1
        last_step = a
        b = sync_in[i] > 0
        if b and not last_sync:
            cur = 0
        last_sync = b
        if cur >= steps:
            cur = 0
        for r in range(8):
            c = rows[r, cur]
            outs[r, i] = 0.0 if c < 0 else (1.0 if c else step_in[i])
        outs[8, i] = 1.0 if cur == 0 else 0.0
    return outs


def _clocked(orc, kind, n_out, B, clock_hz, sync_hz):
    """clock oscillator (square) -> sequencer.step, slower oscillator (square) -> sequencer.sync; the
    sequencer's ports, the clock and the sync pulse on the Output's channels."""
    from srack_b200.patches import hz_to_val
    op = orc.OraclePatch(48000, B, n_out + 2)
    clock = op.module_create("OSCILLATOR")
    syncer = op.module_create("OSCILLATOR")
    seq = op.module_create(kind)
    out = op.module_create("OUTPUT")
    op.set_param(clock, 0, hz_to_val(clock_hz))
    op.set_param(syncer, 0, hz_to_val(sync_hz))
    op.connect(seq, 0, clock, 1)
    op.connect(seq, 1, syncer, 1)
    for port in range(n_out):
        op.connect(out, port, seq, port)
    op.connect(out, n_out, clock, 1)
    op.connect(out, n_out + 1, syncer, 1)
    return op, seq


@pytest.mark.parametrize("B", [64, 1024])
def test_oracle_grid_sequencer_matches_the_restatement(srk, orc, B):
    rng = np.random.default_rng(7)
    for n_steps, spo in ((64, 12), (5, 7), (1, 12)):
        op, seq = _clocked(orc, "GRID_SEQUENCER", 3, B, 400.0, 37.0)
        cells = np.array([SEQ_NONE if rng.random() < 0.3 else cell(int(rng.integers(0, 30)), bool(rng.random() < 0.5))
                          for _ in range(n_steps)], dtype=np.int32)
        op.set_sequence(seq, cells)
        op.set_param(seq, 0, spo)
        st, _ = op.render(1, 6000)
        cv, gate, sync = grid_reference(st[3, :, 0], st[4, :, 0], cells, spo)
        assert (st[0, :, 0] == cv).all() and (st[1, :, 0] == gate).all() and (st[2, :, 0] == sync).all()
        assert len(np.unique(cv)) > 1 or n_steps == 1


def test_oracle_grid_sequencer_defaults(srk, orc):
    """vec![None; 64], steps_per_octave 12 (sequencer.rs:40,44): cv holds `last` = 0, gate 0, sync 1 at step 0."""
    op, seq = _clocked(orc, "GRID_SEQUENCER", 3, 256, 1000.0, 0.001)
    st, _ = op.render(1, 4800)
    assert (st[0] == 0).all() and (st[1] == 0).all()
    cv, gate, sync = grid_reference(st[3, :, 0], st[4, :, 0], [SEQ_NONE] * 64, 12)
    assert (st[2, :, 0] == sync).all() and 0 < sync.sum() < len(sync)


def test_oracle_pattern_sequencer_matches_the_restatement(srk, orc):
    rng = np.random.default_rng(11)
    for n_steps in (64, 16, 3):
        op, seq = _clocked(orc, "PATTERN_SEQUENCER", 9, 512, 600.0, 29.0)
        rows = rng.integers(-1, 2, size=(8, n_steps)).astype(np.int32)
        op.set_sequence(seq, rows)
        st, _ = op.render(1, 5000)
        ref = pattern_reference(st[9, :, 0], st[10, :, 0], rows)
        assert (st[:9, :, 0] == ref).all()


def test_sequence_table_through_the_abi(srk):
    p = srk.Patch()
    grid = p.add_module("Grid Sequencer")
    pat = p.add_module("Pattern Sequencer")
    osc = p.add_module("Oscillator")
    out = p.add_module("Output")
    assert (grid.get_sequence() == SEQ_NONE).all() and grid.get_sequence().shape == (64,)
    assert (pat.get_sequence() == SEQ_NONE).all() and pat.get_sequence().shape == (8, 64)
    assert grid.get_param("GRIDSEQ_STEPS_PER_OCTAVE") == 12.0
    cells = np.array([srk.grid_cell(3, True), SEQ_NONE, srk.grid_cell(15, False)], dtype=np.int32)
    grid.set_sequence(cells)
    assert (grid.get_sequence() == cells).all()
    rows = np.arange(8 * 5).reshape(8, 5) % 3 - 1
    pat.set_sequence(rows)
    assert (pat.get_sequence() == rows).all()
    for bad in (np.zeros(0, np.int32), np.zeros(65, np.int32), np.array([-2], np.int32), np.array([1 << 17], np.int32)):
        with pytest.raises(srk.SrackError) as e:
            grid.set_sequence(bad)
        assert e.value.status == srk.STATUS["ERR_ARG"]
    with pytest.raises(srk.SrackError):
        pat.set_sequence(np.full((8, 4), 2, np.int32))
    with pytest.raises(srk.SrackError) as e:
        osc.set_sequence(cells)
    assert e.value.status == srk.STATUS["ERR_KIND"]
    with pytest.raises(srk.SrackError) as e:
        grid.set_param_per_voice("GRIDSEQ_STEPS_PER_OCTAVE", np.ones(4, np.float32))
    assert e.value.status == srk.STATUS["ERR_PARAM"]
    # editing a table does not invalidate the plan (the reference's ui() edits it under the running module)
    out.set_input(0, grid, 0)
    out.set_input(1, pat, 8)
    p.plan()
    before = p.program(64)
    grid.set_sequence(cells[:2])
    after = p.program(64)  # would raise NOT_PLANNED
    assert [i["op"] for i in before[0]] == [i["op"] for i in after[0]]
    assert [i["n_ch"] for i in after[0] if i["op"] == "GRIDSEQ"] == [2]
    # the pattern sequencer's 9 ports become 3-port instructions; only triples somebody reads exist
    assert [(i["op"], i["flags"]) for i in after[0] if i["op"] == "PATSEQ"] == [("PATSEQ", 6)]

"""The product's planner (C ABI, no GPU needed) against the oracle's literal restatement of
plan_execution (src/synth.rs:107-212): the reference's own `topological_sort` test
(src/synth.rs:537-613) through the ABI, then exact plan-order and cut-wire equality on random
cyclic graphs and random module-list orders."""
import random

import pytest

from test_oracle_kat import _wire_topological_sort_graph


def test_topological_sort_through_the_abi(srk):
    p = srk.Patch(srk.AudioConfig(44100, 64, 2))
    mods, out = _wire_topological_sort_graph(p)
    rng = random.Random(99)
    for _ in range(1000):
        order = mods + [out]
        rng.shuffle(order)
        p.set_module_order(order)
        plan = srk.plan_execution(p)
        idx = {m: i for i, m in enumerate(plan)}
        assert len(plan) == 8
        assert idx[mods[0]] < idx[mods[1]] < idx[mods[2]] < idx[mods[3]] < idx[out]
        assert idx[mods[0]] < idx[mods[4]] < idx[mods[3]]
        assert idx[mods[6]] < idx[mods[4]]
        assert idx[mods[5]] < idx[mods[6]]
        assert p.plan_cuts() == [(mods[5], mods[6])]


KINDS = ["OSCILLATOR", "NOISE", "ADSR", "VCA", "MOOG_FILTER", "MONO_MIXER", "ADD", "SUBTRACT", "MULTIPLY",
         "NON_LINEAR", "GRID_SEQUENCER", "PATTERN_SEQUENCER"]
N_IN = dict(OSCILLATOR=2, NOISE=0, ADSR=1, VCA=2, MOOG_FILTER=2, MONO_MIXER=4, ADD=2, SUBTRACT=2, MULTIPLY=2,
            NON_LINEAR=2, GRID_SEQUENCER=2, PATTERN_SEQUENCER=2, OUTPUT=2)
N_OUT = dict(OSCILLATOR=3, NOISE=1, ADSR=1, VCA=1, MOOG_FILTER=3, MONO_MIXER=1, ADD=1, SUBTRACT=1, MULTIPLY=1,
             NON_LINEAR=1, GRID_SEQUENCER=3, PATTERN_SEQUENCER=9, OUTPUT=0)


def random_graph(rng, n_modules, density):
    kinds = [rng.choice(KINDS) for _ in range(n_modules)]
    kinds.insert(rng.randrange(n_modules + 1), "OUTPUT")
    if rng.random() < 0.2:
        kinds.append("OUTPUT")  # a second Output: only the first one in list order is "the" output
    wires = []
    for sink, k in enumerate(kinds):
        for i in range(N_IN[k]):
            if rng.random() < density:
                src = rng.randrange(len(kinds))
                if src != sink and N_OUT[kinds[src]] > 0:
                    wires.append((sink, i, src, rng.randrange(N_OUT[kinds[src]])))
    return kinds, wires


@pytest.mark.parametrize("seed", range(40))
def test_random_graphs_match_the_oracle_planner(srk, orc, seed):
    rng = random.Random(seed)
    kinds, wires = random_graph(rng, rng.randrange(2, 14), rng.choice([0.3, 0.6, 0.9]))
    gp = srk.Patch()
    op = orc.OraclePatch()
    gm = [gp.module_create(k) for k in kinds]
    om = [op.module_create(k) for k in kinds]
    for sink, i, src, port in wires:
        gp.connect(gm[sink], i, gm[src], port)
        op.connect(om[sink], i, om[src], port)
    for _ in range(6):
        order = list(range(len(kinds)))
        rng.shuffle(order)
        gp.set_module_order([gm[i] for i in order])
        op.set_module_order([om[i] for i in order])
        g_plan = [gm.index(m) for m in gp.plan()]
        o_plan, o_cuts = op.plan()
        assert g_plan == o_plan
        assert [(gm.index(r), gm.index(w)) for r, w in gp.plan_cuts()] == o_cuts
        assert sorted(g_plan) == list(range(len(kinds)))  # every module is planned (synth.rs:193-211)


def test_no_output_gives_an_empty_plan(srk):
    p = srk.Patch()
    p.module_create("OSCILLATOR")
    with pytest.raises(srk.SrackError) as e:
        p.plan()  # ui.rs:76-80: plan cleared when there is no Output
    assert e.value.status == srk.STATUS["ERR_NO_OUTPUT"]


def test_compiled_program_marks_delayed_wires(srk, monkeypatch):
    """A wire whose source runs after its reader becomes a ring (one-block delay)."""
    monkeypatch.setenv("SRK_FUSED", "0")  # the interpreter's chunk length is what is checked at the end
    p = srk.Patch()
    srk.patches.cfg3b(p, 4)
    p.plan()
    info = p.program_info(4)
    assert info["n_rings"] == 1
    q = srk.Patch()
    srk.patches.cfg3(q, 4)
    q.plan()
    assert q.program_info(4)["n_rings"] == 0
    # step never exceeds buffer_size when a ring exists (reader must not outrun the writer)
    r = srk.Patch(srk.AudioConfig(48000, 4, 2))
    srk.patches.cfg3b(r, 4)
    r.plan()
    assert r.program_info(4)["step_samples"] == 4

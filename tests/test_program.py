"""The patch compiler's schedule, checked on the CPU through srk_get_program (no device):
the invariants the voice kernel relies on (program.hpp) for the BASELINE patches and for random
(often cyclic) graphs, under both schedules -- pipelined warps (few voices) and one warp per
voice group (many voices)."""
import random

import pytest

from test_gpu_parity import _fuzz_patch

PIPELINED_V, SOLO_V = 4096, 65536


@pytest.fixture(autouse=True)
def interpreter_schedules(monkeypatch):
    """This module checks the two INTERPRETER schedules (pipelined warps / one warp per group); the fused kernel
    that replaces the one-warp schedule by default has its own tests (tests/test_fused.py)."""
    monkeypatch.setenv("SRK_FUSED", "0")


def _build(srk, builder, B=1024, channels=2):
    p = srk.Patch(srk.AudioConfig(48000, B, channels))
    builder(p, 8)
    p.plan()
    return p


def _check_program(srk, p, n_voices, B):
    info = p.program_info(n_voices)
    instrs, wires = p.program(n_voices)
    assert instrs[-1]["op"] == "END" and all(i["op"] != "END" for i in instrs[:-1])
    code = instrs[:-1]
    assert len(instrs) == info["n_instr"] and len(wires) == info["n_wires"]
    # sorted by warp, stages and warps in range
    assert [i["warp"] for i in code] == sorted(i["warp"] for i in code)
    assert all(i["warp"] < info["n_warps"] and i["stage"] < info["n_stages"] for i in code)
    assert info["block_threads"] == 32 * info["n_warps"] * info["groups_per_block"] <= 512
    assert info["groups_per_block"] == 1 or info["n_warps"] == 1  # only the one-warp schedule shares a block
    # wire rings tile the group's tile array without overlap; ring lengths are powers of two
    covered = sorted((first, first + n) for first, n in wires)
    assert all(n & (n - 1) == 0 and n >= 1 for _, n in wires)
    assert covered[0][0] == 0 if covered else True
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert (covered[-1][1] if covered else 0) == info["n_tiles"]
    pipelined = info["n_warps"] > 1
    for i in code:
        slots_in = [s for s in i["ins"] if s >= 0]
        slots_out = [s for s in i["outs"] if s >= 0]
        assert all(s < len(wires) for s in slots_in + slots_out)
        assert not set(slots_in) & set(slots_out), "an instruction never works in place"
    if pipelined:
        assert info["step_samples"] >= 8
        producers, consumers = {}, {}
        for i in code:
            for s in i["outs"]:
                if s >= 0:
                    producers.setdefault(s, []).append(i)
            for s in i["ins"]:
                if s >= 0:
                    consumers.setdefault(s, []).append(i)
        for s, (first, n_tiles) in enumerate(wires):
            ps, cs = producers.get(s, []), consumers.get(s, [])
            assert ps and cs, "every materialised wire is written and read"
            assert len({q["stage"] for q in ps}) == 1
            if len(ps) > 1:  # only time-split copies (oscillators, their V/oct conversions) share a wire
                n = ps[0]["flags"] >> 4
                assert len({q["op"] for q in ps}) == 1 and ps[0]["op"] in ("OSC", "OSC_DELTA")
                assert all(q["flags"] >> 4 == n and (q["ins"][0] < 0) == (q["op"] == "OSC") for q in ps) and n == len(ps)
                assert sorted(q["flags"] & 15 for q in ps) == list(range(n))
                assert len({(q["state"], q["param"]) for q in ps}) == 1
                assert len({q["warp"] for q in ps}) == n, "copies on one warp would be pointless"
            # a reader runs strictly later than the writer, and the ring holds every chunk in between
            delta = max(c["stage"] for c in cs) - ps[0]["stage"]
            assert min(c["stage"] for c in cs) > ps[0]["stage"]
            assert n_tiles >= delta + 1
        if len(code) <= 16:
            assert len({i["warp"] for i in code}) == len(code), "one instruction per warp while warps last"
        # the slowest stage has SM sub-partition 0 (warps 0, 4, 8, 12) to itself when warps allow
        if len(code) - 1 <= 12 and len(code) >= 2:
            assert sum(1 for i in code if i["warp"] % 4 == 0) == 1
        stores = [i["stage"] for i in code if i["op"] == "RING_STORE"]
        if stores:
            assert info["step_samples"] * (max(stores) + 2) <= B
    else:
        assert all(i["warp"] == 0 and i["stage"] == 0 for i in code)
        assert all(n == 1 for _, n in wires)
        assert all(i["flags"] >> 4 == 0 for i in code if i["op"] == "OSC"), "no time-split copies on one warp"
        written = set()
        for i in code:  # plan order: nothing is read before something wrote it
            assert all(s in written for s in i["ins"] if s >= 0), i
            written.update(s for s in i["outs"] if s >= 0)
        if info["n_rings"]:
            assert info["step_samples"] <= B
    return info, code


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg3b", "cfg4"])
@pytest.mark.parametrize("n_voices", [PIPELINED_V, SOLO_V])
def test_baseline_patch_schedules(srk, name, n_voices):
    p = _build(srk, srk.patches.CONFIGS[name][0])
    info, code = _check_program(srk, p, n_voices, 1024)
    assert (info["n_warps"] > 1) == (n_voices == PIPELINED_V)
    ops = [i["op"] for i in code]
    assert ops.count("OUTPUT") == 1 and ops.count("MIX") == 1
    assert ops.count("RING_LOAD") == ops.count("RING_STORE") == (1 if name == "cfg3b" else 0)


def test_oscillators_are_time_split_only_when_pipelined(srk):
    p = _build(srk, srk.patches.cfg2)
    _, code = _check_program(srk, p, PIPELINED_V, 1024)
    oscs = [i for i in code if i["op"] == "OSC"]
    assert len(oscs) == 8 and {i["flags"] >> 4 for i in oscs} == {4}  # two oscillators x 4 copies
    assert [i["warp"] for i in code if i["op"] == "MOOG"] == [0]        # the critical stage, alone on sub-partition 0
    _, code = _check_program(srk, p, SOLO_V, 1024)
    assert len([i for i in code if i["op"] == "OSC"]) == 2
    # a CV-driven oscillator first loses its V/oct conversion to an OSC_DELTA stage (a wire pair carries
    # the f64 delta); then both are time-split like a CV-less oscillator
    q = _build(srk, srk.patches.cfg3)
    _, code = _check_program(srk, q, PIPELINED_V, 1024)
    deltas = [i for i in code if i["op"] == "OSC_DELTA"]
    assert deltas and all(i["ins"][0] >= 0 and i["outs"][0] >= 0 and i["outs"][1] >= 0 for i in deltas)
    fm = [i for i in code if i["op"] == "OSC" and i["n_ch"] == 1]
    assert fm and all(i["ins"][0] < 0 and i["ins"][2] == deltas[0]["outs"][0] and i["ins"][3] == deltas[0]["outs"][1]
                      for i in fm)
    assert not any(i["op"] == "OSC" and i["ins"][0] >= 0 for i in code)
    _, code = _check_program(srk, q, SOLO_V, 1024)
    assert not any(i["op"] == "OSC_DELTA" for i in code) and any(i["op"] == "OSC" and i["ins"][0] >= 0 for i in code)


@pytest.mark.parametrize("B", [1, 4, 17, 64, 1024])
def test_feedback_patch_respects_the_ring_latency(srk, B):
    p = _build(srk, srk.patches.cfg3b, B=B)
    for n_voices in (PIPELINED_V, SOLO_V):
        info, _ = _check_program(srk, p, n_voices, B)
        assert info["step_samples"] <= B
        if B < 8 * 3:
            assert info["n_warps"] == 1  # too short a block to pipeline across


@pytest.mark.parametrize("seed", range(40))
def test_random_patch_schedules(srk, seed):
    rng = random.Random(5000 + seed)
    kinds, wires = _fuzz_patch(rng, rng.randrange(3, 14))
    B = rng.choice([1, 5, 64, 1024])
    p = srk.Patch(srk.AudioConfig(48000, B, 2))
    mods = [p.module_create(k) for k in kinds]
    for sink, i, src, port in wires:
        p.connect(mods[sink], i, mods[src], port)
    p.plan()
    for n_voices in (1, 33, PIPELINED_V, SOLO_V):
        _check_program(srk, p, n_voices, B)


def test_wide_output_is_chunked_by_four_channels(srk):
    def build(b, n):
        osc = b.module_create("OSCILLATOR")
        out = b.module_create("OUTPUT")
        for c in range(6):
            b.connect(out, c, osc, c % 3)
    p = _build(srk, build, channels=6)
    _, code = _check_program(srk, p, PIPELINED_V, 1024)
    outs = [i for i in code if i["op"] == "OUTPUT"]
    assert [(i["aux"], i["n_ch"]) for i in outs] == [(0, 4), (4, 2)]
    assert [(i["aux"], i["n_ch"]) for i in code if i["op"] == "MIX"] == [(0, 4), (4, 2)]


def test_one_warp_schedule_block_shape(srk):
    """One-warp schedule (no device needed: the sm_100 limits are assumed): a block holds all the groups its SM gets,
    the chunk is the longest that still fits 227 KB of shared memory, and never shorter than 16 samples."""
    limit, n_sm = 227 * 1024, 148
    for name, V, want in (("cfg2", 65536, (14, 32)), ("cfg2", 32768, (7, 64)), ("cfg3", 65536, (14, 32)),
                          ("cfg4", 32768, (7, 32)), ("cfg1", 65536, (14, 64)), ("cfg2", 262144, (14, 32))):
        p = srk.Patch()
        srk.patches.CONFIGS[name][0](p, 8)
        p.plan()
        info = p.program_info(V)
        assert info["n_warps"] == 1 and info["n_stages"] == 1, (name, V)
        assert (info["groups_per_block"], info["step_samples"]) == want, (name, V, info)
        assert info["smem_bytes"] <= limit and info["block_threads"] == 32 * info["groups_per_block"]
        groups = -(-V // 32)
        per_sm = -(-groups // n_sm)
        blocks_per_sm = -(-per_sm // 16)
        assert info["groups_per_block"] == -(-per_sm // blocks_per_sm)
        # the next longer chunk would not have fitted with that many groups
        if info["step_samples"] < 128:
            per_group = (info["smem_bytes"] - 2048) // info["groups_per_block"]  # minus (at most) the program image
            tiles = info["n_tiles"] * 32 * 4
            assert info["smem_bytes"] + info["groups_per_block"] * tiles * info["step_samples"] > limit, (name, V, per_group)

"""The patch specialiser (csrc/fused_gen.cpp -> generated CUDA C++ -> NVRTC for sm_100a -> cubin cache), checked on the CPU:
NVRTC needs no GPU, so "every BASELINE graph's kernels -- the cost model's choice and the alternatives a long render
measures against it -- compile for sm_100a, carry the TMA store, and do not spill" is a CPU test.  The arithmetic of the
generated kernels is checked on the GPU (tests/test_gpu_parity.py runs every test under three schedules); the one GPU test
here keeps a compiler mis-fold the op templates had to work around."""
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from util import assert_parity, build_both

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _builders(srk):
    b = {n: c[0] for n, c in srk.patches.CONFIGS.items()}
    b.update({g.__name__: g for g in srk.patches.CFG5_GRAPHS})
    b["sequenced"] = srk.patches.sequenced
    b["sampler"] = srk.patches.sampler
    return b


def _patch(srk, builder, V, B=1024):
    p = srk.Patch(srk.AudioConfig(48000, B, 2))
    builder(p, V)
    p.plan()
    return p


def _ops(src):
    return re.findall(r"^\s+(Osc|Noise|Moog|MoogCoef|MoogCore|Adsr|Vca|Mixer|Math|GridSeq|PatSeq|Sample)[< ]", src, re.M)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg3b", "cfg4", "cfg5_bandpass", "cfg5_two_osc", "cfg5_no_noise",
                                  "cfg5_gated_sine", "sequenced", "sampler"])
def test_generated_source_follows_the_launch_shape(srk, name):
    builder = _builders(srk)[name]
    full = _patch(srk, builder, 65536)
    few = _patch(srk, builder, 4096)
    src_full, src_few = full.fused_source(65536), few.fused_source(4096)
    assert src_full and '#include "fused_ops.cuh"' in src_full and "srk_fused_kernel" in src_full
    info_full = full.program_info(65536)
    assert info_full["fused"] == 1 and info_full["n_warps"] == 1 and "switch (c.stage)" not in src_full  # a full chip: one warp per group
    assert "MoogCoef" not in src_full  # the ladder filter is split only where stages can separate the halves
    # the same modules whatever the shape (a staged kernel may split a CV-driven ladder filter in two ops)
    norm = lambda ops: sorted("Moog" if o == "MoogCore" else o for o in ops if o != "MoogCoef")  # noqa: E731
    if src_few:  # (an interpreter pipeline may be the cost model's choice at few voices: then there is no fused source)
        info = few.program_info(4096)
        assert norm(_ops(src_few)) == norm(_ops(src_full))
        assert 1 <= info["n_warps"] <= 8 and (info["n_warps"] > 1) == ("switch (c.stage)" in src_few)
        assert _ops(src_few).count("MoogCoef") == _ops(src_few).count("MoogCore")
        assert info["smem_bytes"] <= 227 * 1024
    # equal patches -> equal source -> equal kernel id; another voice-parameter uniformity -> another kernel
    again = _patch(srk, builder, 65536)
    assert again.fused_source(65536) == src_full and again.kernel_id(65536) == full.kernel_id(65536)


def test_every_baseline_kernel_compiles_for_sm_100a_with_tma_and_without_spills(srk, tmp_path, monkeypatch):
    """srk_precompile: the model's choice and the measured alternatives, through NVRTC, into a cubin cache."""
    monkeypatch.setenv("SRK_KERNEL_CACHE", str(tmp_path))
    try:
        n = 0
        for name, V in (("cfg2", 4096), ("cfg2", 65536), ("cfg4", 32768), ("cfg3b", 65536), ("sampler", 4096)):
            n += _patch(srk, _builders(srk)[name], V).precompile(V)
    except srk.SrackError as e:
        pytest.skip(f"no NVRTC on this machine: {e}")
    cubins = sorted(glob.glob(str(tmp_path / "*.cubin")))
    assert n == len(cubins) and n >= 10  # cfg2 @ 4096 alone has eight fused candidates
    # the kernels the cost model launches before anything is measured: no spills at all
    first = {_patch(srk, _builders(srk)[name], V).kernel_id(V).split(":")[1]
             for name, V in (("cfg2", 4096), ("cfg2", 65536), ("cfg4", 32768), ("cfg3b", 65536), ("sampler", 4096))}
    for f in cubins:
        sass = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", f], capture_output=True, text=True).stdout
        assert "sm_100a" in sass and "UTMASTG" in sass and "UTMACMDFLUSH" in sass, f  # TMA bulk tensor stores of the stems tiles
        regs, stack = int(re.search(r"REG:(\d+)", res).group(1)), int(re.search(r"STACK:(\d+)", res).group(1))
        # 40 stack bytes are libdevice's (the huge-argument reduction behind lg_sin_huge, division slow paths).  An
        # ALTERNATIVE launch shape (groups of 8 on a patch with four oscillators) may spill a few registers at the
        # 128-register cap: it then loses the measurement that picks the shape, it is not an error
        assert regs <= 128 and stack <= 160, (f, regs, stack)
        if os.path.basename(f).split(".")[0] in first:
            assert stack <= 48, (f, regs, stack)
    # a second request finds everything cached
    assert _patch(srk, _builders(srk)["cfg2"], 4096).precompile(4096) == 0


def test_second_nvrtc_version_doubles_the_fused_candidates(srk, tmp_path):
    """Neither NVRTC 12.8 nor 12.9 schedules every kernel better (profiles/r06h_tune_all*.txt), so each fused candidate
    is built by both when both are at hand and the measurement picks; SRK_NVRTC_ALT=0 leaves the toolkit's alone.
    The version is part of the kernel id."""
    code = ("import srack_b200 as s; p = s.Patch(); s.patches.cfg2(p, 65536); p.plan(); "
            "print('N', p.precompile(65536), p.kernel_id(65536))")
    out = {}
    for alt in ("0", None):
        env = dict(os.environ, SRK_KERNEL_CACHE=str(tmp_path / f"alt{alt}"))
        env.pop("SRK_NVRTC_ALT", None)
        if alt is not None:
            env["SRK_NVRTC_ALT"] = alt
        os.makedirs(env["SRK_KERNEL_CACHE"], exist_ok=True)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
        assert r.returncode == 0, r.stderr
        line = [l for l in r.stdout.splitlines() if l.startswith("N ")][0].split()
        out[alt] = (int(line[1]), line[2], set(os.path.basename(f) for f in glob.glob(os.path.join(env["SRK_KERNEL_CACHE"], "*.cubin"))))
    n0, id0, files0 = out["0"]
    n1, id1, files1 = out[None]
    if n0 == 0:
        pytest.skip("no NVRTC on this machine")
    assert id0 == id1  # the cost model's first choice is the toolkit compiler's kernel either way
    if n1 == n0:
        pytest.skip("no second NVRTC version in this environment")
    assert n1 == 2 * n0 and files0 < files1 and len(files1) == 2 * len(files0)


def test_forced_knobs_reach_the_generator(srk, monkeypatch):
    p = _patch(srk, srk.patches.cfg2, 4096)
    default = p.fused_source(4096)
    seen = {default}
    for env in ({"SRK_FUSED_GROUP": "8"}, {"SRK_FUSED_STAGES": "3"}, {"SRK_FUSED_SPLIT_MOOG": "0"}, {"SRK_FUSED_PREFETCH": "1"},
                {"SRK_FUSED_TILE_ROWS": "16"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        src = _patch(srk, srk.patches.cfg2, 4096).fused_source(4096)
        assert src and src not in seen, env
        seen.add(src)
        for k in env:
            monkeypatch.delenv(k)
    monkeypatch.setenv("SRK_FUSED", "0")
    assert _patch(srk, srk.patches.cfg2, 4096).fused_source(4096) == ""


@pytest.mark.gpu
def test_constant_fed_detectors(srk, orc, cuda_device):
    """A Math module with no inputs is how a patch spells a constant; fed into a transition detector it made nvcc 12.9
    fold `select(x > 0, select(x <= 0, a, b), c)` into `select(x > 0, a, 0)` at -O2 (the op templates route such outputs
    through fz::opaque).  Constant-high, constant-low and constant-NaN gates into ADSRs, a constant sync into an oscillator."""
    P = srk.PARAM

    def build(b, n_voices, seed=0):
        hi, lo, nan = (b.module_create("ADD") for _ in range(3))
        b.set_param(hi, P["MATH_CONSTANT"], 0.75)
        b.set_param(lo, P["MATH_CONSTANT"], -0.5)
        b.set_param(nan, P["MATH_CONSTANT"], float("nan"))
        adsrs = [b.module_create("ADSR") for _ in range(3)]
        for a, src in zip(adsrs, (hi, lo, nan)):
            for pid, v in zip(("ADSR_A_SEC", "ADSR_D_SEC", "ADSR_S_VAL", "ADSR_R_SEC"), (0.002, 0.004, 0.6, 0.003)):
                b.set_param(a, P[pid], v)
            b.connect(a, 0, src, 0)
        osc = b.module_create("OSCILLATOR")
        b.set_param_per_voice(osc, P["OSC_VAL"], srk.patches._u(3, 1, n_voices, -2.0, 1.0))
        b.connect(osc, 1, hi, 0)  # a constant-high sync: `last` starts true, so it never resets
        mix = b.module_create("MONO_MIXER")
        for i, a in enumerate(adsrs):
            b.connect(mix, i, a, 0)
        b.connect(mix, 3, osc, 2)
        out = b.module_create("OUTPUT")
        b.connect(out, 0, mix, 0)
        b.connect(out, 1, adsrs[0], 0)
        return {}

    gp, op, _, _ = build_both(srk, orc, build, 70)
    gp.plan()
    g, _ = gp.render(70, 3000, stems=True)
    o, _ = op.render(70, 3000)
    # the detector starts with last = true: a gate that is high from the first sample is no transition, but mode None
    # starts the attack on `gate > 0` (adsr.rs:145-150) -- the envelope runs once and sustains
    assert float(o[1].max()) > 0.9 and abs(float(o[1, -1, 0]) - 0.6) < 1e-6
    assert_parity(g, o, exact=True, what="constant-fed detectors")

import os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import srack_b200 as srk
from test_gpu_parity import _fuzz_patch
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rng = random.Random(1000 + seed)
kinds, wires = _fuzz_patch(rng, rng.randrange(3, 12), with_sample=seed >= 24)
B = rng.choice([1, 5, 64, 1024]); V = rng.choice([1, 31, 33, 64])
N = 64 * max(1, 1024 // 64) if B > 64 else 640 // B * B
n_par = dict(OSCILLATOR=1, ADSR=4, MOOG_FILTER=3, MONO_MIXER=4, ADD=1, SUBTRACT=1, MULTIPLY=1, NON_LINEAR=1)
ranges = dict(OSCILLATOR=(-3, 3), ADSR=(0.0, 0.01), MOOG_FILTER=(0.05, 0.9), MONO_MIXER=(0, 1), ADD=(-1, 1),
              SUBTRACT=(-1, 1), MULTIPLY=(-1, 1), NON_LINEAR=(0.5, 2.0))
pvals = {(m, pid): (rng.random() < 0.5, srk.patches._u(seed, 17 * m + pid, V, *ranges[k]))
         for m, k in enumerate(kinds) if k in n_par for pid in range(n_par[k])}
tables = {}
for m, k in enumerate(kinds):
    steps = rng.choice([1, 3, 16, 64])
    if k == "GRID_SEQUENCER":
        tables[m] = np.array([-1 if rng.random() < 0.3 else srk.grid_cell(rng.randrange(24), rng.random() < 0.5)
                              for _ in range(steps)], dtype=np.int32)
    elif k == "PATTERN_SEQUENCER":
        tables[m] = np.array([[rng.choice([-1, 0, 1]) for _ in range(steps)] for _ in range(8)], dtype=np.int32)
waves = {m: (np.random.default_rng(seed * 100 + m).uniform(-1, 1, rng.choice([0, 1, 97, 4000])).astype(np.float32),
             rng.choice([8000.0, 44100.0, 192000.0])) for m, k in enumerate(kinds) if k == "SAMPLE"}
print("B", B, "V", V, "N", N, {k: (v[0], v[1][:2]) for k, v in pvals.items()}, {m: (len(w), r) for m, (w, r) in waves.items()})
res = {}
for fused in ("0", "1"):
    os.environ["SRK_FUSED"] = fused
    b = srk.Patch(srk.AudioConfig(48000, B, 2))
    mods = [b.module_create(k) for k in kinds]
    for sink, i, src, port in wires:
        b.connect(mods[sink], i, mods[src], port)
    for m, cells in tables.items():
        b.set_sequence(mods[m], cells)
    for m, (wave, rate) in waves.items():
        b.set_sample(mods[m], wave, rate)
    for (m, pid), (per_voice, vals) in pvals.items():
        if per_voice:
            b.set_param_per_voice(mods[m], pid, vals)
        else:
            b.set_param(mods[m], pid, float(vals[0]))
    b.plan()
    st, mx = b.render(V, N, stems=True, mix=True)
    res[fused] = st
    print("fused", fused, b.program_info(V))
    print(st[0, :24, 0]); print(st[1, :24, 0])
    if fused == "1":
        print(b.fused_source(V)[:1800])
print("equal", np.array_equal(res["0"], res["1"]), np.argwhere(res["0"] != res["1"])[:10])

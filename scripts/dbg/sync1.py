import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import srack_b200 as srk
P = srk.PARAM
def build(per_voice, const, B, V):
    b = srk.Patch(srk.AudioConfig(48000, B, 2))
    osc = b.module_create("OSCILLATOR"); add = b.module_create("ADD"); out = b.module_create("OUTPUT")
    b.connect(osc, 1, add, 0); b.connect(out, 0, osc, 1); b.connect(out, 1, add, 0)
    b.set_param(add, P["MATH_CONSTANT"], const)
    if per_voice: b.set_param_per_voice(osc, P["OSC_VAL"], np.full(V, -2.73, np.float32))
    else: b.set_param(osc, P["OSC_VAL"], -2.73)
    b.plan()
    return b
for per_voice in (False, True):
  for const in (0.0, 0.86):
    for V in (1, 32):
      r = {}
      for fused in ("0", "1"):
        os.environ["SRK_FUSED"] = fused
        b = build(per_voice, const, 1024, V)
        st, _ = b.render(V, 64, stems=True)
        r[fused] = st
      print("per_voice", per_voice, "const", const, "V", V, "equal", np.array_equal(r["0"], r["1"]))
      if not np.array_equal(r["0"], r["1"]):
        print(" interp", r["0"][0, :12, 0], r["0"][1, :6, 0]); print(" fused ", r["1"][0, :12, 0], r["1"][1, :6, 0])

"""Numeric ids of the C ABI (include/srack_b200.h) and the small helpers that go with them.  Pure Python: importing
this module loads nothing -- the patch descriptions (patches.py) depend on it alone, so a backend other than the product
(the test oracle, bench.py's CPU reference arm) can use them without loading libsrack_b200.so."""

OPS = ["END", "RING_LOAD", "RING_STORE", "OSC", "NOISE", "MOOG", "ADSR", "VCA", "MIXER", "MATH", "OUTPUT", "MIX",
       "MOOG_COEF", "GRIDSEQ", "PATSEQ", "OSC_DELTA", "SAMPLE", "OSC_PHASE", "OSC_SHAPE"]
SEQ_NONE = -1


def grid_cell(val, hold=True):
    """SRK_GRID_CELL: Some((val, hold)) of the grid sequencer's table."""
    return (int(val) & 0xFFFF) | (0x10000 if hold else 0)

# status codes / kinds / params / flags: keep in sync with include/srack_b200.h
# (tests/test_abi.py parses the header and compares)
STATUS = dict(OK=0, ERR_ARG=1, ERR_PORT=2, ERR_KIND=3, ERR_UNSUPPORTED=4, ERR_PARAM=5, ERR_SELF_LOOP=6,
              ERR_NO_OUTPUT=7, ERR_NOT_PLANNED=8, ERR_SIZE=9, ERR_NO_DEVICE=10, ERR_CUDA=11, ERR_LIMIT=12)
KIND = dict(OUTPUT=0, OSCILLATOR=1, NOISE=2, ADSR=3, VCA=4, MOOG_FILTER=5, MONO_MIXER=6, ADD=7, SUBTRACT=8,
            MULTIPLY=9, NON_LINEAR=10, GRID_SEQUENCER=11, PATTERN_SEQUENCER=12, SAMPLE=13)
PARAM = dict(OSC_VAL=0, OSC_ANTIALIASING=1, ADSR_A_SEC=0, ADSR_D_SEC=1, ADSR_S_VAL=2, ADSR_R_SEC=3, VCA_NEGATIVE=0,
             MOOG_FREQ=0, MOOG_RES=1, MOOG_EXP_AMT=2, MIXER_GAIN0=0, MIXER_GAIN1=1, MIXER_GAIN2=2, MIXER_GAIN3=3,
             GRIDSEQ_STEPS_PER_OCTAVE=0,             MATH_CONSTANT=0)
RENDER_DEVICE_OUT = 1
RENDER_ASYNC = 2

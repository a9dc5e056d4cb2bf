"""Generates tests/golden/subtractive_b16.srk: a small .srk patch file written by the schema-driven
restatement of the reference's FileFormat (oracle/srk_file.py; rmp-serde 1.3.0 layout), with
mid-performance DSP state and non-zero port buffers, 16-sample buffers.  The reference ships no .srk file
and cannot be run here, so this fixture pins the two codecs (Python restatement, C++ product) against
regressions, not against the reference.

    python tests/golden/make_srk_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import srk_file as sf  # noqa: E402
from test_srk_file import subtractive_file  # noqa: E402

data = sf.dumps(subtractive_file(B=16))
open(os.path.join(HERE, "subtractive_b16.srk"), "wb").write(data)
print(len(data), "bytes")

"""Writes a raw BGR8 synthetic clip for tests/cpp/api_driver (tools/make_raw_clip.py kind w h n out.bin)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
kind, w, h, n, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
open(out, "wb").write(b"".join(f.tobytes() for f in synth.clip(kind, w, h, n)))

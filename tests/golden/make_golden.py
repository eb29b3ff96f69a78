"""Generates tests/golden/golden_small.npz from the oracle (numpy restatement, cross-checked against
cv2 in tests/test_oracle_restate.py).  The reference has no golden vectors of its own
and cannot be compiled here, so these are minted from the oracle:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_chain as RC, restate as R, synth  # noqa: E402

out = {}
img = synth.frame("scene", 48, 32, 5)
out["scene_48x32_t5"] = img
out["bicubic_x2"] = R.resize_cubic(img, 96, 64)
out["bicubic_x4"] = R.resize_cubic(img, 192, 128)
out["bilateral_d5"] = R.bilateral_u8c3(img, 5, 24.0, 24.0)
scfg = RC.SharpenConfig(adaptive_sigma=False)
out["edge_mask"] = RC.edge_mask(RC.NumpyOps(), img, scfg)
out["sharpen"] = RC.sharpen(RC.NumpyOps(), img, scfg)
cfg = RC.ChainConfig(dst_w=96, dst_h=64, temporal=RC.TEMPORAL_BLEND_FRAMES)
ch = RC.Chain(RC.NumpyOps(), cfg)
for t in range(3):
    out[f"chain_t{t}"] = ch.process(synth.frame("scene", 48, 32, t))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
print({k: v.shape for k, v in out.items()})

"""Generates tests/golden/golden_widened.npz for the rows widened per SURVEY 8(f): every vector comes out of the REAL
OpenCV kernels through cv2 (oracle/ref_chain.py with the Cv2Ops backend; the repaired glue around them is the oracle's),
so the numpy restatement and the CUDA kernels are held against frozen library output even where cv2 is not installed:
    python tests/golden/make_golden_widened.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_chain as RC, restate as R, synth  # noqa: E402

ops = RC.Cv2Ops()
out = {"cv2_version": np.array(cv2.__version__)}
img = synth.frame("scene", 96, 64, 5)
nxt = synth.frame("scene", 96, 64, 6)
out["pyr_down"] = cv2.pyrDown(img)
out["pyr_up"] = cv2.pyrUp(out["pyr_down"], dstsize=(96, 64))
bcfg = RC.BilateralConfig()
out["detail_mask"] = RC.create_detail_mask(ops, img, bcfg)
out["selective"] = RC.selective_bilateral_process(ops, img, RC.BilateralConfig(selective=True, use_multiscale=False))
out["multiscale"] = RC.selective_bilateral_process(ops, img, RC.BilateralConfig(use_multiscale=True, num_scales=3))
scfg = RC.SharpenConfig(adaptive_sigma=True)
out["sigma_map"] = RC.adaptive_sigma_map(ops, RC.texture_map(ops, img))
out["variable_sigma_sharpen"] = RC.sharpen_variable_sigma(ops, img, scfg)
tcfg = RC.TemporalConfig(use_flow=True, scene_detect=True)
flow = ops.farneback(ops.bgr2gray(img), ops.bgr2gray(nxt), tcfg)
out["farneback_flow"] = flow
out["warp"] = ops.remap_flow(img, flow)
big = (flow * np.float32(12.0)).astype(np.float32)          # magnitudes on both sides of motion_threshold
out["reliability_mask"] = RC.flow_reliability_mask(ops, big, tcfg.motion_threshold)
tc = RC.TemporalConsistencyFull(tcfg, ops)
for t in range(5):
    out[f"tc_full_t{t}"] = tc.process(synth.frame("scene", 96, 64, t))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_widened.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})

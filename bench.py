#!/usr/bin/env python
"""bench.py - 1080p->4K enhance-chain frames/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W                     # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               # the reference's OpenCV CPU path

Workload (config.workload): file/--fast mode of the chain
    SelectiveBilateral(PRE) -> cv::resize(INTER_CUBIC) x2 -> AdaptiveSharpening -> SelectiveBilateral(POST)
    -> temporal blend,   1920x1080 -> 3840x2160, 8-bit BGR, synthetic `scene` clip.
A "step" is one contiguous segment of --frames frames per GPU (weak scaling: N GPUs process N segments
of one clip; rank k>0 first runs the temporal warm-up frames of its segment, inside the timed region).
`value` = frames of all ranks / max-over-ranks device time (CUDA events), inputs resident in HBM.
`e2e`   = the same through vp_submit_host/vp_wait with pinned HOST frames (H2D + D2H inside the timed
region, H2D / compute-lane / D2H streams).  The roofline object describes the kernel that takes the most device time.
torch is plumbing only (device buffers, stream, events, process group); every kernel is libvpb200's.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (src_w, src_h, dst_w, dst_h, use_pre, use_sharpen, use_post, temporal)
    "c4": (1920, 1080, 3840, 2160, 1, 1, 1, "frames"),   # C2 chain + temporal (C4 at n_gpus = N)
    "c2": (1920, 1080, 3840, 2160, 1, 1, 1, "none"),
    "c3": (3840, 2160, 7680, 4320, 1, 1, 1, "frames"),
    "c1": (1280, 720, 2560, 1440, 0, 0, 0, "none"),      # bare cv::resize(INTER_CUBIC) x2
    "c5": (960, 540, 3840, 2160, 0, 0, 0, "none"),       # bare cv::resize(INTER_CUBIC) x4
    # C1 as `video_processor <file> --fast` literally runs it: Upscaler(BICUBIC) = CPUImpl's wrapper
    # (pre-blur, bicubic, enhanceBicubicResult) + main.cpp's 0.6/0.4 smoothing
    "c1w": (1280, 720, 2560, 1440, 0, 0, 0, "main"),
}
ENHANCE = {"c1w"}
SEED = 20260


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
# synthetic clip on the device (same integer hash as oracle/synth.py; tests/test_bench_synth.py pins it)
# ---------------------------------------------------------------------------------------------------
def _mix32(h):
    M = 0xFFFFFFFF
    h = h & M
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & M
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & M
    h = h ^ (h >> 15)
    return h


def torch_frame(kind, w, h, t, seed, device):
    import torch
    x = torch.arange(w, dtype=torch.int64, device=device)[None, :, None]
    y = torch.arange(h, dtype=torch.int64, device=device)[:, None, None]
    c = torch.arange(3, dtype=torch.int64, device=device)[None, None, :]
    M = 0xFFFFFFFF
    k = ((x * 0x9E3779B1) & M) ^ ((y * 0x85EBCA77) & M) ^ ((c * 0xC2B2AE3D) & M) ^ (((t + seed) * 0x27D4EB2F) & M)
    n = _mix32(k) & 0xFF
    if kind == "noise":
        return n.to(torch.uint8).contiguous()
    if kind == "flat":
        return (128 + (n & 3)).to(torch.uint8).contiguous()

    def tri(v, p):
        return torch.abs((v % (2 * p)) - p) * 255 // p
    v = tri(x + 2 * t + 13 * c, 96) * 3 // 8 + tri(y + t, 64) * 3 // 8 + 64 * (((x >> 5) ^ (y >> 5) ^ (t >> 4)) & 1) + (n & 15)
    return v.clamp(0, 255).to(torch.uint8).contiguous()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_for(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest ncu --set full
    capture summarised in profiles/traffic.json (tools/ncu_summary.py), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        entries = t[sorted(t)[-1]]
        # PRE is the first k_bilateral launch of a frame, POST the last ("#n" = order in the capture; older captures
        # only carry the grid size, where PRE had the smaller grid)
        bil = sorted((k for k in entries if "k_bilateral" in k),
                     key=lambda k: (int(k.split("#")[1]) if "#" in k else 0, int(k.split("@grid")[1].split("#")[0])))
        pick = {"stats": [k for k in entries if "k_stats" in k],
                "upsharp": [k for k in entries if "k_upsharp" in k or "k_bicubic" in k],
                "bilateral_pre": bil[:1], "bilateral_post": bil[-1:]}.get(kernel, [])
        return entries[pick[0]] if pick else None
    except Exception:
        return None


def warp_inst_per_frame():
    """smsp__inst_executed.sum per frame of the C4 chain from the newest committed ncu --set full capture
    (profiles/kernel_counters.json, written by tools/ncu_summary.py): {kernel: warp instructions}, tag."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as f:
            t = json.load(f)
        tag = sorted(t)[-1]
        return {k: v["warp_inst"] for k, v in t[tag].items()}, tag
    except Exception:
        return None, None


def copy_ceiling(dev, barrier, world, dist, inb, outb, depth, seconds=0.6):
    """tools/copy_peak.py's `both` leg (plain pinned cudaMemcpyAsync loops of this workload's per-frame transfers) on
    every rank at once: the box's host<->device ceiling at this N, measured in the same process group as `e2e`."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("copy_peak", os.path.join(ROOT, "tools", "copy_peak.py"))
    cp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cp)
    d, u, sec = cp.measure(dev, "both", seconds, barrier, d2h_bytes=outb, h2d_bytes=inb, depth=depth)
    tb = torch.tensor([d, u], dtype=torch.float64, device=dev)
    tt = torch.tensor([sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    d, u, sec = float(tb[0].item()), float(tb[1].item()), float(tt.item())
    return {"d2h_gbs": d / sec / 1e9, "h2d_gbs": u / sec / 1e9, "total_gbs": (d + u) / sec / 1e9, "frames_per_s": d / outb / sec}


def single_frame_latency(capi, torch, h, process, frames, outs, sw, dw, dh, stream, nlat):
    """BASELINE config C5's figure: one frame in flight, synchronise after each (device time from CUDA events, wall time
    launch-to-done, and the synchronous pinned-host call)."""
    def pct(v, q):
        v = sorted(v)
        return v[min(len(v) - 1, int(q * len(v)))]
    capi.lib.vp_reset_temporal(h)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nlat)]
    wall = []
    for i in range(nlat):
        t0 = time.perf_counter()
        evs[i][0].record()
        capi.check(process(h, frames[i % len(frames)].data_ptr(), sw * 3, outs[i % len(outs)].data_ptr(), dw * 3, stream), h)
        evs[i][1].record()
        evs[i][1].synchronize()
        wall.append((time.perf_counter() - t0) * 1e6)
    dev_us = [a.elapsed_time(b) * 1e3 for a, b in evs]
    hin1 = frames[0].cpu().pin_memory()
    hout1 = torch.empty((dh, dw, 3), dtype=torch.uint8).pin_memory()
    host = []
    for i in range(min(nlat, 300)):
        t0 = time.perf_counter()
        capi.check(capi.lib.vp_process_host(h, hin1.data_ptr(), sw * 3, hout1.data_ptr(), dw * 3), h)
        host.append((time.perf_counter() - t0) * 1e6)
    return {"frames": nlat, "unit": "us",
            "device": {"p50": round(pct(dev_us, .5), 2), "p99": round(pct(dev_us, .99), 2)},
            "launch_to_done_wall": {"p50": round(pct(wall, .5), 2), "p99": round(pct(wall, .99), 2)},
            "host_frame_sync_api": {"p50": round(pct(host, .5), 2), "p99": round(pct(host, .99), 2),
                                    "note": "vp_process_host: pinned H2D + kernels + pinned D2H, one frame in flight"}}


USE_FLOW = False             # --flow
ADAPTIVE_SIGMA = False       # --adaptive-sigma: AdaptiveSharpening's variable-sigma branch (repaired semantics)
BILATERAL_MODE = "plain"     # --bilateral-mode: plain (north-star chain) | selective | multiscale (repaired semantics)


def apply_bilateral_mode(pre, post):
    """Upscaler::initializeEnhancements' own flags (src/upscaler.cpp:668-725) on a pair of bilateral configs
    (ctypes structs or the oracle's dataclasses: same field names)."""
    if BILATERAL_MODE == "plain":
        return
    for b, ep, ns in ((pre, 2.0, 3), (post, 1.5, 2)):
        b.selective = 1
        b.use_multiscale = int(BILATERAL_MODE == "multiscale")
        b.num_scales = ns
        b.detail_threshold = 30.0
        b.edge_preserve = ep


def build_cfg(capi, wl, temporal):
    sw, sh, dw, dh, pre, sharpen, post, tdef = WORKLOADS[wl]
    cfg = capi.default_chain_config(sw, sh, dw, dh)
    cfg.use_pre, cfg.use_sharpen, cfg.use_post = pre, sharpen, post
    cfg.bicubic_enhance = 1 if wl in ENHANCE else 0
    t = temporal or tdef
    cfg.temporal.mode = {"none": capi.VP_TEMPORAL_NONE, "main": capi.VP_TEMPORAL_MAIN_BLEND, "frames": capi.VP_TEMPORAL_BLEND_FRAMES}[t]
    apply_bilateral_mode(cfg.pre, cfg.post)
    cfg.sharpen.adaptive_sigma = int(ADAPTIVE_SIGMA)
    if USE_FLOW:
        cfg.temporal.use_flow = 1
        cfg.temporal.scene_detect = 1
    return cfg, t


def algorithmic_bytes(wl, temporal, cfg):
    """in + out (+ state reads + state write when the temporal stage is on), per frame."""
    sw, sh, dw, dh = WORKLOADS[wl][:4]
    inb, outb = sw * sh * 3, dw * dh * 3
    nprev = {"none": 0, "main": 1, "frames": max(cfg.temporal.buffer_size, 0)}[temporal]
    chain = inb + outb + (nprev * outb + outb if nprev else 0)
    if wl in ENHANCE:   # profile kinds are reused: stats = gauss3, upsharp = bicubic, blend = canny (nms + hysteresis)
        return chain, {"stats": 2 * inb, "upsharp": inb + outb, "blend": outb + 2 * (outb // 3),
                       "bilateral_post": 2 * outb + outb // 3 + (nprev * outb + outb if nprev else 0)}, inb, outb
    per_kernel = {
        "stats": inb,
        "bilateral_pre": 2 * inb,
        "upsharp": inb + outb,
        "bilateral_post": 2 * outb + (nprev * outb + outb if nprev else 0),
        "blend": outb + outb + (nprev * outb + outb if nprev else 0),
    }
    return chain, per_kernel, inb, outb


# ---------------------------------------------------------------------------------------------------
def cpu_chain_fps(wl, temporal, nframes, kind, threads=None):
    """The reference's OpenCV CPU path: oracle/ref_chain.py over the real cv2 kernels (the one place
    bench.py executes oracle/)."""
    import cv2
    from oracle import ref_chain as RC, synth
    sw, sh, dw, dh, pre, sharpen, post, tdef = WORKLOADS[wl]
    t = temporal or tdef
    ocfg = RC.ChainConfig(dst_w=dw, dst_h=dh, use_pre=bool(pre), use_sharpen=bool(sharpen), use_post=bool(post),
                          temporal={"none": RC.TEMPORAL_NONE, "main": RC.TEMPORAL_MAIN_BLEND, "frames": RC.TEMPORAL_BLEND_FRAMES}[t],
                          bicubic_enhance=wl in ENHANCE)
    apply_bilateral_mode(ocfg.pre, ocfg.post)
    ocfg.sharpen.adaptive_sigma = bool(ADAPTIVE_SIGMA)
    if USE_FLOW:
        ocfg.temporal_cfg.use_flow = True
        ocfg.temporal_cfg.scene_detect = True
    ops = RC.Cv2Ops(threads=threads or os.cpu_count())
    chain = RC.Chain(ops, ocfg)
    frames = synth.clip(kind, sw, sh, nframes, SEED)
    chain.process(frames[0])   # warm the temporal state and cv2's thread pool (untimed)
    t0 = time.perf_counter()
    for f in frames:
        chain.process(f)
    dt = time.perf_counter() - t0
    return nframes / dt, cv2.getNumThreads(), dt


def run_reference(args, rank):
    if rank != 0:
        return
    import cv2
    wl = args.workload
    per_step = args.ref_frames
    sw, sh, dw, dh = WORKLOADS[wl][:4]
    from oracle import ref_chain as RC, synth
    pre, sharpen, post, tdef = WORKLOADS[wl][4:]
    t = args.temporal or tdef
    ocfg = RC.ChainConfig(dst_w=dw, dst_h=dh, use_pre=bool(pre), use_sharpen=bool(sharpen), use_post=bool(post),
                          temporal={"none": RC.TEMPORAL_NONE, "main": RC.TEMPORAL_MAIN_BLEND, "frames": RC.TEMPORAL_BLEND_FRAMES}[t],
                          bicubic_enhance=wl in ENHANCE)
    apply_bilateral_mode(ocfg.pre, ocfg.post)
    ocfg.sharpen.adaptive_sigma = bool(ADAPTIVE_SIGMA)
    if USE_FLOW:
        ocfg.temporal_cfg.use_flow = True
        ocfg.temporal_cfg.scene_detect = True
    chain = RC.Chain(RC.Cv2Ops(threads=os.cpu_count()), ocfg)
    frames = synth.clip(args.kind, sw, sh, per_step, SEED)
    for _ in range(args.warmup):
        for f in frames:
            chain.process(f)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for f in frames:
            chain.process(f)
    dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "enhance_chain_frames_per_s", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(wl, t), "clip": args.kind, "frames_per_step": per_step,
                   "note": "reference CPU path = oracle/ref_chain.py over the real OpenCV kernels (cv2 %s); the reference's "
                           "C++ cannot be compiled here (no OpenCV C++), see DESIGN.md" % cv2.__version__},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cv2.getNumThreads(), "kind": "port",
                         "sample": f"{per_step} frames x {args.steps} steps of the same workload"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(wl, temporal):
    sw, sh, dw, dh, pre, sharpen, post, _ = WORKLOADS[wl]
    if wl in ENHANCE:
        stages = ["gauss3x3", "bicubic", "enhanceBicubicResult(bilateral5+canny+sharpen)"] + ([f"temporal_{temporal}"] if temporal != "none" else [])
        return f"{wl.upper()}: {sw}x{sh}->{dw}x{dh} BGR8 Upscaler(BICUBIC) " + "+".join(stages)
    stages = (["bilateral_pre"] if pre else []) + ["bicubic"] + (["adaptive_sharpen"] if sharpen else []) + \
             (["bilateral_post"] if post else []) + ([f"temporal_{temporal}"] if temporal != "none" else [])
    mode = "" if BILATERAL_MODE == "plain" else f" [SelectiveBilateral {BILATERAL_MODE} branch, repaired semantics]"
    if ADAPTIVE_SIGMA:
        mode += " [AdaptiveSharpening adaptive_sigma branch, repaired semantics]"
    if USE_FLOW:
        mode += " [TemporalConsistency with Farneback flow, warp and reliability masks]"
    return f"{wl.upper()}: {sw}x{sh}->{dw}x{dh} BGR8 " + "+".join(stages) + mode


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--temporal", default=None, choices=["none", "main", "frames"])
    ap.add_argument("--kind", default="scene", choices=["scene", "noise", "flat"])
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step (one segment)")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames per step of the reference arm")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c5-latency", action="store_true", help="skip the C5 single-frame latency sample of the default line")
    ap.add_argument("--zero-copy-out", action="store_true",
                    help="experiment (DESIGN.md 5): the last kernel writes its frames straight into mapped pinned host memory "
                         "instead of device memory + a D2H copy; reported as `zero_copy_out`, never as value / e2e")
    ap.add_argument("--scene-detect", action="store_true", help="enable TemporalConsistency::detectSceneChange (blend becomes its own kernel)")
    ap.add_argument("--lanes", type=int, default=0, help="stream lanes of the batch API (0 = library default)")
    ap.add_argument("--per-frame", action="store_true", help="one vp_process_device call per frame on one stream (no batch API)")
    ap.add_argument("--latency-frames", type=int, default=None,
                    help="single-frame latency sample (p50/p99); default 1000 for workload c5, else 0")
    ap.add_argument("--bilateral-mode", default="plain", choices=["plain", "selective", "multiscale"],
                    help="SelectiveBilateral branch of both bilateral stages (default: the north-star plain branch)")
    ap.add_argument("--flow", action="store_true", help="TemporalConsistency's Farneback / warp / reliability-mask front end (SURVEY 8(f)3)")
    ap.add_argument("--adaptive-sigma", action="store_true", help="AdaptiveSharpening's adaptive_sigma branch (default: the north-star unsharp mask)")
    args = ap.parse_args()
    global BILATERAL_MODE, ADAPTIVE_SIGMA, USE_FLOW
    BILATERAL_MODE = args.bilateral_mode
    ADAPTIVE_SIGMA = args.adaptive_sigma
    USE_FLOW = args.flow

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    if args.warmup < 3:
        print("warning: --warmup < 3 breaks the timing rules", file=sys.stderr)

    import torch
    import torch.distributed as dist
    from video_processor_b200 import build as vb
    if rank == 0:
        vb.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on file descriptor 1 when the first communicator is built: keep stdout to the
        # one JSON line by pointing fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from video_processor_b200 import capi

    wl = args.workload
    cfg, temporal = build_cfg(capi, wl, args.temporal)
    if args.scene_detect:
        cfg.temporal.scene_detect = 1
    sw, sh, dw, dh = WORKLOADS[wl][:4]
    chain_bytes, kbytes, inb, outb = algorithmic_bytes(wl, temporal, cfg)
    F = args.frames
    from video_processor_b200 import sharding
    # weak scaling: the clip has world*F frames, rank k owns [k*F, (k+1)*F) and first feeds the warm-up frames
    first, warm, seg_start, seg_stop = sharding.plan(world * F, world, rank, temporal, cfg.temporal.buffer_size)
    assert (seg_start, seg_stop) == (rank * F, (rank + 1) * F)

    # device-resident clip segment of this rank: frames [rank*F - warm, (rank+1)*F)
    seg = [torch_frame(args.kind, sw, sh, rank * F - warm + i, SEED, dev) for i in range(F + warm)]
    NOUT = max(2, min(8, (512 << 20) // outb))     # output ring larger than L2 together with the inputs
    outs = [torch.empty((dh, dw, 3), dtype=torch.uint8, device=dev) for _ in range(NOUT)]
    ctx = capi.Context(cfg, device=local_rank)
    # a non-default torch stream: the chain's kernels are launched on it (vp_process_device's stream
    # argument) and torch.cuda.Event brackets them on the same stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    h = ctx.handle
    process = capi.lib.vp_process_device

    # file/--fast mode: the whole segment is handed over in one call (vp_process_batch_device), frames
    # alternate between two internal streams; the context's compute stream is made to follow torch's
    # stream so the CUDA events below bracket the work
    src_ptrs = (C.c_void_p * (warm + F))(*[t.data_ptr() for t in seg])
    dst_ptrs = (C.c_void_p * (warm + F))(*([None] * warm + [outs[i % NOUT].data_ptr() for i in range(F)]))
    batch = capi.lib.vp_process_batch_device
    if args.lanes:
        capi.check(capi.lib.vp_set_batch_lanes(h, args.lanes), h)

    def step_per_frame():
        capi.lib.vp_reset_temporal(h)
        for i in range(warm):
            capi.check(process(h, seg[i].data_ptr(), sw * 3, None, 0, stream), h)
        for i in range(F):
            capi.check(process(h, seg[warm + i].data_ptr(), sw * 3, outs[i % NOUT].data_ptr(), dw * 3, stream), h)

    def step():
        if args.per_frame:
            return step_per_frame()
        capi.lib.vp_reset_temporal(h)
        capi.check(batch(h, warm + F, src_ptrs, sw * 3, dst_ptrs, dw * 3, stream), h)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # both APIs take torch's stream (the batch call forks its second lane from it and joins it again),
    # so these CUDA events bracket all the device work of the timed steps
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    # parity_check (rank 0): the frames the timed batch call left in the output ring, to be compared below with what
    # one vp_process_device call per frame writes there
    timed_tail = [o.clone() for o in outs] if rank == 0 and not args.per_frame else None
    barrier()
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tl = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    ms_max = float(tms.item())
    fps = world * F * args.steps / (ms_max / 1e3)

    # ---- per-kernel device times (second pass, events around every launch) -> roofline ---------------
    roof = None
    kernels = {}
    parity = {}
    if rank == 0:
        capi.lib.vp_profile_enable(h, 1)
        step_per_frame()
        torch.cuda.synchronize()
        prof = ctx.profile_read()
        capi.lib.vp_profile_enable(h, 0)
        if timed_tail is not None:
            import hashlib
            same = all(torch.equal(a, b) for a, b in zip(timed_tail, outs))
            parity["batch_api_vs_per_frame"] = {"result": "ok" if same else "MISMATCH", "frames_compared": min(NOUT, F),
                                                "sha256_last_frame": hashlib.sha256(outs[(F - 1) % NOUT].cpu().numpy().tobytes()).hexdigest()[:16]}
            del timed_tail
        peak, peak_src = peaks()
        tot = sum(v[0] for v in prof.values()) or 1.0
        for k, (tms_k, n) in prof.items():
            if n:
                us = tms_k / n * 1e3
                kernels[k] = {"launches_per_step": n, "avg_us": round(us, 2), "share": round(tms_k / tot, 4),
                              "alg_bytes": kbytes.get(k), "achieved_gbs": round(kbytes[k] / (us * 1e-6) / 1e9, 1) if k in kbytes else None}
        dom = max((k for k in kernels if kernels[k]["achieved_gbs"] is not None), key=lambda k: kernels[k]["share"])
        traffic = traffic_for(dom)
        ach = kernels[dom]["achieved_gbs"]
        per_gpu_fps = fps / world
        chain = {"alg_bytes_per_frame": chain_bytes, "achieved": round(chain_bytes * per_gpu_fps / 1e9, 1),
                 "frac": round(chain_bytes * per_gpu_fps / 1e9 / peak, 4),
                 "counting": "in + out + state reads + state write as the code moves them (blendFrames reads buffer_size previous frames)"}
        if temporal == "frames":
            # SURVEY 8(d) counted one state read (its a7 row says "<= 2 previous"); the code blends buffer_size = 3
            # previous frames, hence two figures for the same run
            b8d = inb + outb + 2 * outb
            chain["survey_8d"] = {"alg_bytes_per_frame": b8d, "achieved": round(b8d * per_gpu_fps / 1e9, 1),
                                  "frac": round(b8d * per_gpu_fps / 1e9 / peak, 4)}
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "avg_us": kernels[dom]["avg_us"], "share_of_step": kernels[dom]["share"],
                "chain": chain, "kernels": kernels}
        # the bound that binds these stencils: warp-instruction issue (one per SM sub-partition and clock)
        winst, wtag = warp_inst_per_frame()
        if winst and wl in ("c4", "c2") and BILATERAL_MODE == "plain" and not ADAPTIVE_SIGMA and not USE_FLOW:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            per_frame = sum(winst.values())
            ipeak = sms * 4 * mhz * 1e6
            roof["secondary"] = {"bound": "issue", "unit": "warp-inst/s", "warp_inst_per_frame": per_frame, "per_kernel": winst,
                                 "source": f"smsp__inst_executed.sum, profiles/kernel_counters.json [{wtag}]",
                                 "achieved": per_frame * per_gpu_fps, "peak": ipeak, "peak_def": f"{sms} SMs x 4 schedulers x {mhz:.0f} MHz",
                                 "frac": round(per_frame * per_gpu_fps / ipeak, 4)}

    # ---- experiment: output frames written by the POST kernel itself into mapped pinned host memory ----
    zero_copy = None
    if args.zero_copy_out and not args.per_frame:
        hz = [torch.empty((dh, dw, 3), dtype=torch.uint8).pin_memory() for _ in range(NOUT)]     # UVA: device-visible at the same address
        zdst = (C.c_void_p * (warm + F))(*([None] * warm + [hz[i % NOUT].data_ptr() for i in range(F)]))

        def zstep():
            capi.lib.vp_reset_temporal(h)
            capi.check(batch(h, warm + F, src_ptrs, sw * 3, zdst, dw * 3, stream), h)

        for _ in range(2):
            zstep()
        barrier()
        z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        z0.record()
        for _ in range(max(1, min(args.steps, 3))):
            zstep()
        z1.record()
        torch.cuda.synchronize()
        zms = z0.elapsed_time(z1) / max(1, min(args.steps, 3))
        same = bool(torch.equal(hz[(F - 1) % NOUT], outs[(F - 1) % NOUT].cpu())) if rank == 0 else None
        zero_copy = {"frames_per_s_per_gpu": F / (zms / 1e3), "ms_per_step": zms, "equals_device_output": same,
                     "what": "device-resident inputs, output frames stored by the last kernel into mapped pinned host memory (no D2H copy)"}

    # ---- e2e: pinned host frames through vp_submit_host / vp_wait -----------------------------------
    e2e = None
    if not args.no_e2e:
        nin = min(F + warm, 16)
        hin = [seg[i].cpu().pin_memory() for i in range(nin)]
        depth = capi.lib.vp_pipeline_depth(h)
        hout = [torch.empty((dh, dw, 3), dtype=torch.uint8).pin_memory() for _ in range(depth + 1)]
        submit, wait = capi.lib.vp_submit_host, capi.lib.vp_wait

        def e2e_step():
            capi.lib.vp_reset_temporal(h)
            inflight = 0
            for i in range(warm + F):
                if inflight == depth:
                    capi.check(wait(h), h)
                    inflight -= 1
                dst = hout[i % (depth + 1)].data_ptr() if i >= warm else None
                capi.check(submit(h, hin[i % nin].data_ptr(), sw * 3, dst, dw * 3), h)
                inflight += 1
            while inflight:
                capi.check(wait(h), h)
                inflight -= 1

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        nsteps = max(1, min(args.steps, 3))
        for _ in range(nsteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        twarm = torch.tensor([warm], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(twarm, op=dist.ReduceOp.SUM)
        e2e_fps = world * F * nsteps / float(tdt.item())
        h2d_step, d2h_step = (world * F + int(twarm.item())) * inb, world * F * outb
        e2e = {"value": e2e_fps, "unit": "frames/s",
               "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
               "timing": "host wall clock around submit/wait of whole steps, synchronised both sides, max over ranks",
               "api": "vp_submit_host/vp_wait (pinned host frames, H2D + compute lanes + D2H streams, depth %d)" % depth}
        # the same transfers as plain pinned copy loops on every rank at once: what the box's PCIe / host fabric allows
        # at this N (tools/copy_peak.py); e2e is bound by it, not by a kernel
        try:
            ceil = copy_ceiling(dev, barrier, world, dist, inb, outb, depth)
            ach = (h2d_step + d2h_step) * nsteps / float(tdt.item()) / 1e9
            e2e["roofline"] = {"bound": "host<->device copies (PCIe / host fabric)", "unit": "GB/s",
                               "achieved_gbs": round(ach, 2), "peak_gbs": round(ceil["total_gbs"], 2),
                               "frac": round(ach / ceil["total_gbs"], 4),
                               "peak_d2h_gbs": round(ceil["d2h_gbs"], 2), "peak_h2d_gbs": round(ceil["h2d_gbs"], 2),
                               "frames_per_s_ceiling": round(ceil["frames_per_s"], 1),
                               "peak_source": f"measured in this run at N={world}: concurrent pinned cudaMemcpyAsync loops, "
                                              f"{outb} B d2h + {inb} B h2d per frame, depth {depth} (tools/copy_peak.py)"}
        except Exception as ex:   # the ceiling is a report, never a reason to lose the line
            e2e["roofline"] = {"error": repr(ex)}
        # parity_check of the host path (rank 0): the first frames of the segment through submit/wait against one
        # vp_process_device call per frame
        if rank == 0:
            npar = min(8, F)
            capi.lib.vp_reset_temporal(h)
            ref_out = [torch.empty((dh, dw, 3), dtype=torch.uint8, device=dev) for _ in range(npar)]
            for i in range(warm + npar):
                capi.check(process(h, seg[i].data_ptr(), sw * 3, ref_out[i - warm].data_ptr() if i >= warm else None, dw * 3, stream), h)
            torch.cuda.synchronize()
            capi.lib.vp_reset_temporal(h)
            pin_in = [seg[i].cpu().pin_memory() for i in range(warm + npar)]
            pin_out = [torch.zeros((dh, dw, 3), dtype=torch.uint8).pin_memory() for _ in range(npar)]
            inflight = 0
            for i in range(warm + npar):
                if inflight == depth:
                    capi.check(wait(h), h)
                    inflight -= 1
                capi.check(submit(h, pin_in[i].data_ptr(), sw * 3, pin_out[i - warm].data_ptr() if i >= warm else None, dw * 3), h)
                inflight += 1
            while inflight:
                capi.check(wait(h), h)
                inflight -= 1
            same = all(torch.equal(a.cpu(), b) for a, b in zip(ref_out, pin_out))
            parity["host_api_vs_per_frame"] = {"result": "ok" if same else "MISMATCH", "frames_compared": npar}
            del ref_out, pin_in, pin_out

    # ---- single-frame latency (BASELINE config C5): one frame in flight, synchronise after each ---------
    latency = None
    nlat = args.latency_frames if args.latency_frames is not None else (1000 if wl == "c5" else 0)
    if rank == 0 and nlat > 0:
        latency = single_frame_latency(capi, torch, h, process, seg, outs, sw, dw, dh, stream, nlat)
    # BASELINE config C5 (bare INTER_CUBIC x4 960x540 -> 3840x2160, single-frame latency) rides along with the default
    # line so the driver-run record carries it: its own small context, 1000 single frames
    latency_c5 = None
    if rank == 0 and wl == "c4" and not args.no_c5_latency and BILATERAL_MODE == "plain" and not ADAPTIVE_SIGMA and not USE_FLOW:
        cfg5, _ = build_cfg(capi, "c5", None)
        w5, h5, dw5, dh5 = WORKLOADS["c5"][:4]
        ctx5 = capi.Context(cfg5, device=local_rank)
        fr5 = [torch_frame(args.kind, w5, h5, i, SEED + 5, dev) for i in range(32)]
        out5 = outs if (dw5, dh5) == (dw, dh) else [torch.empty((dh5, dw5, 3), dtype=torch.uint8, device=dev) for _ in range(4)]
        latency_c5 = single_frame_latency(capi, torch, ctx5.handle, process, fr5, out5, w5, dw5, dh5, stream, 1000)
        latency_c5["workload"] = workload_name("c5", "none")
        ctx5.close()

    cpu = None
    if rank == 0 and world == 1 and args.cpu_frames > 0:
        v, cores, dt = cpu_chain_fps(wl, temporal, args.cpu_frames, args.kind)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_frames} frames of the same workload ({dt:.1f} s), oracle chain over cv2 kernels"}

    if rank == 0:
        line = {
            "metric": "enhance_chain_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(wl, temporal), "clip": args.kind, "frames_per_step_per_gpu": F,
                       "temporal_warmup_frames_rank_gt0": {"none": 0, "main": 1, "frames": max(cfg.temporal.buffer_size, 0)}[temporal],
                       "sharding": "contiguous frame segments, no collective",
                       "l2": f"inputs ({F}x{inb >> 20} MiB) and output ring ({NOUT}x{outb >> 20} MiB) exceed the 126 MB L2",
                       "roofline_pass": "per-kernel CUDA events in a separate pass of one step, one stream, one frame at a time",
                       "api": "vp_process_device per frame" if args.per_frame else "vp_process_batch_device (four stream lanes)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(tl.item()), "batch_graph_replays": int(capi.lib.vp_batch_graph_replays(h)), "zero_copy_out": zero_copy, "roofline": roof, "cpu_baseline": cpu, "latency": latency, "latency_c5": latency_c5,
            "parity_check": ("ok" if parity and all(v["result"] == "ok" for v in parity.values()) else ("FAILED" if parity else None)),
            "parity_detail": parity or None,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""copy_peak.py - the host<->device copy ceiling the `e2e` leg of bench.py runs against.

    python tools/copy_peak.py [--gpus 1,2,4,8] [--seconds 1.0] [--out profiles/r02_copy_peak.json]

For every N in --gpus: N processes (one per GPU, as bench.py runs), each with pinned host buffers and two
CUDA streams, loop plain copies of exactly the chain's per-frame transfers - one cudaMemcpyAsync per copy
(`tensor.copy_(non_blocking=True)` between a pinned host tensor and a device tensor of the same shape):

    d2h    24,883,200 B  (one 3840x2160 BGR8 output frame)  device -> pinned host
    h2d     6,220,800 B  (one 1920x1080 BGR8 input frame)   pinned host -> device
    both   the two loops at once on their own streams, one h2d per d2h (what vp_submit_host / vp_wait do per frame)

All ranks start behind a barrier; a rank's rate is bytes / its own CUDA-event time, the job's rate is total
bytes / max-over-ranks time.  `frames_per_s_ceiling` = the `both` leg's d2h frames per second over all ranks: the
number an e2e run of the C4 chain cannot exceed on this box.  No kernel of libvpb200 runs here.

bench.py imports `measure()` and runs the `both` leg inside its own process group, so every bench line carries the
ceiling of the box and the N it ran on (`e2e.roofline`).
"""
import argparse
import json
import os
import sys
import time

D2H_BYTES = 3840 * 2160 * 3
H2D_BYTES = 1920 * 1080 * 3


def measure(dev, mode, seconds, barrier, d2h_bytes=D2H_BYTES, h2d_bytes=H2D_BYTES, nbuf=4, depth=3):
    """One rank's copy loop.  Returns (d2h_bytes_moved, h2d_bytes_moved, elapsed_s) with the time taken from CUDA
    events on the copy streams.  `barrier()` is called right before the timed region starts."""
    import torch
    torch.cuda.set_device(dev)
    dd = [torch.empty(d2h_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    hd = [torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    du = [torch.empty(h2d_bytes, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    hu = [torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    for t in hd + hu:
        t.fill_(7)                                   # touch every page before timing
    s_d2h, s_h2d = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    do_d, do_u = mode in ("d2h", "both"), mode in ("h2d", "both")

    evs = [torch.cuda.Event() for _ in range(depth)]

    def burst(n):
        for i in range(n):
            if do_u:
                with torch.cuda.stream(s_h2d):
                    if do_d and i >= depth:
                        # frame i goes up only once frame i - depth has come down (the pipeline's depth), so the two
                        # directions stay interleaved for the whole burst instead of the small uploads finishing early
                        s_h2d.wait_event(evs[i % depth])
                    du[i % nbuf].copy_(hu[i % nbuf], non_blocking=True)
            if do_d:
                with torch.cuda.stream(s_d2h):
                    hd[i % nbuf].copy_(dd[i % nbuf], non_blocking=True)
                    if do_u:
                        evs[i % depth].record(s_d2h)

    burst(8)
    torch.cuda.synchronize()
    # calibrate the burst length for ~`seconds` of copying
    t0 = time.perf_counter()
    burst(16)
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / 16
    n = max(16, int(seconds / max(per, 1e-6)))
    barrier()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e0[0].record(s_d2h); e0[1].record(s_h2d)
    burst(n)
    e1[0].record(s_d2h); e1[1].record(s_h2d)
    torch.cuda.synchronize()
    ms = max(e0[0].elapsed_time(e1[0]) if do_d else 0.0, e0[1].elapsed_time(e1[1]) if do_u else 0.0)
    return (n * d2h_bytes if do_d else 0), (n * h2d_bytes if do_u else 0), ms / 1e3


def gpu_locality(index):
    """(numa_node, local_cpulist) of GPU `index` from sysfs, or (None, None)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        return int(open(path + "/numa_node").read()), open(path + "/local_cpulist").read().strip()
    except Exception:
        return None, None


def bind_local(index):
    """Pin this process to the CPUs next to GPU `index` BEFORE any pinned allocation (first touch then places the
    pages on that node).  Returns the cpu list used, or None when sysfs offers nothing narrower than all CPUs."""
    node, cpus = gpu_locality(index)
    if not cpus:
        return None
    ids = set()
    for part in cpus.split(","):
        a, _, b = part.partition("-")
        ids.update(range(int(a), int(b or a) + 1))
    ids &= os.sched_getaffinity(0)
    if not ids:
        return None
    os.sched_setaffinity(0, ids)
    return cpus


def _worker(rank, world, mode, seconds, bar, q, bind=False):
    import torch
    where = bind_local(rank) if bind else None
    d, u, s = measure(torch.device("cuda", rank), mode, seconds, bar.wait)
    q.put((rank, d, u, s, where))


def run(world, mode, seconds, bind=False):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    bar, q = ctx.Barrier(world), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, mode, seconds, bar, q, bind)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join()
    tmax = max(r[3] for r in res)
    d2h, h2d = sum(r[1] for r in res), sum(r[2] for r in res)
    out = {"n_gpus": world, "mode": mode, "seconds": round(tmax, 3),
           "d2h_gbs": round(d2h / tmax / 1e9, 2), "h2d_gbs": round(h2d / tmax / 1e9, 2),
           "total_gbs": round((d2h + h2d) / tmax / 1e9, 2),
           "per_rank_gbs": [round((r[1] + r[2]) / r[3] / 1e9, 2) for r in res]}
    if bind:
        out["bound_to_cpus"] = [r[4] for r in res]
    if mode == "both":
        out["frames_per_s_ceiling"] = round(d2h / D2H_BYTES / tmax, 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--modes", default="d2h,h2d,both")
    ap.add_argument("--bind", action="store_true", help="pin every rank to its GPU's local CPUs before allocating pinned memory")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    rows = []
    for n in [int(x) for x in a.gpus.split(",")]:
        if n > have:
            print(f"skip N={n}: only {have} device(s)", file=sys.stderr)
            continue
        for mode in a.modes.split(","):
            r = run(n, mode, a.seconds, a.bind)
            print(json.dumps(r), flush=True)
            rows.append(r)
    doc = {"what": "concurrent pinned-memory copy ceiling, one process per GPU, one cudaMemcpyAsync per copy "
                   "(24,883,200 B d2h / 6,220,800 B h2d per frame of the C4 chain)",
           "gpu": torch.cuda.get_device_name(0), "devices_on_box": have, "host_cpus": os.cpu_count(),
           "gpu_locality": [gpu_locality(i) for i in range(have)], "rows": rows}
    if a.out:
        with open(a.out, "w") as f:
            json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()

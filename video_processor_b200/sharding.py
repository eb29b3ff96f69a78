"""File / --fast mode frame partitioning across GPUs (multi-GPU): contiguous segments, each
preceded by the temporal warm-up frames its first output depends on.  No collective is involved: the
temporal stage only looks back `warmup` unblended frames (src/main.cpp:277: 1 frame;
TemporalConsistency with buffer_size b: b frames - every buffered frame is blended,
src/temporal_consistency.cpp:127 with :97, :162-169)."""


def segment(n_frames, world, rank):
    """[start, stop) of rank's contiguous segment: floor(k*N/G) boundaries."""
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def warmup_frames(temporal_mode, buffer_size=3):
    """temporal_mode: 'none' | 'main' | 'frames' (VP_TEMPORAL_*)."""
    if temporal_mode in (0, "none"):
        return 0
    if temporal_mode in (1, "main"):
        return 1
    return max(int(buffer_size), 0)


def plan(n_frames, world, rank, temporal_mode, buffer_size=3):
    """Returns (first_frame_to_feed, n_warmup, start, stop): frames [first, start) are fed with no
    output (vp_process_device(d_dst=NULL) / Upscaler::warmup), frames [start, stop) produce output."""
    start, stop = segment(n_frames, world, rank)
    w = min(warmup_frames(temporal_mode, buffer_size), start)
    return start - w, w, start, stop

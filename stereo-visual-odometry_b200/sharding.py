"""Frame sharding for throughput mode: extraction, stereo matching and LK are independent per stereo
frame; only the temporal match / track needs frames t-1 and t.  A sequence of F frames has F-1 units (unit u uses frames
u and u+1).  Rank r of W gets a contiguous run of units plus the one frame before it, so chunks overlap by exactly one
frame and no collective is needed on the data path."""


def shard_units(n_frames, world, rank):
    """-> (first_frame, n_frames_for_rank): frames [first, first+n) carry units [first, first+n-1)."""
    units = max(n_frames - 1, 0)
    u0, u1 = units * rank // world, units * (rank + 1) // world
    if u1 <= u0:
        return u0, 0
    return u0, u1 - u0 + 1


def batches(first_frame, n_frames, max_frames):
    """Split a rank's frame run into svo_frontend_batch calls of at most max_frames frames."""
    out, f = [], first_frame
    while f < first_frame + n_frames:
        n = min(max_frames, first_frame + n_frames - f)
        out.append((f, n))
        f += n
    return out

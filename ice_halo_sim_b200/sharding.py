"""Ray-index sharding for multi-GPU runs.

Root rays are i.i.d. and the engine's RNG is counter-based on the global ray index, so R ranks tracing
disjoint contiguous index ranges of a single-population layer generate exactly the roots of a 1-rank run (and,
with the gate / transit streams keyed by the same index, statistically independent continuation layers); the only
exchange is one sum reduce of the XYZ accumulator at frame end. With several populations the population of a
global index depends on the session split (PartitionCrystalRayNum works on session-local counts), so such frames
are statistically, not ray-for-ray, equal across world sizes.
"""


def shard_range(total, rank, world):
    """Contiguous [begin, end) of `total` items owned by `rank` of `world` (sizes differ by at most 1)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def session_plan(total_rays, rank, world, session_rays, index_base=0):
    """[(global_index_of_first_ray, ray_count)] sessions covering this rank's shard."""
    begin, end = shard_range(total_rays, rank, world)
    out = []
    while begin < end:
        n = min(int(session_rays), end - begin)
        out.append((index_base + begin, n))
        begin += n
    return out

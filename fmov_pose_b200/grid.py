"""Grid-partitioned dense SDF query (BASELINE config C5): the res^3 grid of `extract_fields`
(models/renderer.py:9-37) is split into contiguous slabs of x-planes, one per rank; every rank evaluates its slab with
one `fmov_sdf_query_grid` launch and the slabs are gathered (NCCL all-gather over NVLink, or gloo on CPU in the tests)
into the full `u` grid that marching cubes consumes (models/renderer.py:43)."""
import torch


def slab_of(resolution, world, rank):
    """(first_plane, n_planes) of `rank`: the first `resolution % world` ranks take one extra plane."""
    base, extra = divmod(int(resolution), int(world))
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n


def gather_slabs(local, resolution, group=None):
    """local: this rank's [n_planes * res^2] slab -> the full [res^3] grid on every rank.  Slabs differ by at most one
    plane, so they are padded to the largest and all-gathered in one collective."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if group is not None else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    plane = resolution * resolution
    n_max = slab_of(resolution, world, 0)[1]
    first, n = slab_of(resolution, world, rank)
    assert local.numel() == n * plane, (local.numel(), n, plane)
    if n < n_max:
        local = torch.cat([local, local.new_zeros((n_max - n) * plane)])
    buf = local.new_empty(world * n_max * plane)
    dist.all_gather_into_tensor(buf, local.contiguous(), group=group)
    if resolution % world == 0:
        return buf
    parts = [buf[r * n_max * plane: r * n_max * plane + slab_of(resolution, world, r)[1] * plane] for r in range(world)]
    return torch.cat(parts)


def extract_fields_sharded(renderer, bound_min, bound_max, resolution, group=None, to_host=False, precise="act"):
    """u = -sdf on the res^3 grid with the x-planes partitioned across the ranks of `group`.
    -> [res,res,res] tensor on every rank (device; pinned host memory when to_host: what validate_mesh hands to
    marching cubes, exp_runner.py:1630-1640)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0
    first, n = slab_of(resolution, world, rank)
    plane = resolution * resolution
    local = renderer.extract_fields(bound_min, bound_max, resolution, first=first * plane, count=n * plane, precise=precise)
    u = gather_slabs(local, resolution, group).view(resolution, resolution, resolution)
    if to_host:
        host = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
        host.copy_(u, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host
    return u

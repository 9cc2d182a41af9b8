"""Multi-GPU plumbing: one process per GPU (torch.distributed), scene replicated, work sharded, ONE reduce of the fp32
film.  The reference has no distributed path (its only parallelism is rayon over rows,
src/renderer.rs:87-90); (pixel, sample) independence is what makes the shard boundaries invisible: the Philox
stream is keyed on the GLOBAL pixel and sample index, so the reduced film equals the single-GPU film up to fp32
summation order.

`render_sharded` is backend-agnostic on purpose: on a B200 the per-rank render is CoreScene.render_accumulate into a
CUDA tensor reduced over NCCL/NVLink; the gloo CPU tests drive the very same sharding / reduce logic with the oracle
as the per-rank renderer.
"""
import torch
import torch.distributed as dist


def sample_range(spp, rank, world):
    """Contiguous, balanced split of [0, spp) — ranks differ by at most one sample."""
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard(mode, spp, rank, world):
    """-> (sample_begin, sample_end, tile_mod, tile_rem) for this rank.

    "samples": every rank renders the whole frame for its sample range (perfect load balance, any image size).
    "tiles":   every rank renders all samples of the 32x32 tiles with index % world == rank (interleaved, so sky and
               geometry tiles mix on every rank)."""
    if mode == "samples":
        b, e = sample_range(spp, rank, world)
        return b, e, 0, 0
    if mode == "tiles":
        return 0, spp, world, rank
    raise ValueError(mode)


def render_sharded(render_fn, spp, mode="samples", group=None, dst=0):
    """render_fn(sample_begin, sample_end, tile_mod, tile_rem) -> 1-D float32 tensor: this rank's radiance SUM film
    (W*H*3).  Returns the mean-radiance film (sum / spp) on rank `dst`, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if mode == "samples" and world > spp:  # decided from (spp, world) alone, so every rank raises — none waits in the reduce
        raise ValueError("more ranks than samples: use mode='tiles'")
    b, e, tm, tr = shard(mode, spp, rank, world)
    film = render_fn(b, e, tm, tr)
    if world > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)  # the single exchange step of the path
    if rank != dst:
        return None
    return film * (1.0 / spp)  # renderer.rs:85,103


def core_render_fn(pt, core_scene, camera, base_settings, accum, stream_ptr=None, stats_out=None):
    """Per-rank renderer on a B200: zero the device film, ptc_render_accumulate into it on torch's current stream."""
    def fn(sample_begin, sample_end, tile_mod, tile_rem):
        st = pt.RenderSettings.from_buffer_copy(base_settings)
        st.sample_begin, st.sample_end, st.tile_mod, st.tile_rem = sample_begin, sample_end, tile_mod, tile_rem
        accum.zero_()
        s = core_scene.render_accumulate(camera, st, accum.data_ptr(),
                                         stream_ptr if stream_ptr is not None else torch.cuda.current_stream().cuda_stream)
        if stats_out is not None:
            stats_out.append(s)
        return accum
    return fn

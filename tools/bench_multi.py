#!/usr/bin/env python3
"""Multi-GPU throughput of one BASELINE configuration, sharded the way SURVEY.md 8(e) says: scene replicated, work split
by interleaved 32x32 tiles or by sample range, ONE NCCL reduce of the fp32 film to rank 0, then the film resolve — all
inside the timed region.  bench.py stays the headline (C2, weak scaling over samples); this tool produces the 1/2/4/8
GPU table for the synthetic 2 M-triangle scene (config C5) and any other config.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_multi.py --config c5 --mode tiles --scaling strong --spp 64 [--steps 3]

  strong: the job is fixed (the config's frame at --spp samples); value = its paths / max-over-ranks time
  weak:   every rank adds --spp samples of the full frame (mode samples) — the job grows with N
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def scene_for(pt, name):
    sc = os.path.join(ROOT, "scenes")
    if name == "c5":
        return pt.synthetic_scene(cells=1000), "C5 synthetic 2,000,000-triangle height field 3840x2160, depth 16"
    if name == "c2":
        return pt.load_scene_from_json(os.path.join(sc, "semesterbild.json")), "C2 semesterbild.json 800x600, depth 30"
    if name == "c4":
        return pt.load_scene_from_json(os.path.join(sc, "veach-mis", "scene.json")), "C4 veach-mis 1280x720, depth 16"
    raise SystemExit("unknown config " + name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5")
    ap.add_argument("--mode", default="tiles", choices=["tiles", "samples"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--pool", type=int, default=3 << 22)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import ptload
    pt = ptload.load()
    from importlib import import_module
    D = import_module("raytracer-rust_b200.dist")

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)  # NCCL's banner must not land on stdout
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    scene, label = scene_for(pt, a.config)
    w, h, _, depth = scene.settings
    cs = scene.to_core().commit(local)
    total_spp = a.spp * world if a.scaling == "weak" else a.spp
    mode = "samples" if a.scaling == "weak" else a.mode
    if mode == "samples" and total_spp < world:
        raise SystemExit("more ranks than samples")
    accum = torch.zeros(h * w * 3, dtype=torch.float32, device=dev)
    out_u32 = torch.zeros(h * w, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    base = scene.render_settings(spp=total_spp, seed=0, pool_paths=a.pool)
    stats = []
    fn = D.core_render_fn(pt, cs, scene.camera, base, accum, stream.cuda_stream, stats)
    b, e, tm, tr = D.shard(mode, total_spp, rank, world)

    def step():
        fn(b, e, tm, tr)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            pt._ck(pt.core().ptc_resolve_device(accum.data_ptr(), h * w, 1.0 / total_spp, out_u32.data_ptr(), stream.cuda_stream))

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    del stats[:]
    tot = 0.0
    for _ in range(a.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot], dtype=torch.float64, device=dev)
    cnt = torch.tensor([sum(s.paths for s in stats), sum(s.rays for s in stats), max(s.render_ms for s in stats)], dtype=torch.float64, device=dev)
    mx = cnt[2:3].clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt[:2], op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"config": label, "n_gpus": world, "scaling": a.scaling, "sharding": mode, "spp_total": total_spp,
                          "steps": a.steps, "ms_per_step": ms / a.steps, "mpaths_per_s": float(cnt[0].item()) / ms / 1e3,
                          "mrays_per_s": float(cnt[1].item()) / ms / 1e3, "slowest_rank_render_ms": float(mx.item()),
                          "film_reduce_bytes": h * w * 3 * 4 if world > 1 else 0, "pool_paths": a.pool}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

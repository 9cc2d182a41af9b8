#!/usr/bin/env python3
"""bench.py — the headline benchmark: BASELINE.json's metric (Mpaths/s, with Mrays/s beside it) on config C2,
`semesterbild.json` at its native 800x600 / 256 spp / 30 bounces, on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full render of the frame (render_scene, src/renderer.rs:67-123 of the reference, timed like the
reference times it: renderer.rs:82,109 — scene load, BVH build and PNG encode are outside).

  value  device-resident: scene resident in HBM, film accumulated in HBM (ptc_render_accumulate) and, for N > 1, one
         NCCL reduce of the fp32 film to rank 0 + the film resolve, all inside the timed region.
  e2e    the reference-facing call with HOST buffers: ptc_render_u32, i.e. `render_scene` returning its Vec<u32>
         (film resolved on the device, packed image copied to the host).
N > 1 is weak scaling over samples: every rank renders the full frame at the config's spp with its own sample
range [rank*spp, (rank+1)*spp) of an N*spp job (Philox is keyed on the global sample index, so the reduced film is
the N*spp image); value = all ranks' paths / max-over-ranks time.

`--impl reference` times the CPU restatement of the reference (oracle/, the reference's own ChaCha stream, all host
threads) on a bounded sample of the same frame.  The Rust reference itself cannot be built in this image (no
cargo/rustc).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCENE = os.path.join(ROOT, "scenes", "semesterbild.json")
WORKLOAD = "C2 semesterbild.json 800x600, 256 spp, max_bounces 30 (3 cubes, 4748-triangle mesh, glass sphere; sky-lit)"
METRIC = "Mpaths/s (semesterbild 800x600x256spp, depth 30)"


def measured_traffic():
    """DRAM bytes per extend ray from the committed ncu capture (dram__bytes_read.sum + dram__bytes_write.sum of the three
    extend kernels of one full-pool iteration, profiles/r1_extend_traffic.json), and that capture's per-kernel figures."""
    p = os.path.join(ROOT, "profiles", "r1_extend_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dram_bytes_per_ray"]), d.get("ncu_per_kernel")
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args):
    """CPU arm: the oracle (C++ restatement of the reference's renderer) on all host threads, ChaCha stream."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ptload
    pt = ptload.load()
    from bindings import OracleScene, RNG_CHACHA
    scene = pt.load_scene_from_json(SCENE)
    w, h, spp, depth = scene.settings
    orc = OracleScene(scene)
    cores = os.cpu_count() or 1
    # calibrate: one sample per pixel, then size a step to ~6 s of CPU work
    _, st = orc.render(scene.camera, w, h, 1, depth, rng_mode=RNG_CHACHA)
    per_spp = max(st.seconds, 1e-3)
    spp_step = int(max(1, min(spp, round(6.0 / per_spp))))
    for _ in range(args.warmup):
        orc.render(scene.camera, w, h, spp_step, depth, rng_mode=RNG_CHACHA)
    paths = rays = 0
    secs = 0.0
    for _ in range(args.steps):
        _, st = orc.render(scene.camera, w, h, spp_step, depth, rng_mode=RNG_CHACHA)
        paths += st.paths
        rays += st.rays
        secs += st.seconds
    value = paths / secs / 1e6
    sample = f"{w}x{h} at {spp_step} spp per step (of {spp}), depth {depth}; Mpaths/s does not depend on spp"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mrays_per_s": rays / secs / 1e6,
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the reference CPU path (oracle/), row-parallel like rayon, ChaCha12 per-row "
                                 "RNG; `cargo run --release` is impossible here (no Rust toolchain)"},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pool", type=int, default=3 << 22, help="path-pool slots (12 Mi: C2 on B200 4 Mi 48.7 ms, 8 Mi 45.9, 12 Mi 45.0, 16 Mi 44.9)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import ptload
    pt = ptload.load()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or pt.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the path tracer has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at communicator creation; stdout carries exactly one JSON line,
        # so park fd 1 on stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    scene = pt.load_scene_from_json(SCENE)
    w, h, spp, depth = scene.settings
    cs = scene.to_core().commit(local)
    cam = scene.camera
    stream = torch.cuda.current_stream()
    accum = torch.zeros(h * w * 3, dtype=torch.float32, device=dev)
    out_u32 = torch.zeros(h * w, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    total_spp = spp * world

    def settings(flags=0):
        return scene.render_settings(spp=total_spp, sample_begin=rank * spp, sample_end=(rank + 1) * spp, seed=0,
                                     pool_paths=args.pool, flags=flags)

    def step_device(flags=0):
        accum.zero_()
        st = cs.render_accumulate(cam, settings(flags), accum.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            pt._ck(pt.core().ptc_resolve_device(accum.data_ptr(), h * w, 1.0 / total_spp, out_u32.data_ptr(), stream.cuda_stream))
        return st

    host_rgb = np.empty((h, w, 3), np.float32)

    def step_e2e():
        # the reference-facing call: host buffers in and out
        if world == 1:
            _, st = cs.render_u32(cam, settings())  # render_scene -> Vec<u32>, resolved on the device
            return st
        st = step_device()
        if rank == 0:
            out_u32.cpu()
        return st

    def timed(fn, n):
        tot_ms, stats = 0.0, []
        for _ in range(n):
            flush.fill_(1)  # evict L2 between timed iterations (untimed)
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            stats.append(fn())
            e1.record(stream)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            # the wavefront loop synchronises its own stream before returning, so wall and event time agree; e2e steps
            # that use the library's internal stream are only visible to the wall clock
            tot_ms += max(e0.elapsed_time(e1), wall if fn is step_e2e else 0.0)
        t = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stats

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # `value`: the production path, exactly K steps.  Then K more steps of the same work with PTC_FLAG_TIMING, which
    # brackets every stage launch with CUDA events on the launching stream (a few per cent slower): these give the
    # per-kernel share and the roofline's launch durations.  Clocks are sampled across both.
    ms_total, vstats = timed(lambda: step_device(0), args.steps)
    _, stats = timed(lambda: step_device(pt.FLAG_TIMING), args.steps)
    clocks = sampler.stop() if rank == 0 else None

    paths_rank = sum(s.paths for s in vstats)
    rays_rank = sum(s.rays for s in vstats)
    cnt = torch.tensor([paths_rank, rays_rank], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    paths_all, rays_all = float(cnt[0].item()), float(cnt[1].item())
    value = paths_all / ms_total / 1e3
    launches = sum(s.kernel_launches for s in vstats) + args.steps * 2  # + memset + resolve

    for _ in range(2):
        step_e2e()
    ms_e2e, e2e_stats = timed(step_e2e, args.steps)
    e2e_paths = torch.tensor([sum(s.paths for s in e2e_stats)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(e2e_paths, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_paths.item()) / ms_e2e / 1e3
    if world == 1:
        d2h = h * w * 4 + 64 * (1 + spp // 8)  # the packed image + the control-block snapshots the host polls
        h2d = 256                               # control block + launch parameters (the scene is resident)
    else:
        d2h, h2d = h * w * 4, 256

    # ---- roofline of the dominant kernel (extend), rank 0 view
    ext_ms = sum(s.extend_ms for s in stats)
    ext_launches = sum(s.extend_launches for s in stats)
    shade_ms = sum(s.shade_ms for s in stats)
    render_ms = sum(s.render_ms for s in stats)
    c = cs.render(cam, scene.render_settings(spp=8, seed=0, flags=pt.FLAG_COUNTERS))[1]  # instrumented build, untimed
    nodes_per_ray, tris_per_ray = c.nodes_visited / c.rays, c.tris_tested / c.rays
    objs = scene.objects
    cost = {pt.OBJ_SPHERE: 25, pt.OBJ_PLANE: 14, pt.OBJ_QUAD: 35, pt.OBJ_CUBE: 80, pt.OBJ_MESH: 64}
    analytic_instr = sum(cost[o.type] for o in objs)
    bytes_per_ray = 32 + 16 + 80.0 * nodes_per_ray + 48.0 * tris_per_ray
    instr_per_ray = analytic_instr + 170.0 * nodes_per_ray + 45.0 * tris_per_ray
    hbm_peak, sm_max_mhz, peak_kind = peaks()
    rays_per_launch = rays_rank / max(1, ext_launches)
    ext_s = ext_ms / 1e3
    achieved_gbs = bytes_per_ray * rays_rank / ext_s / 1e9 if ext_s > 0 else 0.0
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max_mhz
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = sms * 128 * sm_mhz * 1e6 / 1e12  # T instr/s at the clock sampled under load
    achieved_tinstr = instr_per_ray * rays_rank / ext_s / 1e12 if ext_s > 0 else 0.0

    dram_per_ray, ncu_kernels = measured_traffic()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"samples x{world}" if world > 1 else "single GPU",
                       "spp_total": total_spp, "l2": "256 MiB flush write between timed steps", "pool_paths": args.pool,
                       "rng": "Philox4x32-10 keyed (pixel, sample, bounce)"},
            "mrays_per_s": rays_all / ms_total / 1e3, "rays_per_path": rays_all / paths_all,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernel_share": {"k_extend_pre": sum(s.pre_ms for s in stats) / render_ms,
                             "k_traverse": sum(s.traverse_ms for s in stats) / render_ms,
                             "k_extend_post": sum(s.post_ms for s in stats) / render_ms,
                             "k_shade (incl. path regeneration)": shade_ms / render_ms,
                             "launch gaps": max(0.0, 1.0 - (ext_ms + shade_ms) / render_ms)},
            "roofline": {"kernel": "extend stage = k_extend_pre + k_traverse + k_extend_post (one logical kernel, timed together)", "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak,
                         "traffic": dram_per_ray * rays_per_launch if dram_per_ray else None,
                         "traffic_note": "bytes per launch of the stage: DRAM bytes per ray of the committed ncu capture "
                                         "(profiles/r1_extend_traffic.json) x rays per launch; below the algorithmic bytes because "
                                         "the BVH nodes and triangles are served from L1/L2",
                         "algorithmic_bytes_per_launch": bytes_per_ray * rays_per_launch, "peak_source": peak_kind,
                         "algorithmic_bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "rays_per_launch": rays_per_launch, "avg_launch_ms": ext_ms / max(1, ext_launches),
                         "extend_grays_per_s": rays_rank / ext_s / 1e9 if ext_s > 0 else 0.0,
                         "note": "algorithmic bytes = 32 B ray + 16 B hit + 80 B per wide node popped + 48 B per triangle tested (SURVEY.md 8d); "
                                 "the BVH + triangle working set is < 1 MB (L1/L2 resident), so HBM is the schema's bound, not the binding one: "
                                 "the stage is instruction-issue bound (profiles/), see roofline_fp32"},
            "roofline_fp32": {"bound": "fp32 issue", "achieved": achieved_tinstr, "peak": fp32_peak, "unit": "Tinstr/s",
                              "frac": achieved_tinstr / fp32_peak, "algorithmic_instr_per_ray": instr_per_ray,
                              "peak_source": f"{sms} SMs x 128 lanes x {sm_mhz:.0f} MHz sampled under load",
                              "ncu": ncu_kernels,
                              "note": "algorithmic count with FMA = 1 and no divergence; what the kernels EXECUTE (unfused IEEE "
                                      "arithmetic for bit-exact hit records, incoherent rays) keeps the issue slots busy 74 % / 60 % / "
                                      "47 % of the time in pre / traverse / post (ncu, profiles/r1_v4_stages_ncu_summary.txt)"},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(pt, scene)
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline(pt, scene):
    from bindings import OracleScene, RNG_CHACHA
    w, h, spp, depth = scene.settings
    orc = OracleScene(scene)
    _, st = orc.render(scene.camera, w, h, 1, depth, rng_mode=RNG_CHACHA)
    n = int(max(1, min(spp, round(12.0 / max(st.seconds, 1e-3)))))
    _, st = orc.render(scene.camera, w, h, n, depth, rng_mode=RNG_CHACHA)
    return {"value": st.paths / st.seconds / 1e6, "unit": "Mpaths/s", "cores": os.cpu_count() or 1, "kind": "port",
            "mrays_per_s": st.rays / st.seconds / 1e6,
            "sample": f"{w}x{h} at {n} spp (of {spp}), depth {depth}, {st.seconds:.1f} s of CPU work; C++ restatement of the reference "
                      "(oracle/), row-parallel, ChaCha12 per-row RNG as the reference"}


if __name__ == "__main__":
    sys.exit(main())

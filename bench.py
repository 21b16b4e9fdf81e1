#!/usr/bin/env python
"""bench.py — headline measurement of the wavefront path-tracing hot path (BASELINE.json / SURVEY.md §8d).

  python bench.py --gpus N --steps K --warmup W [--workload instanced10m_4k|cornell_1080p|sky10m_4k|build50m|build10m|build100k] [--impl reference]

A step is one frame (1 sample per pixel) of the wavefront pipeline: generate -> closest-hit trace -> shade (+NEE) ->
{closest-hit trace, any-hit shadow trace} per bounce -> accumulate.  Metric: Mrays/s = (extension + shadow rays traced) /
device time / 1e6, as SURVEY.md §8(d) defines it; spp/s (frames/s) rides along in `spp_per_s`.

N > 1 (launched by torchrun, one rank per GPU): the scene is replicated, rank g renders its own block of K frame indices
(sample partition, weak scaling) and the float accumulation buffers are summed with one NCCL all-reduce inside the timed
region.  Timing is CUDA events on the stream the kernels run on; the reported time is the max over ranks.

--impl reference runs the UNMODIFIED reference CUDA kernels (oracle/_ref/libnexus_ref.so, compiled from /root/reference by
oracle/Makefile with the reference's own flags) through a headless replay of PathTracer::Render on the same GPU, same scene,
same metric: the reference has no CPU implementation of this path, its implementation IS CUDA, and north_star names "the
reference's own CUDA renderer on the same B200" as the baseline.  No product code (libnexus_b200.so) is loaded on that arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

DEFAULT_TRACE_MODE = "lane"      # what nx_ctx_create selects when NX_TRACE_MODE is unset (nx_common.cuh: trace_mode)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration north_star's target is quoted on; fits one GPU (about 2.5 GB resident)
    "instanced10m_4k": dict(res=(3840, 2160), desc="10M-triangle instanced scene (1024 BLAS x 9798 tris, 1026 instances, TLAS), OpenPBR "
                            "dielectric/metal/translucent, NEE+MIS, pathLength 8, 3840x2160, 1 spp/step"),
    # BASELINE.json configs[1]
    "cornell_1080p": dict(res=(1920, 1080), desc="Cornell box (32 triangles, 8 instances, diffuse + 35x area light), NEE+MIS, pathLength 10, 1920x1080, 1 spp/step"),
    # BASELINE.json configs[4]'s scene (8 GPUs: `--gpus 8`): configs[2]'s geometry lit only by a procedural HDR sky with a sun disc
    "sky10m_4k": dict(res=(3840, 2160), desc="10M-triangle instanced scene (1024 BLAS, 1026 instances), no area light, procedural 4096x2048 RGBA32F HDR sky + sun disc, "
                      "pathLength 8, 3840x2160, 1 spp/step"),
    # BASELINE.json configs[3]: the NexusBVH benchmark mesh (Test/src/Main.cpp:31-64) at 50M triangles; metric Mprims/s
    "build50m": dict(res=None, n=50_000_000, desc="H-PLOC BVH2 build + CWBVH8 collapse of the NexusBVH benchmark mesh (random small triangles on a 1000^3 lattice), "
                     "50,000,000 triangles, 32-bit Morton keys (prioritizeSpeed), 1 build/step"),
    "build10m": dict(res=None, n=10_000_000, desc="as build50m with 10,000,000 triangles (the size NexusBVH's README quotes)"),
    # BASELINE.json configs[0]: the reference's CPU-runnable case.  Our arm builds the same mesh on the GPU; the CPU path (binned-SAH
    # BVH2 + SAH-optimal BVH8 collapse, oracle/oracle_sah.cpp) is timed beside it on the host cores as cpu_baseline
    "build100k": dict(res=None, n=100_352, mesh="uv_sphere", desc="procedurally tessellated UV sphere, 224 x 224 x 2 = 100,352 triangles: H-PLOC BVH2 build + CWBVH8 collapse on the GPU; "
                      "cpu_baseline = CPU binned-SAH BVH2 build + SAH-optimal BVH8 collapse of the same mesh on the host cores"),
    # reduced variant for quick functional checks (not a bench line)
    "instanced_small": dict(res=(640, 360), desc="64-BLAS reduced instanced scene, 640x360 (functional check only)"),
}


def make_desc(workload):
    from nexus_b200 import scenes
    if workload == "instanced10m_4k":
        d = scenes.instanced_scene(n_blas=1024, n_instances=1024, path_length=8)
    elif workload == "cornell_1080p":
        d = scenes.cornell_box(path_length=10)
    elif workload == "sky10m_4k":
        d = scenes.instanced_scene(n_blas=1024, n_instances=1024, path_length=8)
        # no area light: the emissive quad keeps its place in the instance list (same TLAS) but emits nothing
        d["materials"][1].emissionColor = (0.0, 0.0, 0.0); d["materials"][1].intensity = 0.0
        d["hdr"] = scenes.procedural_sky()
        d["settings"].backgroundIntensity = 1.0
    elif workload == "instanced_small":
        d = scenes.instanced_scene(n_blas=64, n_instances=64, nu=24, nv=24, path_length=8)
    else:
        raise SystemExit(f"unknown workload {workload}")
    return scenes.with_triangle_data(d)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, reasons, mx, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            # under load = samples at or above the median power of the upper half
            out.update(sm_mhz=float(np.median(sm[len(sm) // 4:])), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(work, kind):
    """closest hit 24 + 20 B/ray, shadow 44 B/ray; 80 B per node visited, 4 + 36 B per triangle tested,
    144 B per instance entered."""
    per_ray = 44
    return per_ray * work["rays"] + 80 * work["nodes"] + 40 * work["tris"] + 144 * work["insts"]


def kernel_source_sha():
    """Hash of the CUDA sources: a committed ncu figure is only reported for the code it was measured on."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "nexus_b200", "csrc", "*.cu*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


NCU_METRICS = "smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max"


def ncu_counters(workload, kernel_regex, launches_per_step, extra_env=None):
    """Hardware counters of one step's launches of a kernel, measured NOW: a child `bench.py --count-pass` (scene set-up, one warm
    step, one counted step) runs under ncu with five raw metrics.  Instruction and DRAM-byte counts do not depend on the replay;
    no time is taken from this pass.  Returns per-launch averages over the last step's launches, or None (ncu missing / not permitted)."""
    import csv
    import shutil
    if not shutil.which("ncu"):
        return None
    log = tempfile.mktemp(suffix=".csv")
    cmd = ["ncu", "--metrics", NCU_METRICS, "--clock-control", "none", "-k", "regex:" + kernel_regex, "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--workload", workload, "--count-pass"]
    env = dict(os.environ); env.update(extra_env or {})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600, check=True)
        rows = [r for r in csv.reader(l for l in open(log) if l.startswith('"'))]
    except (OSError, subprocess.SubprocessError):
        return None
    finally:
        if os.path.exists(log):
            os.unlink(log)
    if not rows:
        return None
    hdr = rows[0]
    iid, iname, ival = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        try:
            per.setdefault(int(r[iid]), {})[r[iname]] = float(r[ival].replace(",", ""))
        except (ValueError, IndexError):
            continue
    ids = sorted(per)[-launches_per_step:]
    if len(ids) < launches_per_step:
        return None
    out = {m: sum(per[i].get(m, 0.0) for i in ids) / len(ids) for m in NCU_METRICS.split(",")}
    out["launches"] = len(ids)
    return out


def run_count_pass(args):
    """Child of ncu_counters: the workload's scene, one warm step, one counted step; prints nothing."""
    import nexus_b200 as nx
    from nexus_b200 import scenes
    ctx = nx.Context(0)
    if args.workload.startswith("build"):
        import torch
        wl = WORKLOADS[args.workload]
        dev_t = torch.from_numpy(build_mesh(wl)).cuda()
        for _ in range(2):
            nx.BuildBVH8Device(ctx, dev_t.data_ptr(), wl["n"], 1, True).Free()
        ctx.synchronize()
        return
    desc = make_desc(args.workload)
    res = WORKLOADS[args.workload]["res"]
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    pt.Render(scene, frames=1, firstFrame=1); ctx.synchronize()
    pt.Render(scene, frames=1, firstFrame=2); ctx.synchronize()


# ----------------------------------------------------------------------------------------------- our arm ----
def run_ours(args):
    import torch
    import torch.distributed as dist
    import nexus_b200 as nx
    from nexus_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, args.warmup
    wl = WORKLOADS[args.workload]
    res = wl["res"]

    ctx = nx.Context(local)          # raises when the CUDA library or a GPU is missing: there is no fallback
    ctx_trace_mode = os.environ.get("NX_TRACE_MODE", DEFAULT_TRACE_MODE)
    desc = make_desc(args.workload)
    t_scene = time.time()
    # N > 1: the BLAS builds are sharded (rank g builds meshes g, g + N, ...; one NCCL all-gather hands every rank every BLAS,
    # SURVEY.md 8(e)); the TLAS is built by every rank.  --replicated-build makes every rank build everything.
    # With instance merging (the default) the meshes of this scene end up in ONE world-space BLAS that every rank builds itself (a few
    # milliseconds); --sharded-build turns merging off and shards the per-mesh BLAS builds over the ranks instead.
    sharded = world > 1 and len(desc["meshes"]) > 1 and args.sharded_build
    if sharded:
        ctx.SetInstanceMerging(False)
    if world > 1:                                      # NCCL communicator set-up (lazy, ~1 s) is not a scene cost: do it before the clock starts
        dist.all_reduce(torch.zeros(1, device="cuda")); torch.cuda.synchronize()
        t_scene = time.time()
    if sharded:
        from nexus_b200.multigpu import build_scene_sharded
        scene = build_scene_sharded(ctx, desc, res)
    else:
        scene = scenes.build(ctx, desc, res)
    ctx.synchronize()
    t_scene = time.time() - t_scene
    pt = nx.PathTracer(ctx, res)
    stream = torch.cuda.ExternalStream(int(nx.lib().nx_ctx_stream(ctx._h)), device=torch.device("cuda", local))
    from nexus_b200.multigpu import accumulation_tensor, frame_block, reduce_accumulation
    acc_t = accumulation_tensor(pt, torch.device("cuda", local)) if world > 1 else None   # zero-copy view for NCCL

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up: W untimed steps (also warms NCCL)
    pt.Render(scene, frames=max(W, 1), firstFrame=1)
    if world > 1:
        with torch.cuda.stream(stream):
            # the exact sequence of the timed region: the float64 checksum kernel is loaded on first use (13 ms measured inside the
            # timed region of a 2-GPU run when only the all-reduce had been warmed)
            acc_t.sum(dtype=torch.float64)
            reduce_accumulation(acc_t, max(W, 1))
    barrier()
    pt.ResetFrameNumber()

    # ---- timed region: exactly K steps per rank, kernel events on, + the NCCL reduce for N > 1
    first = frame_block(rank, world, K)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    pt.SetProfiling(events=False, work=False)   # no per-launch events inside the headline region: they cost 7 % on the small Cornell launches
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    wall0 = time.time()
    e0.record(stream)
    pt.Render(scene, frames=K, firstFrame=first)
    e1.record(stream)
    local_sum = None
    if world > 1:
        with torch.cuda.stream(stream):
            local_sum = acc_t.sum(dtype=torch.float64)          # checksum of this rank's contribution (100 MB read, inside the timed region)
            pt.SetAccumulatedFrames(reduce_accumulation(acc_t, K))
    e2.record(stream)
    barrier()
    wall = time.time() - wall0
    reduce_check = None
    if world > 1:
        # multi-GPU correctness, in the run the driver records: the reduced buffer must be the sum of what the ranks rendered
        with torch.cuda.stream(stream):
            reduced_sum = acc_t.sum(dtype=torch.float64)
            dist.all_reduce(local_sum)
        barrier()
        want, got = float(local_sum), float(reduced_sum)
        rel = abs(got - want) / max(abs(want), 1e-30)
        assert rel < 1e-5, f"reduced accumulation checksum {got} != sum of the per-rank checksums {want} (relative {rel:.2e})"
        reduce_check = {"sum_of_rank_checksums": want, "reduced_checksum": got, "rel_err": rel}
    clocks = sampler.stop() if rank == 0 else None
    ms_total, ms_reduce = e0.elapsed_time(e2), e1.elapsed_time(e2)
    st = pt.Stats()
    rays_local = st["extension_rays"] + st["shadow_rays"]
    agg = torch.tensor([ms_total, ms_reduce], device="cuda", dtype=torch.float64)
    cnt = torch.tensor([rays_local, st["extension_rays"], st["shadow_rays"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, ms_reduce = float(agg[0]), float(agg[1])
    rays_all = float(cnt[0])
    value = rays_all / (ms_total * 1e-3) / 1e6
    mean_radiance = float(pt.ReadAccumulation().mean()) if rank == 0 else 0.0

    # ---- roofline of the dominant kernel (closest-hit traversal): the same K frames again with a CUDA event pair around every
    # launch (on the stream it is launched on), then the algorithmic bytes from a counted, untimed replay of the same frames
    pt.SetProfiling(events=True, work=False)
    pt.ResetFrameNumber()
    pt.Render(scene, frames=K, firstFrame=first)
    ctx.synchronize()
    prof = pt.Profile()
    pt.SetProfiling(events=False, work=True)
    pt.ResetFrameNumber()
    pt.Render(scene, frames=K, firstFrame=first)
    ctx.synchronize()
    workp = pt.Profile()
    pt.SetProfiling(events=False, work=False)
    tc = prof["trace_closest"]
    cw, aw = workp["closest_work"], workp["any_work"]
    bytes_closest = algorithmic_bytes(cw, "closest")
    hbm, hbm_src = peaks()
    avg_launch_ms = tc["ms"] / max(tc["launches"], 1)
    achieved = bytes_closest / max(tc["launches"], 1) / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    # What bounds the kernel is instruction issue, not HBM (ncu: 0.4-0.9 GB of DRAM traffic per launch against 6.9 GB requested, L2
    # throughput under 20 %): the roofline is thread-instructions per second against SMs x 4 schedulers x 32 lanes x SM clock.  The
    # instruction and DRAM-byte counts come from an ncu pass over one step of THIS build, run now (ncu_counters); the time is the
    # live CUDA-event time above.  The request-byte rate (SURVEY.md 8d's formula) is kept beside it under "hbm".
    launches_per_step = max(tc["launches"] // K, 1)
    counters = None
    if rank == 0 and world == 1 and not args.no_ncu:
        counters = ncu_counters(args.workload, "trace_closest", launches_per_step)
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    sms = ctx.sm_count
    issue_peak = sms * 4 * 32 * sm_mhz * 1e6 / 1e9                      # G thread-instructions / s
    hbm_part = {"achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4), "peak_source": hbm_src,
                "what": "request bytes (44 B/ray + 80 B/node + 40 B/triangle + 144 B/instance entry, SURVEY.md 8d) / live launch time; most of it is served by L1/L2",
                "algorithmic_bytes_per_launch": int(bytes_closest / max(tc["launches"], 1))}
    if counters:
        tinst, winst = counters["smsp__thread_inst_executed.sum"], counters["smsp__inst_executed.sum"]
        a_issue = tinst / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
        traffic = int(counters["dram__bytes_read.sum"] + counters["dram__bytes_write.sum"])
        hbm_part["dram_frac_of_peak"] = round(traffic / (avg_launch_ms * 1e-3) / 1e9 / hbm, 4) if avg_launch_ms > 0 else None
        roofline = {"bound": "issue", "kernel": "trace_closest (%s loop)" % ctx_trace_mode, "achieved": round(a_issue, 1), "peak": round(issue_peak, 1), "unit": "Gthread-inst/s",
                    "frac": round(a_issue / issue_peak, 4), "traffic": traffic,
                    "peak_source": f"{sms} SMs x 4 schedulers x 32 lanes x {sm_mhz:.0f} MHz (SM clock sampled under load in the timed region)",
                    "warp_inst_per_launch": int(winst), "thread_inst_per_launch": int(tinst), "lanes_per_inst": round(tinst / max(winst, 1.0), 2),
                    "issue_slot_util": round(winst / (avg_launch_ms * 1e-3) / (sms * 4 * sm_mhz * 1e6), 4) if avg_launch_ms > 0 else None,
                    "counters": "ncu pass over one step of this build, run inside this bench invocation (instruction / DRAM-byte counts only; times are live CUDA events)",
                    "hbm": hbm_part}
    else:
        roofline = {"bound": "issue", "kernel": "trace_closest (%s loop)" % ctx_trace_mode, "achieved": None, "peak": round(issue_peak, 1), "unit": "Gthread-inst/s", "frac": None, "traffic": None,
                    "counters": "unavailable: no ncu pass in this run (N > 1, --no-ncu, or ncu not permitted); see profiles/ for the round's captures", "hbm": hbm_part}
    roofline.update({"avg_launch_ms": round(avg_launch_ms, 4), "launches": tc["launches"],
                "per_ray": {"nodes": round(cw["nodes"] / max(cw["rays"], 1), 2), "tris": round(cw["tris"] / max(cw["rays"], 1), 2),
                            "insts": round(cw["insts"] / max(cw["rays"], 1), 3)},
                "shadow_per_ray": {"nodes": round(aw["nodes"] / max(aw["rays"], 1), 2), "tris": round(aw["tris"] / max(aw["rays"], 1), 2),
                                   "insts": round(aw["insts"] / max(aw["rays"], 1), 3)},
                "kernel_ms_per_step": {k: round(prof[k]["ms"] / K, 4) for k in pt.KERNELS},
                "timed_pass": "a second pass over the same K frames with a CUDA event pair around every launch and the two trace streams serialised, so each kernel is timed alone (the headline region runs them overlapped and without per-launch events)"})

    # ---- e2e: the call a host application makes per frame, HOST buffers on both sides, copies inside the timed region:
    # camera + render settings from host structs (H2D: the kernel parameter block), one frame, tone-mapped RGBA8 frame read
    # back into pinned host memory (what Renderer::Render + UnpackToTexture move per frame in the reference).
    pt.ResetFrameNumber()
    host_rgba = [torch.empty((res[1], res[0]), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32) for _ in range(2)]
    cam = desc["camera"]
    # one untimed warm-up of this exact call sequence: the first read-back after the counted replay above carries a one-time
    # cost of tens of milliseconds (first use of the display kernel / stream-ordered pool growth) that is not part of a step
    scene.SetCamera(cam); scene.SetRenderSettings(desc["settings"])
    for k in range(2):
        pt.Render(scene, frames=1, firstFrame=first); pt.PresentWait(pt.Present(scene, host_rgba[k]))
    pt.ReadRGBA8(scene, out=host_rgba[0]); pt.Stats()
    pt.ResetFrameNumber()

    def e2e_loop(pipelined):
        """K steps of: host camera + settings in -> Render(1 frame) -> tone-mapped RGBA8 frame in pinned host memory, every step.
        pipelined: the reference's display path (Render returns without synchronising, the pixel buffer is consumed a frame
        later): frame i's resolve + copy are queued behind it, the host waits for frame i-1's image while frame i renders, so
        the 33 MB read-back of one frame overlaps the traversal of the next.  Otherwise every step ends with a blocking read."""
        barrier()
        t0 = time.time()
        rays, step_ms, prev = 0, [], None
        for i in range(K):
            ts = time.time()
            scene.SetCamera(cam)
            scene.SetRenderSettings(desc["settings"])
            pt.Render(scene, frames=1, firstFrame=first + i)
            if pipelined:
                ticket = pt.Present(scene, host_rgba[i & 1])
                if prev is not None:
                    s1 = pt.PresentWait(prev); rays += s1["extension_rays"] + s1["shadow_rays"]
                prev = ticket
            else:
                pt.ReadRGBA8(scene, out=host_rgba[0])
                s1 = pt.Stats(); rays += s1["extension_rays"] + s1["shadow_rays"]
            step_ms.append((time.time() - ts) * 1e3)
        if pipelined:
            s1 = pt.PresentWait(prev); rays += s1["extension_rays"] + s1["shadow_rays"]     # the last frame's image, inside the timed region
        barrier()
        return time.time() - t0, rays, step_ms

    e2e_s, e2e_rays, step_ms = e2e_loop(True)
    pt.ResetFrameNumber()
    sync_s, sync_rays, _ = e2e_loop(False)
    e2e_t = torch.tensor([e2e_s, sync_s], device="cuda", dtype=torch.float64)
    e2e_c = torch.tensor([float(e2e_rays), float(sync_rays)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_c, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_c[0]) / float(e2e_t[0]) / 1e6
    # H2D per step: nx_camera (48 B) + nx_render_settings (32 B), marshalled into the kernels' parameter block
    e2e = {"value": round(e2e_value, 1), "unit": "Mrays/s", "h2d_bytes_per_step": 48 + 32, "d2h_bytes_per_step": 4 * res[0] * res[1] + 32,
           "ms_per_step": round(float(e2e_t[0]) * 1e3 / K, 3), "ms_per_step_median": round(float(np.median(step_ms)), 3), "ms_per_step_max": round(float(np.max(step_ms)), 3),
           "what": "every step: SetCamera+SetRenderSettings (host structs) -> Render(1 frame) -> Present (display transform + RGBA8 frame + queue totals copied to pinned host memory on a copy stream); the host waits for frame i-1's image while frame i renders (double-buffered, like the reference's pixel-buffer display path); the last image is awaited inside the timed region",
           "blocking_read": {"value": round(float(e2e_c[1]) / float(e2e_t[1]) / 1e6, 1), "ms_per_step": round(float(e2e_t[1]) * 1e3 / K, 3),
                             "what": "same steps with a blocking ReadRGBA8 + Stats at the end of every step (no overlap)"}}

    # ---- like for like (N = 1 only): the same frames on BLASes / TLAS collapsed by the reference GPU converter's rule, i.e. trees
    # identical to the ones the reference renders with.  The headline above uses the product default, the SAH-optimal collapse of
    # the reference's CPU BVH8Builder run on the GPU (same hits, fewer node visits); this line isolates what the trees contribute.
    two_level = None
    if world == 1 and not args.no_like_for_like:
        # the product's trees WITHOUT instance merging: every mesh its own SAH-optimal BLAS under a TLAS of 1,026 instances (round 1's default)
        ctx.SetInstanceMerging(False)
        scene_tl = scenes.build(ctx, desc, res)
        ctx.SetInstanceMerging(True)
        pt.ResetFrameNumber()
        pt.Render(scene_tl, frames=3, firstFrame=1); ctx.synchronize()
        pt.ResetFrameNumber()
        kk = min(K, 8)
        pt.Render(scene_tl, frames=kk, firstFrame=first); ctx.synchronize()
        s3 = pt.Stats()
        two_level = {"what": "instance merging off: one SAH-optimal BLAS per mesh under a TLAS over all instances (clipped bounds, bounding-sphere cull)",
                     "value": round((s3["extension_rays"] + s3["shadow_rays"]) / s3["device_ms"] / 1e3, 1), "unit": "Mrays/s", "ms_per_step": round(s3["device_ms"] / kk, 4), "steps": kk}
        scene_tl.close()
    like = None
    if world == 1 and not args.no_like_for_like:
        ctx.SetSceneCollapse(nx.COLLAPSE_REFERENCE_GPU, 0)
        scene_ref = scenes.build(ctx, desc, res)
        ctx.SetSceneCollapse(nx.COLLAPSE_SAH_OPTIMAL, 2)
        pt.ResetFrameNumber()
        pt.Render(scene_ref, frames=3, firstFrame=1); ctx.synchronize()
        pt.ResetFrameNumber()
        kk = min(K, 8)
        pt.Render(scene_ref, frames=kk, firstFrame=first); ctx.synchronize()
        s2 = pt.Stats()
        like = {"collapse": "reference_gpu (NexusBVH-identical trees)", "value": round((s2["extension_rays"] + s2["shadow_rays"]) / s2["device_ms"] / 1e3, 1), "unit": "Mrays/s",
                "ms_per_step": round(s2["device_ms"] / kk, 4), "steps": kk}
        scene_ref.close()

    # ---- CPU baseline (rank 0, N = 1 only): the CPU oracle's closest-hit traversal on a bounded sample of this workload's primary rays
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(desc, scene, res)

    config5 = None
    if world > 1 and args.workload == "instanced10m_4k" and not args.no_config5:
        config5 = run_config5(nx, ctx, pt, world, rank, res, stream, barrier, sharded)

    if rank == 0:
        line = {"metric": "Mrays/s", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "description": wl["desc"], "resolution": list(res), "path_length": desc["settings"].pathLength},
                "notes": {"partition": f"sample partition: rank g renders frames [1+g*K, (g+1)*K]; NCCL all-reduce(sum) of float[3*W*H] accumulation ({'%.1f' % (12e-6 * res[0] * res[1])} MB)" if world > 1 else "single GPU",
                          "bvh": "H-PLOC BVH2 + SAH-optimal CWBVH8 collapse (reference CPU BVH8Builder's C(n,i) table on the GPU, <= 2 primitives per leaf); instances whose mesh is used once are transformed to world space and share one BLAS under the TLAS (instance merging); same hits as the reference's trees: two_level = merging off, like_for_like = NexusBVH-identical trees",
                          "l2": "per-step working set (ray/hit/state queues %.0f MB at this resolution + BVH/triangles) exceeds the 126 MB L2; no flush needed" % (212e-6 * res[0] * res[1]),
                          "traversal": ctx_trace_mode},
                "spp_per_s": round(world * K / (ms_total * 1e-3), 2),
                "rays_per_step": int(rays_all / (K * world)), "extension_rays": int(cnt[1]), "shadow_rays": int(cnt[2]),
                "primary_Mrays_per_s": round(world * K * res[0] * res[1] / (ms_total * 1e-3) / 1e6, 1),
                "reduce_ms": round(ms_reduce, 3), "wall_s": round(wall, 3), "scene_setup_s": round(t_scene, 2),
                "blas_builds": f"merging off, per-mesh BLAS builds sharded round-robin over {world} ranks + NCCL all-gather" if sharded else "every rank builds its own trees (merged world-space BLAS + the BLASes of shared / moved meshes)",
                "mean_radiance": round(mean_radiance, 5),
                "gpu_launches": int(st["kernel_launches"]), "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "two_level": two_level, "like_for_like": like,
                "reduce_check": reduce_check, "config5": config5}
        print(json.dumps(line), flush=True)
    pt.close(); scene.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_config5(nx, ctx, pt, world, rank, res, stream, barrier, sharded, total_spp=64):
    """BASELINE.json configs[4] (SURVEY.md 8d "Config 5 inputs") inside an N > 1 run: configs[2]'s geometry under a procedural HDR sky,
    64 samples per pixel in total = 64 / N frames per rank (sample partition), one NCCL all-reduce of the float accumulation sums, the
    reduce timed separately.  In-run correctness: the float64 checksum of the reduced buffer must equal the all-reduced sum of the
    per-rank checksums taken before the reduce."""
    import torch
    import torch.distributed as dist
    from nexus_b200 import scenes
    from nexus_b200.multigpu import accumulation_tensor, build_scene_sharded, frame_block
    desc = make_desc("sky10m_4k")
    t0 = time.time()
    scene = build_scene_sharded(ctx, desc, res) if sharded else scenes.build(ctx, desc, res)
    ctx.synchronize()
    t_scene = time.time() - t0
    per = max(1, total_spp // world)
    acc = accumulation_tensor(pt, torch.device("cuda", ctx.device))
    pt.ResetFrameNumber()
    pt.Render(scene, frames=2, firstFrame=1)
    with torch.cuda.stream(stream):
        dist.all_reduce(acc)
    barrier()
    pt.ResetFrameNumber()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    e0.record(stream)
    pt.Render(scene, frames=per, firstFrame=frame_block(rank, world, per))
    e1.record(stream)
    with torch.cuda.stream(stream):
        local_sum = acc.sum(dtype=torch.float64)                 # 100 MB read, ~20 us: inside the timed region
        dist.all_reduce(acc)
    e2.record(stream)
    barrier()
    with torch.cuda.stream(stream):
        reduced_sum = acc.sum(dtype=torch.float64)
        dist.all_reduce(local_sum)
    barrier()
    pt.SetAccumulatedFrames(per * world)
    st = pt.Stats()
    t = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2)], device="cuda", dtype=torch.float64)
    c = torch.tensor([float(st["extension_rays"] + st["shadow_rays"])], device="cuda", dtype=torch.float64)
    t_min = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c)
    dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
    want, got = float(local_sum), float(reduced_sum)
    rel = abs(got - want) / max(abs(want), 1e-30)
    assert rel < 1e-5, f"reduced accumulation checksum {got} != sum of the per-rank checksums {want} (relative {rel:.2e})"
    # the collective alone, ranks aligned by a barrier first (the in-region figure above also contains the skew between ranks)
    scratch = torch.empty_like(acc)
    alone = []
    for _ in range(5):
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a0.record(stream); dist.all_reduce(scratch); a1.record(stream)
        barrier()
        alone.append(a0.elapsed_time(a1))
    alone_t = torch.tensor([float(np.median(alone))], device="cuda", dtype=torch.float64)
    dist.all_reduce(alone_t, op=dist.ReduceOp.MAX)
    mean = float(pt.ReadAccumulation().mean()) if rank == 0 else 0.0
    scene.close()
    ms_total, ms_reduce = float(t[0]), float(t[1])
    return {"workload": "sky10m_4k", "what": "BASELINE configs[4]: 8xB200 sample-partitioned 4K render, procedural HDR environment map, 64 spp total, NCCL all-reduce of the accumulation buffers",
            "spp_total": per * world, "frames_per_rank": per, "value": round(float(c[0]) / (ms_total * 1e-3) / 1e6, 1), "unit": "Mrays/s",
            "spp_per_s": round(per * world / (ms_total * 1e-3), 2), "ms_total": round(ms_total, 3), "ms_per_frame": round((ms_total - ms_reduce) / per, 4),
            "reduce_ms": round(ms_reduce, 3), "reduce_ms_last_arrival": round(float(t_min[1]), 3),
            "reduce_what": "reduce_ms: checksum + all-reduce as the rank that finished rendering FIRST sees it (max over ranks: contains its wait for the slowest rank); reduce_ms_last_arrival: the same on the rank that arrived last (min over ranks: the collective itself); reduce_alone_ms: the all-reduce with the ranks aligned by a barrier",
            "reduce_alone_ms": round(float(alone_t[0]), 3), "reduce_bytes": int(acc.numel() * 4),
            "reduce_busbw_GBs": round(2.0 * (world - 1) / world * acc.numel() * 4 / (float(alone_t[0]) * 1e-3) / 1e9, 1),
            "reduce_check": {"sum_of_rank_checksums": want, "reduced_checksum": got, "rel_err": rel}, "mean_radiance": round(mean, 5), "scene_setup_s": round(t_scene, 2)}


def cpu_baseline(desc, scene, res, budget_s=15.0):
    """CPU port (oracle/oracle_trace.cpp) of the closest-hit traversal, all host cores, on a bounded sample of the workload's
    primary rays.  The oracle is used here as the thing being timed on the CPU, never on the product path."""
    import oracle_lib as O
    from nexus_b200 import scenes
    import nexus_b200 as nx
    cores = os.cpu_count() or 1
    ora = O.oracle_scene_from_product(desc, scene)
    n = 20000
    o, d = scenes.camera_rays(desc["camera"], res, n=n, seed=1)
    rays = nx.make_rays(o, d)
    t0 = time.time(); ora.trace_closest(rays, threads=cores); dt = time.time() - t0
    n2 = int(min(4_000_000, max(n, n * budget_s / max(dt, 1e-4))))
    o, d = scenes.camera_rays(desc["camera"], res, n=n2, seed=2)
    rays = nx.make_rays(o, d)
    t0 = time.time(); ora.trace_closest(rays, threads=cores); dt = time.time() - t0
    return {"value": round(len(rays) / dt / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{len(rays)} primary (camera) rays of this workload, closest-hit traversal only, {dt:.1f} s"}



# ------------------------------------------------------------------------------------- builder workloads ----
# algorithmic bytes per primitive with 32-bit keys, each BVH2 node counted once per phase.
BUILD_BYTES_FIXED = 340           # + 80 * N8 / n for the CWBVH8 node writes
HPLOC_BYTES_PER_PRIM = 32 + 32 + 16 + 8 + 8   # leaf bounds read, inner node write, clusterIdx load/store, parentIdx, two Morton neighbours


def build_mesh(wl):
    from nexus_b200 import scenes
    if wl.get("mesh") == "uv_sphere":
        return np.ascontiguousarray(scenes.uv_sphere(224, 224).reshape(-1, 9), np.float32)
    return scenes.test_triangles(wl["n"])


def cpu_sah_baseline(tris):
    """BASELINE.json configs[0]'s CPU path on this box's host cores: binned-SAH BVH2 (all cores) + SAH-optimal collapse (sequential,
    as the reference's recursive BVH8Builder is), best of 3."""
    import oracle_lib as O
    cores = os.cpu_count() or 1
    pb, sb = O.prim_bounds(tris, 1)
    n = len(pb)
    best = None
    for _ in range(3):
        t0 = time.time(); n2 = O.sah_build_bvh2(pb, threads=cores); t1 = time.time(); n8, pidx, root_cost = O.sah_collapse(n2, n); t2 = time.time()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    out = {"value": round(n / best[0] / 1e6, 4), "unit": "Mprims/s", "cores": cores, "kind": "port",
           "sample": f"all {n} triangles: binned-SAH BVH2 build {best[1] * 1e3:.1f} ms on {cores} threads + SAH-optimal BVH8 collapse {best[2] * 1e3:.1f} ms on 1 thread "
                     "(oracle/oracle_sah.cpp; the reference snapshot contains no CPU BVH2 builder and its CPU collapse is dead code, SURVEY.md header note 1)",
           "bvh8_nodes": int(len(n8)), "bvh2_sah_leaf_cost": round(O.sah_bvh2_cost(n2, n), 3), "collapse_root_cost": round(float(root_cost), 4),
           "bvh8_sah": round(O.bvh8_cost(n8, sb), 4)}
    t0 = time.time(); O.sah_build_bvh2(pb, threads=1); t1 = time.time() - t0       # single-thread figure beside the all-core one
    out["single_thread"] = {"bvh2_build_ms": round(t1 * 1e3, 1), "Mprims_per_s": round(n / (t1 + best[2]) / 1e6, 4)}
    if O.have_refcpu():
        # the collapse half also through the reference's own code: Nexus/src/Geometry/BVH/BVH8Builder.cpp compiled unmodified (oracle/_ref)
        tr = []
        for _ in range(3):
            t0 = time.time(); rn, rp, rc = O.ref_cpu_collapse(n2, n); tr.append(time.time() - t0)
        out["reference_collapse"] = {"ms": round(min(tr) * 1e3, 1), "what": "the same BVH2 through the unmodified reference BVH8Builder (g++ -O2, 1 thread)",
                                     "identical_to_port": bool((rn == n8).all() and (rp == pidx).all() and rc == root_cost),
                                     "Mprims_per_s_with_it": round(n / (best[1] + min(tr)) / 1e6, 4)}
    return out


def run_build_ours(args):
    import torch
    import torch.distributed as dist
    import nexus_b200 as nx
    from nexus_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, args.warmup
    wl = WORKLOADS[args.workload]
    n = wl["n"]
    speed = True
    ctx = nx.Context(local)
    stream = torch.cuda.ExternalStream(int(nx.lib().nx_ctx_stream(ctx._h)), device=torch.device("cuda", local))
    t_gen = time.time()
    host = torch.empty((n, 9), dtype=torch.float32, pin_memory=True)
    host.numpy()[:] = build_mesh(wl)
    t_gen = time.time() - t_gen
    dev_t = torch.empty((n, 9), dtype=torch.float32, device="cuda")
    dev_t.copy_(host)
    dev = dev_t.data_ptr()

    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(W):
        nx.BuildBVH8Device(ctx, dev, n, 1, speed).Free()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    nodes = 0
    for _ in range(K):                       # replicas only: every rank builds the whole mesh (one hierarchy does not shard)
        b = nx.BuildBVH8Device(ctx, dev, n, 1, speed)
        nodes = b.nodeCount
        b.Free()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    value = world * K * n / (ms_total * 1e-3) / 1e6

    # per-stage device times (the builder's own CUDA events, BVHBuildMetrics layout) over K more builds, and the SAH costs
    m = nx.BenchmarkBuild(ctx, dev, n, 1, speed, 1, K)
    mo = nx.BenchmarkBuild(ctx, dev, n, 1, speed, 1, K, collapse=nx.COLLAPSE_SAH_OPTIMAL, maxLeafPrims=2)
    m64 = nx.BenchmarkBuild(ctx, dev, n, 1, False, 1, min(K, 4))       # prioritizeSpeed = false: 64-bit Morton keys (SURVEY.md 8d, config 4 reports both)
    hbm, hbm_src = peaks()
    hploc_ms = m["bvh2_ms"]
    achieved = HPLOC_BYTES_PER_PRIM * n / (hploc_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload, {}).get("hploc_dram_bytes_per_launch")
    whole_bytes = (BUILD_BYTES_FIXED + 80.0 * nodes / n) * n
    roofline = {"bound": "hbm", "kernel": "hploc_kernel", "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                "traffic": traffic, "peak_source": hbm_src, "avg_launch_ms": round(hploc_ms, 4), "launches": K,
                "algorithmic_bytes_per_launch": int(HPLOC_BYTES_PER_PRIM * n),
                "whole_build": {"bytes_per_prim": round(whole_bytes / n, 1), "achieved_gbs": round(whole_bytes / (ms_total / K * 1e-3) / 1e9, 1),
                                "frac": round(whole_bytes / (ms_total / K * 1e-3) / 1e9 / hbm, 4)},
                "stage_ms": {k: round(v, 4) for k, v in m.items() if k.endswith("_ms")}}

    # e2e: what Mesh::Mesh does per mesh (N/Assets/Mesh.h:29-40): host triangles -> device, BuildBVH8, handle (bounds, counts) back.
    # Pipelined like an importer that loads a scene of many meshes: two device input buffers, the H2D copy of mesh i+1 runs on a copy
    # stream while mesh i is being built; every copy and every build is inside the timed region (the first copy is fully exposed).
    # The copies run on a torch stream (torch's allocator must never see the context's stream, which dies with the context); the
    # builder's stream waits for each through an event.
    bufs = [dev_t, torch.empty_like(dev_t)]
    copy_stream = torch.cuda.Stream()
    copied = [None, None]

    def start_copy(k):
        with torch.cuda.stream(copy_stream):
            bufs[k].copy_(host, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream); copied[k] = ev

    barrier()
    t0 = time.time()
    start_copy(0)
    for i in range(K):
        k = i & 1
        if i + 1 < K:
            start_copy(k ^ 1)                     # the build that read this buffer finished (synchronised) two lines below, one step ago
        stream.wait_event(copied[k])
        b = nx.BuildBVH8Device(ctx, bufs[k].data_ptr(), n, 1, speed)
        ctx.synchronize()
        _ = (b.nodeCount, b.bounds)
        b.Free()
    barrier()
    e2e_s = torch.tensor([time.time() - t0], device="cuda", dtype=torch.float64)
    # and unpipelined: copy, then build, one after the other (what a single Mesh::Mesh call costs)
    barrier()
    t0 = time.time()
    for _ in range(min(K, 4)):
        dev_t.copy_(host, non_blocking=True)
        ev = torch.cuda.Event(); ev.record(); stream.wait_event(ev)
        b = nx.BuildBVH8Device(ctx, dev, n, 1, speed)
        ctx.synchronize()
        b.Free()
    barrier()
    serial_s = (time.time() - t0) / min(K, 4)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = {"value": round(world * K * n / float(e2e_s[0]) / 1e6, 1), "unit": "Mprims/s", "h2d_bytes_per_step": 36 * n, "d2h_bytes_per_step": 48,
           "ms_per_step": round(float(e2e_s[0]) * 1e3 / K, 3),
           "what": "pinned host triangles -> device copy -> BuildBVH8 -> handle (bounds, node count) on the host, every step; double-buffered input, the copy of step i+1 overlaps the build of step i",
           "unpipelined": {"value": round(n / serial_s / 1e6, 1), "ms_per_step": round(serial_s * 1e3, 3), "what": "copy, then build, strictly one after the other"}}
    del bufs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and wl.get("mesh") == "uv_sphere":
        cpu = cpu_sah_baseline(host.numpy())
    elif rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib as O
        ns = 1_000_000
        sample = host.numpy()[:ns]
        t0 = time.time(); n8, _, _ = O.cpu_build_bvh8(sample, 1, 0); dt = time.time() - t0
        cpu = {"value": round(ns / dt / 1e6, 4), "unit": "Mprims/s", "cores": 1, "kind": "port",
               "sample": f"first {ns} triangles of this mesh through the CPU restatement of the whole pipeline (bounds, Morton, H-PLOC, collapse), single thread, {dt:.1f} s"}
    if rank == 0:
        line = {"metric": "Mprims/s", "value": round(value, 1), "unit": "Mprims/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
                "config": {"workload": args.workload, "description": wl["desc"], "triangles": n, "prioritize_speed": speed},
                "notes": {"partition": "replicas only: one global sort + one hierarchy does not shard; each rank builds the whole mesh" if world > 1 else "single GPU",
                          "l2": "input (%.1f GB) and every intermediate array exceed the 126 MB L2; no flush needed" % (36e-9 * n)},
                # the reference's own metric (BVHBuildMetrics::totalTime = the sum of the per-stage CUDA-event times, which is what
                # the reference arm reports as its value): the same definition for this arm, beside the stricter whole-step `value`
                "Mprims_per_s_stage_sum": round(n / m["total_ms"] / 1e3, 1),
                "bvh8_nodes": int(nodes), "bvh2_sah": round(m["bvh2_cost"], 4), "bvh8_sah": round(m["bvh8_cost"], 4), "avg_children_per_node": round(m["avg_children_per_node"], 3),
                "sah_optimal_collapse": {"what": "same build with nx_build_config.collapse = NX_COLLAPSE_SAH_OPTIMAL, max_leaf_prims 2 (what the renderer builds its BLASes with)",
                                         "total_ms": round(mo["total_ms"], 4), "bvh8_ms": round(mo["bvh8_ms"], 4), "bvh8_nodes": int(mo["node_count"]), "bvh8_sah": round(mo["bvh8_cost"], 4),
                                         "Mprims_per_s": round(n / mo["total_ms"] / 1e3, 1)},
                "morton64": {"what": "same build with prioritizeSpeed = false (64-bit Morton keys, sort over bits [1, 64))", "total_ms": round(m64["total_ms"], 4),
                             "Mprims_per_s": round(n / m64["total_ms"] / 1e3, 1), "stage_ms": {k: round(v, 4) for k, v in m64.items() if k.endswith("_ms")},
                             "bvh8_nodes": int(m64["node_count"]), "bvh2_sah": round(m64["bvh2_cost"], 4), "bvh8_sah": round(m64["bvh8_cost"], 4)},
                "host_generate_s": round(t_gen, 1), "gpu_launches": int(K * 9), "library_launches": 0,   # per build: leaf bounds, Morton keys, sort prefix, 4 onesweep passes, H-PLOC, collapse (profiles/r02_ncu_launches_build10m.csv); no library kernel
                 "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    del dev_t, host
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_build_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import ctypes as C
    import oracle_lib as O
    from nexus_b200 import scenes      # scene/mesh generators only (numpy); the product library is not loaded on this arm
    K, W = args.steps, args.warmup
    wl = WORKLOADS[args.workload]
    n = wl["n"]
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libnexus_ref.so was not built (needs /root/reference at build time)"}))
        return
    tris = build_mesh(wl)
    mm = np.zeros(9, np.float32); cnt = C.c_uint32(0); ms2 = np.zeros(2, np.float32)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start(); time.sleep(0.3)
    rc = O.ref().nxref_benchmark_bvh8(tris.ctypes.data_as(C.c_void_p), C.c_uint32(n), 1, 1, W, K, mm.ctypes.data_as(C.c_void_p), C.byref(cnt), ms2.ctypes.data_as(C.c_void_p))
    clocks = sampler.stop()
    assert rc == 0
    mm64 = np.zeros(9, np.float32); cnt64 = C.c_uint32(0); ms64 = np.zeros(2, np.float32)      # prioritizeSpeed = false (64-bit Morton keys)
    rc = O.ref().nxref_benchmark_bvh8(tris.ctypes.data_as(C.c_void_p), C.c_uint32(n), 1, 0, 1, min(K, 4), mm64.ctypes.data_as(C.c_void_p), C.byref(cnt64), ms64.ctypes.data_as(C.c_void_p))
    assert rc == 0
    # The reference's own metric (BVHBuildMetrics::totalTime: the sum of its per-stage CUDA-event times, what NexusBVH's README
    # quotes) is the value; what a caller of BuildBVH8 actually waits for (cudaMallocAsync + cudaFree of every array inside each
    # build) is reported beside it as caller_ms_per_step and is several times longer.
    ms = float(mm[5]) * K
    value = K * n / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": "Mprims/s", "value": round(value, 1), "unit": "Mprims/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 4), "caller_ms_per_step": round(float(ms2[0]) / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "triangles": n, "prioritize_speed": True},
            "bvh8_nodes": int(cnt.value), "bvh2_sah": round(float(mm[6]), 4), "bvh8_sah": round(float(mm[7]), 4), "avg_children_per_node": round(float(mm[8]), 3),
            "stage_ms": {"computeSceneBoundsTime": round(float(mm[0]), 4), "computeMortonCodesTime": round(float(mm[1]), 4), "radixSortTime": round(float(mm[2]), 4),
                         "bvhBuildTime": round(float(mm[3]), 4), "bvh8ConversionTime": round(float(mm[4]), 4), "totalTime": round(float(mm[5]), 4)},
            "morton64": {"what": "same build with prioritizeSpeed = false (64-bit Morton keys)", "total_ms": round(float(mm64[5]), 4),
                         "Mprims_per_s": round(n / float(mm64[5]) / 1e3, 1), "bvh8_nodes": int(cnt64.value), "bvh2_sah": round(float(mm64[6]), 4), "bvh8_sah": round(float(mm64[7]), 4)},
            "clocks": clocks,
            "cpu_baseline": {"value": round(value, 1), "unit": "Mprims/s", "cores": 0, "kind": "reference",
                             "sample": f"{K} builds of this mesh through the unmodified NexusBVH (compiled -arch=sm_100a --use_fast_math from /root/reference) on the same B200; "
                                       "the reference has no CPU builder in this snapshot"},
            "e2e": {"value": round(value, 1), "unit": "Mprims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)

# ----------------------------------------------------------------------------------------- reference arm ----
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    K, W = args.steps, args.warmup
    wl = WORKLOADS[args.workload]
    res = wl["res"]
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libnexus_ref.so was not built (needs /root/reference at build time)"}))
        return
    desc = make_desc(args.workload)
    O.ref_load_scene_standalone(desc, res)
    O.ref_render(1, max(W, 1))
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start(); time.sleep(0.3)
    # the reference's running mean restarts at frame 1; its RNG is keyed on the frame number, so use the same indices as our arm
    ms, ext, sh = O.ref_render(1, K)
    clocks = sampler.stop()
    value = (ext + sh) / (ms * 1e-3) / 1e6
    img = O.ref_read_accum(res)
    line = {"impl": "reference", "metric": "Mrays/s", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "resolution": list(res), "path_length": desc["settings"].pathLength},
            "spp_per_s": round(K / (ms * 1e-3), 2), "rays_per_step": int((ext + sh) / K), "extension_rays": ext, "shadow_rays": sh,
            "primary_Mrays_per_s": round(K * res[0] * res[1] / (ms * 1e-3) / 1e6, 1), "mean_radiance": round(float(img.mean()), 5),
            "clocks": clocks,
            "cpu_baseline": {"value": round(value, 1), "unit": "Mrays/s", "cores": 0, "kind": "reference",
                             "sample": f"{K} full frames of this workload through the unmodified reference CUDA kernels (compiled -arch=sm_100a --use_fast_math "
                                       "from /root/reference) on the same B200; the reference has no CPU implementation of this path"},
            "e2e": {"value": round(value, 1), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="instanced10m_4k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="skip the ncu counter pass behind roofline.frac / roofline.traffic")
    ap.add_argument("--count-pass", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-like-for-like", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the BASELINE configs[4] block (HDR sky, 64 spp total)")
    ap.add_argument("--sharded-build", action="store_true", help="N > 1: instance merging off, per-mesh BLAS builds sharded over the ranks + NCCL all-gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    build = args.workload.startswith("build")
    if args.count_pass:
        return run_count_pass(args)
    if args.impl == "reference":
        (run_build_reference if build else run_reference)(args)
    else:
        (run_build_ours if build else run_ours)(args)


if __name__ == "__main__":
    main()

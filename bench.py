#!/usr/bin/env python
"""Benchmark of the rasterize -> sample -> approximate-gradient hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

One "step" = one forward + backward of ``rasterize_rgba`` over one batch of synthetic views
(BASELINE.json config 2 at N = 1: teapot, 64 views per GPU, 512x512, no anti-aliasing, RGBA with a
texture_size-4 atlas; upstream gradient G ~ N(0,1) fixed).  Metric:
megapixel-views per second = views * S^2 / 1e6 / time.  For N > 1 the driver launches this file under
torchrun; every rank renders its own 64 views (no data-path collective: views are independent), the
step time is the max over ranks, `value` the aggregate.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: `roofline`, `cpu_baseline`,
`e2e`, `gpu_launches`, `clocks`, `kernels_ms`.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if "reference" in sys.argv[1:]:
    # the CPU arm uses every host core; torchrun exports OMP_NUM_THREADS=1, and the OpenMP runtime reads it
    # once, when torch loads it
    try:
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    except AttributeError:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (views per GPU, S, anti_aliasing, mode, texture_size)
    "cfg2": dict(views=64, S=512, aa=False, mode="rgba", ts=4, mesh="teapot"),
    "cfg2s": dict(views=64, S=512, aa=False, mode="silhouettes", ts=0, mesh="teapot"),
    "cfg1": dict(views=1, S=256, aa=True, mode="rgba", ts=16, mesh="teapot"),
    "cfg5": dict(views=32, S=512, aa=True, mode="rgb", ts=4, mesh="teapot"),
    # config 2 through the Renderer facade: world-space vertices -> camera transform -> rasterize
    "cfg2r": dict(views=64, S=512, aa=False, mode="rgba", ts=4, mesh="teapot", renderer=True),
    # multi-view optimisation of ONE shared 100k-face mesh: gradient all-reduce across ranks
    "cfg3": dict(views=8, S=512, aa=False, mode="silhouettes", ts=0, mesh="sphere", shared=True),
    # 1M independent ~5 px triangles: stresses binning, list sorting and z-test contention
    "cfg4": dict(views=16, S=1024, aa=False, mode="silhouettes", ts=0, mesh="random1m"),
}
METRIC = "megapixel-views/sec fwd+bwd"
UNIT = "Mpix-views/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(nv, nf, T, P, C, S, depth=False):
    """per view, forward and backward shares."""
    fwd = 12 * nv + 12 * nf + 12 * T + 4 * P + 12 * P + (4 * P if depth else 0) + 4 * C * S * S
    bwd = 4 * C * S * S + 16 * P + (4 * P if depth else 0) + 12 * nv + 12 * T + 12 * T
    return fwd, bwd


def sphere_mesh(n=225, perturb=0.0, seed=0):
    """UV sphere, n x n vertices (n=225: 50 625 vertices, 100 352 faces); radius 0.8 times a seeded
    low-frequency perturbation."""
    i = torch.arange(n, dtype=torch.float32)
    theta = (i / (n - 1) * 3.14159265)[:, None].expand(n, n)
    phi = (i / (n - 1) * 2 * 3.14159265)[None, :].expand(n, n)
    r = torch.full((n, n), 0.8)
    if perturb:
        g = torch.Generator().manual_seed(seed)
        for _ in range(6):
            a, kt, kp, ph = (torch.rand(1, generator=g).item() for _ in range(4))
            r = r * (1 + perturb * (a - 0.5) * torch.sin((1 + int(kt * 4)) * theta + 6.28 * ph) *
                     torch.cos((1 + int(kp * 4)) * phi))
    v = torch.stack((r * torch.sin(theta) * torch.cos(phi), r * torch.cos(theta), r * torch.sin(theta) * torch.sin(phi)), -1)
    idx = (torch.arange(n - 1)[:, None] * n + torch.arange(n - 1)[None, :]).reshape(-1)
    f = torch.cat((torch.stack((idx, idx + 1, idx + n), 1), torch.stack((idx + 1, idx + n + 1, idx + n), 1)), 0)
    return v.reshape(-1, 3).contiguous(), f.to(torch.int32).contiguous()


def random_triangle_mesh(nf=1000000, seed=0):
    """nf independent triangles: centres U(-1,1)^3, corner offsets N(0, 0.01^2); faces = arange."""
    g = torch.Generator().manual_seed(seed)
    c = torch.rand((nf, 1, 3), generator=g) * 2 - 1
    v = (c + torch.randn((nf, 3, 3), generator=g) * 0.01).reshape(-1, 3)
    v = v - v.min(0).values[None]                       # load_obj-style normalisation (load_obj.py:157-161)
    v = v / v.abs().max() * 2
    v = v - v.max(0).values[None] / 2
    return v.contiguous(), torch.arange(nf * 3, dtype=torch.int32).reshape(nf, 3)


def make_inputs(w, seed, device, nr):
    B, S = w["views"], w["S"]
    g = torch.Generator().manual_seed(seed)
    elev = torch.rand(B, generator=g) * 80. - 20.
    azim = torch.rand(B, generator=g) * 360.
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), elev, azim)
    mesh = w.get("mesh", "teapot")
    if mesh == "teapot":
        d = np.load(os.path.join(ROOT, "tests", "golden", "teapot.npz"))
        v_world, faces = torch.from_numpy(d["vertices"]), torch.from_numpy(d["faces"])
    elif mesh == "sphere":
        v_world, faces = sphere_mesh(225, perturb=0.3, seed=0)
    else:
        v_world, faces = random_triangle_mesh(1000000, seed=0)
    tdev = device if (mesh == "random1m" and str(device) != "cpu") else "cpu"   # 3M vertices x 16 views: transform on the GPU
    vs = nr.perspective(nr.look_at(v_world.to(tdev)[None].expand(B, -1, -1), eye.to(tdev))).contiguous()
    out = dict(vertices=vs, faces=faces, nv=vs.shape[1], nf=faces.shape[0], T=0, eye=eye, v_world=v_world)
    if w["mode"] in ("rgb", "rgba"):
        vt_np, ft_np, tex_np = nr.create_textures(faces.shape[0], w["ts"])
        gt = torch.Generator().manual_seed(0)
        out["textures"] = torch.rand((B,) + tex_np.shape, generator=gt)
        out["vt"] = torch.from_numpy(vt_np)[None].repeat(B, 1, 1).contiguous()
        out["ft"] = torch.from_numpy(ft_np)
        out["T"] = tex_np.shape[1] * tex_np.shape[2]
    C = {"rgba": 4, "rgb": 3, "silhouettes": 1, "depth": 1}[w["mode"]]
    gg = torch.Generator().manual_seed(1)
    shape = (B, C, S, S) if C > 1 else (B, S, S)
    out["G"] = torch.randn(shape, generator=gg)
    out["C"] = C
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            self._stop_evt.wait(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


# ------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch.distributed as dist
    import neural_renderer_v2_pytorch_b200 as nr
    from neural_renderer_v2_pytorch_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its "NCCL version ..." banner (NCCL_DEBUG=VERSION / WARN) on stdout, next to the JSON
        # line; send its log to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    B, S = w["views"], w["S"]
    R = S * 2 if w["aa"] else S
    inp = make_inputs(w, seed=1000 + rank, device=dev, nr=nr)
    C = inp["C"]
    L = _lib.lib()

    faces = inp["faces"].to(dev)
    G = inp["G"].to(dev)
    rgb = w["mode"] in ("rgb", "rgba")
    ft = inp["ft"].to(dev) if rgb else None
    vt = inp["vt"].to(dev) if rgb else None
    fn = {"rgba": nr.rasterize_rgba, "rgb": nr.rasterize_rgb, "silhouettes": nr.rasterize_silhouettes,
          "depth": nr.rasterize_depth}[w["mode"]]

    def step(v, tex):
        hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"])
        p = nr.RasterizeParam(vertices_textures=vt, faces_textures=ft, textures=tex) if rgb else nr.RasterizeParam()
        images = fn(v, faces, p, hp)
        images.backward(G)           # == ((images * G).sum()).backward(), SURVEY.md 8(d)
        return images

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm: `value`
    shared = bool(w.get("shared"))
    if shared:
        # ONE mesh parameter for all views of all ranks: world -> screen transform inside the step,
        # squared-error loss against the unperturbed sphere, gradient summed over views and ranks
        param = inp["v_world"][None].to(dev).requires_grad_(True)
        eye_d = inp["eye"].to(dev)
        tv, _ = sphere_mesh(225, 0.0)
        with torch.no_grad():
            target = fn(nr.perspective(nr.look_at(tv.to(dev)[None].expand(B, -1, -1), eye_d)), faces,
                        nr.RasterizeParam(), nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"]))
        params = [param]
        rend = nr.Renderer()
        rend.image_size, rend.anti_aliasing, rend.viewpoints = S, w["aa"], eye_d

        def step_fn():
            # Renderer.render_silhouettes: fused camera transform + rasterizer; the backward sums the
            # gradient over the local views and all-reduces it over the ranks (parallel.py)
            images = rend.render_silhouettes(nr.parallel.share_across_views(param, B), faces)
            ((images - target) ** 2).sum().backward()
            return images
    elif w.get("renderer"):
        # the call a user of the reference makes: Renderer.render(world vertices, ...)
        v_dev = inp["v_world"][None].repeat(B, 1, 1).to(dev).requires_grad_(True)
        tex_dev = inp["textures"].to(dev).requires_grad_(True)
        params = [v_dev, tex_dev]
        rend = nr.Renderer()
        rend.image_size, rend.anti_aliasing, rend.viewpoints = S, w["aa"], inp["eye"].to(dev)
        rend.fused_camera = not args.unfused_camera

        def step_fn():
            images = rend.render(v_dev, faces, vt, ft, tex_dev)
            images.backward(G)
            return images
    else:
        v_dev = inp["vertices"].to(dev).requires_grad_(True)
        tex_dev = inp["textures"].to(dev).requires_grad_(True) if rgb else None
        params = [v_dev] + ([tex_dev] if rgb else [])

        def step_fn():
            return step(v_dev, tex_dev)

    def eager_step():
        for p_ in params:
            p_.grad = None
        return step_fn()

    for _ in range(max(args.warmup, 3)):
        eager_step()
    barrier()
    # the whole step (forward + backward, a handful of launches, and for a shared mesh the NCCL
    # all-reduce of its gradient) is captured once in a CUDA graph and replayed: every replay does the
    # full work on the device; --eager launches from Python instead
    use_graph = not args.eager
    run_step = eager_step
    if use_graph:
        try:
            run_step = nr.capture_step(step_fn, params=params, warmup=2)
        except Exception as e:      # e.g. a collective that cannot be captured on this NCCL build
            if not (shared and world > 1):
                raise
            sys.stderr.write("bench.py: graph capture with the all-reduce failed (%s); launching eagerly\n" % e)
            use_graph = False
            torch.cuda.synchronize(dev)
    for _ in range(max(args.warmup, 3)):      # warm replays: clocks ramp up, lazy kernel loading is over
        run_step()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        run_step()
        if sampler and i == args.steps // 2:
            sampler.sample()
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * S * S / 1e6 / (ms_step / 1e3)

    # ---- per-kernel pass (same inputs, right after the timed region): roofline numbers
    L.nr_profile_enable(1)
    for _ in range(args.steps):
        eager_step()
    torch.cuda.synchronize(dev)
    L.nr_profile_enable(0)
    import ctypes
    ms = (ctypes.c_float * _lib.NR_PROF_SLOTS)()
    cnt = (ctypes.c_int32 * _lib.NR_PROF_SLOTS)()
    L.nr_profile_collect(ms, cnt)
    kern = {n: (ms[i] / cnt[i]) for i, n in enumerate(_lib.PROF_SLOT_NAMES) if cnt[i]}
    if "setup_count" in kern and "scan_tiles" not in kern:
        kern["bin_view"] = kern.pop("setup_count")      # small meshes: the one-kernel cluster binning uses this slot
    # kernels of this library per step (memset nodes not counted)
    launches_per_step = sum(cnt[i] for i, n in enumerate(_lib.PROF_SLOT_NAMES) if n != "memset") / args.steps

    # ---- end-to-end arm: host (pinned) inputs -> H2D -> fwd + bwd -> D2H of the results
    hosts = [p_.detach().cpu().pin_memory() for p_ in params]
    gv_host = torch.empty_like(hosts[0]).pin_memory()
    chk_host = torch.empty(1).pin_memory()

    # Double-buffered pipeline, as a data loader would drive it: while step i runs on the compute
    # stream, the inputs of step i+1 are uploaded on a copy stream into a staging set; a device-to-
    # device copy moves them into the (static) inputs of the captured step.  Every step uploads its
    # own inputs and reads its results back; all of it is inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    compute_stream = torch.cuda.current_stream(dev)
    stages = [[torch.empty_like(p_.data) for p_ in params] for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]       # staging set filled
    ev_free = [torch.cuda.Event() for _ in range(2)]     # staging set consumed

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[slot])
            for s_, h_ in zip(stages[slot], hosts):
                s_.copy_(h_, non_blocking=True)
            ev_up[slot].record(copy_stream)

    def e2e_run(n):
        for ev in ev_free:
            ev.record(compute_stream)
        upload(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)
            compute_stream.wait_event(ev_up[slot])
            for p_, s_ in zip(params, stages[slot]):
                p_.data.copy_(s_, non_blocking=True)
            ev_free[slot].record(compute_stream)
            images = run_step()
            gv_host.copy_(params[0].grad, non_blocking=True)
            chk_host.copy_(images.sum().reshape(1), non_blocking=True)

    e2e_run(max(args.warmup, 4))
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * B * S * S / 1e6 / (ms_e2e / args.steps / 1e3)
    h2d = sum(h_.numel() * 4 for h_ in hosts)
    d2h = gv_host.numel() * 4 + 4

    def finish():
        """Leave without tearing NCCL down: destroy_process_group() can block for minutes on a communicator
        whose all-reduce lives in a captured CUDA graph, and nothing is left to clean up anyway."""
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    # ---- roofline of the dominant kernel
    peak, peak_src = peaks()
    fwd_b, bwd_b = algorithmic_bytes(inp["nv"], inp["nf"], inp["T"], R * R, C, S)
    shares = {"raster": fwd_b * B, "backward": bwd_b * B}
    dom = max((k for k in kern if k in shares), key=lambda k: kern[k])
    achieved = shares[dom] / (kern[dom] / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    step_bytes = (fwd_b + bwd_b) * B
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": shares[dom], "kernel_ms": round(kern[dom], 4),
                "peak_source": peak_src,
                "step_frac": round(step_bytes / (ms_step / 1e3) / 1e9 / peak, 4),
                "how": "CUDA events around every launch on the launch stream, %d profiled steps right after "
                       "the timed region" % args.steps}

    cpu = cpu_baseline_sample(args.workload, budget_s=12.0) if world == 1 and not args.no_cpu_baseline else None

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s (%d v / %d f), %d views per GPU, %dx%d, anti_aliasing=%s, %s%s"
                               % (args.workload, w.get("mesh", "teapot"), inp["nv"], inp["nf"], B, S, S, w["aa"], w["mode"],
                                  (", texture_size %d (T=%d texels/view)" % (w["ts"], inp["T"])) if rgb else ""),
                   "views_per_gpu": B, "global_views": B * world, "image_size": S, "parallelism": "dp%d" % world,
                   "l2": "per-step working set %.2f GB > 126 MB L2, no explicit flush" % (step_bytes / 1e9)},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 4),
                "pipeline": "inputs of step i+1 uploaded from pinned memory on a copy stream while step i runs; "
                            "per step: H2D of all differentiable inputs, fwd+bwd, D2H of vertex gradients + checksum"},
        "gpu_launches": int(round(launches_per_step * args.steps)),
        "launch_mode": "cuda graph replay of the whole step" if use_graph else "eager (python)",
        "host_ms_per_step": round(host_ms, 4),
        "clocks": clocks,
        "roofline": roofline,
        "kernels_ms": {k: round(v, 4) for k, v in kern.items()},
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    finish()


# ------------------------------------------------------------------------------------- CPU side
def oracle_step(inp, w, views):
    """One fwd + bwd of the CPU oracle (C z-buffer / weight map with OpenMP + torch-CPU stages)."""
    from oracle import pipeline as ref
    v = inp["vertices"][:views].clone().requires_grad_(True)
    kw = {}
    rgb = w["mode"] in ("rgb", "rgba")
    if rgb:
        tex = inp["textures"][:views].clone().requires_grad_(True)
        kw = dict(vertices_textures=inp["vt"][:views], faces_textures=inp["ft"].numpy(), textures=tex)
    img = ref.rasterize(v, inp["faces"], w["S"], w["aa"], draw_rgb=rgb,
                        draw_silhouettes=w["mode"] in ("silhouettes", "rgba"), draw_depth=w["mode"] == "depth", **kw)
    G = inp["G"][:views]
    img.backward(G if G.ndim == 4 else G[:, None])
    return img


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_sample(workload, budget_s=12.0):
    """Times the oracle port on a bounded sample of the workload (about `budget_s` seconds)."""
    import neural_renderer_v2_pytorch_b200 as nr
    w = WORKLOADS[workload]
    cores = cpu_cores()
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    inp = make_inputs(w, seed=1000, device="cpu", nr=nr)
    oracle_step(inp, w, 1)                      # warm-up (builds / loads the C library)
    t0 = time.perf_counter()
    oracle_step(inp, w, 1)
    t1 = time.perf_counter() - t0
    views = int(max(1, min(w["views"], budget_s / max(t1, 1e-3))))
    reps, t0 = 0, time.perf_counter()
    while True:                                 # about budget_s seconds of CPU work
        oracle_step(inp, w, views)
        reps += 1
        t = time.perf_counter() - t0
        if t >= budget_s * 0.8 or reps >= 16:
            break
    return {"value": round(reps * views * w["S"] ** 2 / 1e6 / t, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d x (%d of %d views of %s, one fwd+bwd), %.1f s; C z-buffer/weight-map restatement (OpenMP) + "
                      "torch-CPU stages; the reference itself has no CPU z-buffer (rasterize_cuda.cpp:60-61)"
                      % (reps, views, w["views"], workload, t)}


def run_reference(args):
    """--impl reference: the oracle port of the reference's path on the host cores.
    (oracle/_ref holds the reference's CUDA kernels, which need a GPU; the reference ships no CPU
    implementation of the z-buffer, so the CPU arm is kind = "port".)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import neural_renderer_v2_pytorch_b200 as nr
    w = WORKLOADS[args.workload]
    cores = cpu_cores()
    torch.set_num_threads(cores)
    inp = make_inputs(w, seed=1000, device="cpu", nr=nr)
    oracle_step(inp, w, 1)
    t0 = time.perf_counter()
    oracle_step(inp, w, 1)
    t1 = time.perf_counter() - t0
    total_steps = args.steps + max(args.warmup, 1)
    views = int(max(1, min(w["views"], 150.0 / (t1 * total_steps))))
    for _ in range(max(args.warmup, 1)):
        oracle_step(inp, w, views)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(inp, w, views)
    t = time.perf_counter() - t0
    value = views * w["S"] ** 2 * args.steps / 1e6 / t
    sample = ("each step = %d of %d views of %s (fwd+bwd); C z-buffer/weight-map restatement (OpenMP) + "
              "torch-CPU stages" % (views, w["views"], args.workload))
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: teapot, %dx%d, anti_aliasing=%s, %s; bounded sample of %d views per step"
                                   % (args.workload, w["S"], w["S"], w["aa"], w["mode"], views)},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused-camera", action="store_true", help="cfg2r: camera transform as torch ops")
    ap.add_argument("--eager", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

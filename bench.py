#!/usr/bin/env python
"""bench.py — headline benchmark of the path-integration hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config C1..C5] [--spp S]
                  [--scaling weak|strong] [--no-extras]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[0], the configuration the metric is quoted on): C1 =
scene07 Cornell box, 600x600, 400 spp, depth 50, integrator 1 (Russian roulette).  One STEP is
one full render of that configuration per GPU (144 M camera paths, ~470 M rays).  --config
selects another BASELINE configuration (ray_tracing-rendering_b200/configs.py) as the headline;
--spp reduces the samples per pixel of a step (stated in config.workload and config.spp_run).

native arm      value   Mpaths/s with the scene resident in HBM: K steps bracketed by CUDA events
                        (barrier + synchronize on both sides, max over ranks); with N > 1 ranks
                        every rank renders 400 spp of its own sample slice (weak scaling) and the
                        images are reduced INSIDE the library: a kernel stages float3 means (the
                        division by spp fused in), ncclReduce on the library's own communicator
                        (rtb_comm_init / rtb_render_reduce), inside the timed region.
                e2e     the same metric through the host-buffer C-ABI: rtb_scene_upload (H2D of
                        the scene blob) + the render + D2H of the accumulators into pinned host
                        memory, every step.
                roofline for the dominant kernel against the roof that binds it (FP32 issue for
                        the shared-memory scenes, L2 for BVH scenes; the SURVEY 8(d) HBM figure is
                        kept as a secondary key), cpu_baseline = the unmodified reference's
                        Renderer::render on this box's host cores.
                configs every other BASELINE configuration (C2, C3, C4, C4env, C5), measured in the
                        same run on rank 0's GPU at its FULL sample count (N = 1 only; C5 = 4.2 s per step).
                strong  strong scaling inside the same run: C1 (400 spp in total) and C5 (4K,
                        1 M spheres, its full 1024 spp in total per step) with the work split over all N GPUs.
reference arm   --impl reference: the unmodified reference (oracle/_ref, compiled from
                /root/reference in the build container) on all host threads; rank 0 only.
Prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "ray_tracing-rendering_b200"

W, H, SPP, DEPTH, INTEGRATOR, SCENE = 600, 600, 400, 50, 1, 7
WORKLOAD = "scene07 Cornell box 600x600 spp=400 kMaxDepth=50 integrator 1 (Russian roulette)"
CONFIG = "C1"


def select_config(name, spp):
    """Points the module-level workload description at a BASELINE configuration."""
    global W, H, SPP, DEPTH, INTEGRATOR, SCENE, WORKLOAD, CONFIG
    cfg = importlib.import_module(PKG + ".configs").get(name)
    W, H, SPP, DEPTH, INTEGRATOR, SCENE, CONFIG = cfg.width, cfg.height, cfg.spp, cfg.depth, cfg.integrator, cfg.scene_id, name
    WORKLOAD = cfg.workload
    if spp and spp != cfg.spp:
        SPP = spp
        WORKLOAD += f" [step reduced to spp={spp} of {cfg.spp}]"
    return cfg


class ClockSampler(threading.Thread):
    """nvidia-smi clocks line of B200_PROFILING.md, sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.done, self.nvml = False, False
        self.t0 = self.t1 = None   # the timed region (samples outside it are dropped)

    def run(self):
        # NVML in-process (a sample every ~5 ms: the default C1 timed region is only ~90 ms long, less
        # than one period of `nvidia-smi -lms 100`, whose start-up alone can outlast the whole run);
        # the nvidia-smi loop of the recipe is the fallback when NVML cannot be loaded.
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = ((0x8, 5), (0x40, 6), (0x20, 7), (0x4, 8))   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap -> row columns
            self.nvml = True
            while not self.done:
                t = time.time()
                r = reasons_fn(h)
                row = [str(self.index), str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "", hex(r), "", "", "", ""]
                for bit, col in bits:
                    row[col] = "Active" if r & bit else "Not Active"
                self.rows.append([t] + row)
                time.sleep(0.005)
            return
        except Exception:
            if self.rows:
                return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([time.time()] + [c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.done = True
        if self.proc:
            self.proc.terminate()
        self.join(2)
        rows = [r[1:] for r in self.rows if self.t0 is None or self.t0 <= r[0] <= (self.t1 or r[0])]
        if not rows and self.rows:   # region shorter than one sampling period: the sample closest to it
            mid = 0.5 * ((self.t0 or 0) + (self.t1 or 0))
            rows = [min(self.rows, key=lambda r: abs(r[0] - mid))[1:]]
        sm = sorted(float(r[1]) for r in rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_reference(steps, warmup, budget_s):
    """The reference's CPU implementation of the path on all host threads:
    (Mpaths/s, cores, kind, sample text, secs per step, spp of a step).  kind "reference" = the unmodified
    reference compiled from /root/reference (oracle/_ref, Renderer::render untouched); where that
    library did not travel, kind "port" = the CPU restatement oracle/port on the same scene."""
    from oracle import refbind
    if SCENE < 0:
        raise RuntimeError("synthetic configuration: the reference has no such scene to time")
    if refbind.available():
        t_probe, _, _, _ = refbind.render_timed(SCENE, INTEGRATOR, W, 8, DEPTH)
        per_step = budget_s / max(steps + warmup, 1)
        spp = int(max(1, min(SPP, per_step / max(t_probe / 8, 1e-3))))
        for _ in range(warmup):
            refbind.render_timed(SCENE, INTEGRATOR, W, spp, DEPTH)
        secs = 0.0
        for _ in range(steps):
            t, _, _, _ = refbind.render_timed(SCENE, INTEGRATOR, W, spp, DEPTH)
            secs += t
        cores, kind = refbind.hardware_threads(), "reference"
        what = "Renderer::render of the unmodified reference"
    else:
        from oracle import portbind
        if not portbind.available():
            raise RuntimeError("neither oracle/_ref/libref_oracle.so nor oracle/liboracle_port.so is built")
        scenes = importlib.import_module(PKG + ".scenes")
        sc = portbind.PortScene(importlib.import_module(PKG + ".configs").get(CONFIG).blob())
        _, _, _, t_probe = sc.render_linear(INTEGRATOR, W, H, 2, DEPTH, want_images=False)
        per_step = budget_s / max(steps + warmup, 1)
        spp = int(max(1, min(SPP, per_step / max(t_probe / 2, 1e-3))))
        for _ in range(warmup):
            sc.render_linear(INTEGRATOR, W, H, spp, DEPTH, want_images=False)
        secs = 0.0
        for i in range(steps):
            secs += sc.render_linear(INTEGRATOR, W, H, spp, DEPTH, seed=i + 1, want_images=False)[3]
        cores, kind = portbind.hardware_threads(), "port"
        what = "CPU restatement oracle/port (brute-force closest hit, no BVH)"
    paths = W * H * spp * steps
    sample = f"{W}x{H} at {spp} of {SPP} spp per step, {steps} steps (cost is linear in spp); {what}, all hardware threads"
    return paths / secs / 1e6, cores, kind, sample, secs / steps, spp


L2_POLICY = ("native arm: no L2 flush between steps — the fused schedule (scenes of <= 64 primitive records) keeps the "
             "scene in shared memory and the path state in registers and rewrites every accumulator with atomics "
             "after a memset; the wavefront schedule streams its 8 Mi-entry queues (>= 1 GB read + written per "
             "iteration, evict-first) through the 126 MB L2.  reference arm: host CPU")


def config_dict(spp_run=None):
    """The SAME dict in both arms (the driver compares them; equal whenever the reference arm's time budget lets
    it run the full sample count): what differs between the arms lives under the top-level key `run`."""
    return {"workload": WORKLOAD, "name": CONFIG, "integrator": INTEGRATOR, "width": W, "height": H, "spp": SPP,
            "spp_run": SPP if spp_run is None else spp_run, "max_depth": DEPTH, "l2": L2_POLICY}


def run_reference(args, rank):
    if rank != 0:
        return
    value, cores, kind, sample, ms, spp_run = cpu_reference(args.steps, args.warmup, 120.0)
    out = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": config_dict(spp_run=spp_run),
           "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


class Bench:
    """One rank of the native arm: context, stream, library communicator, measurement helpers."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.pkg = importlib.import_module(PKG)
        self.binding = importlib.import_module(PKG + ".binding")
        self.configs = importlib.import_module(PKG + ".configs")
        self.abi = importlib.import_module(PKG + ".abi")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            # keep stdout to the ONE JSON line: NCCL prints its version banner to fd 1 when a
            # communicator comes up, so fd 1 points at stderr until both communicators are through
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize(self.dev)
                self.ctx = self.pkg.Context(local_rank)
                ids = [self.binding.Context.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                self.ctx.comm_init(world, rank, ids[0])      # the library's own NCCL communicator
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
        else:
            self.ctx = self.pkg.Context(local_rank)
        self.stream = torch.cuda.Stream(self.dev)   # kernels, events and the library's ncclReduce all run on this stream
        torch.cuda.synchronize(self.dev)
        torch.cuda.set_stream(self.stream)
        self.micro = load_json("profiles", "microbench.json")
        self.captured = load_json("profiles", "extend_traffic.json")
        self.hbm_peak, self.hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        peaks = load_json("MEASURED_PEAKS.json")
        if "hbm_gbs" in peaks:
            self.hbm_peak, self.hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"

    def sync(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def measure(self, name, spp_step, steps, warmup, strong, with_e2e=True, clocks=False):
        """K steps of configuration `name` at spp_step samples per pixel per step — per GPU (weak) or
        in total (strong).  Returns device-timed and end-to-end figures (max over ranks)."""
        torch, dist, ctx = self.torch, self.dist, self.ctx
        cfg = self.configs.get(name)
        w, h = cfg.width, cfg.height
        total_spp = spp_step if strong else spp_step * self.world
        blob = cfg.blob()                       # the product's own builder; no reference code
        ctx.upload_scene(blob)
        accum = torch.zeros((h, w, 4), dtype=torch.float32, device=self.dev) if self.world == 1 else None
        host = torch.empty((h, w, 4), dtype=torch.float32).pin_memory() if (self.rank == 0 and with_e2e) else None

        def params(flags=0, seed=1, spp=None):
            # N > 1: the library plans the split of the WHOLE job over the ranks (rtb_render_reduce)
            return ctx.params(w, h, total_spp if spp is None else spp, cfg.integrator, cfg.depth, 3, seed, 0, 1, 0, flags, 0, 1)

        def step(seed):
            if self.world == 1:
                return ctx.render_device(params(seed=seed), accum.data_ptr(), self.stream.cuda_stream)
            return ctx.render_reduce(params(seed=seed), stream=self.stream.cuda_stream)

        def step_e2e(seed):
            ctx.upload_scene(blob)                                       # H2D: the scene tables
            if self.world == 1:
                _, st = ctx.render(params(seed=seed), out=host.numpy())  # render + D2H of the accumulators
                return st
            return ctx.render_reduce(params(seed=seed), out=host.numpy() if self.rank == 0 else None,
                                     stream=self.stream.cuda_stream)

        def timed(fn):
            sampler = None
            if clocks:
                sampler = ClockSampler(self.local_rank)   # started before the warm-up: nvidia-smi takes a while to come up
                sampler.start()
            for i in range(warmup):
                fn(1000 + i)
            self.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            agg = {"paths": 0, "rays": 0, "launches": 0}
            if sampler:
                sampler.t0 = time.time()
            e0.record(self.stream)
            st = None
            for i in range(steps):
                st = fn(i + 1)
                agg["paths"] += st["paths"]
                agg["rays"] += st["rays_closest"] + st["rays_shadow"]
                agg["launches"] += st["kernel_launches"]
            e1.record(self.stream)
            self.sync()
            clk = None
            if sampler:
                sampler.t1 = time.time()
                clk = sampler.stop()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
            tot = torch.tensor([agg["paths"], agg["rays"], agg["launches"]], dtype=torch.float64, device=self.dev)
            if self.world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            return float(ms.item()), [float(x) for x in tot.tolist()], clk, st

        ms, tot, clk, last = timed(step)
        out = {"name": name, "workload": cfg.workload, "width": w, "height": h, "integrator": cfg.integrator,
               "spp": cfg.spp, "spp_run": total_spp if strong else spp_step, "spp_total_per_step": total_spp,
               "mpaths_per_s": tot[0] / (ms * 1e-3) / 1e6, "mrays_per_s": tot[1] / (ms * 1e-3) / 1e6,
               "ms_per_step": ms / steps, "gpu_launches": int(tot[2]), "schedule": "fused" if last["schedule"] == 1 else "wavefront",
               "clocks": clk, "blob_bytes": len(blob)}
        if with_e2e:
            ms2, tot2, _, _ = timed(step_e2e)
            out["e2e"] = {"value": tot2[0] / (ms2 * 1e-3) / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": len(blob),
                          "d2h_bytes_per_step": w * h * 16, "ms_per_step": ms2 / steps}
        self._last = (cfg, blob, accum, params)
        return out

    def roofline(self, name):
        """Dominant kernel of the configuration measured last (N = 1): per-launch CUDA-event durations
        measured live on the kernels' stream, algorithmic work from the same kernel's counting variant."""
        cfg, blob, accum, params = self._last
        ctx, binding = self.ctx, self.binding
        st_t = ctx.render_device(params(binding.RENDER_TIME_EXTEND, seed=77), accum.data_ptr(), self.stream.cuda_stream)
        st_c = ctx.render_device(params(binding.RENDER_COUNT_VISITS, seed=77, spp=min(cfg.spp, 16)), accum.data_ptr(),
                                 self.stream.cuda_stream)
        fused = st_t.get("schedule") == 1
        rays_c = max(st_c["rays_closest"] + st_c["rays_shadow"], 1)
        n_node, n_prim = st_c["nodes_visited"] / rays_c, st_c["prim_tests"] / rays_c
        prims = self.abi.parse_blob(blob)["prims"]["type"]
        frac_sphere = float((prims <= 1).mean()) if len(prims) else 0.0
        # fused: the one kernel traces closest-hit AND shadow rays; wavefront: k_extend traces the closest-hit rays
        rays_dom = st_t["rays_closest"] + (st_t["rays_shadow"] if fused else 0)
        secs = max(st_t["extend_ms"], 1e-9) * 1e-3
        grays = rays_dom / secs / 1e9
        launches = max(st_t["extend_launches"], 1)
        cap = self.captured.get(name, {})
        common = {"rays_per_launch": rays_dom / launches, "us_per_launch": 1e6 * secs / launches,
                  "share_of_step": st_t["extend_ms"] / st_t["device_ms"], "grays_per_s_in_kernel": grays,
                  "stage_ms": st_t.get("stage_ms"), "traffic": cap.get("dram_bytes_per_launch"), "ncu": cap.get("ncu")}
        ncu = cap.get("ncu") or {}
        if ncu.get("issue_slots_busy_pct") is not None:
            # the roof that actually binds these kernels (instruction issue, then the ALU pipe), from the committed
            # ncu capture of the same kernel on the same configuration (not re-measured by this run)
            common["issue_roof"] = {"frac": ncu["issue_slots_busy_pct"] / 100.0, "alu_pipe_frac": ncu.get("alu_pipe_pct", 0) / 100.0,
                                    "active_lanes_of_32": ncu.get("active_lanes_of_32"), "source": ncu.get("source")}
        if fused:
            # The scene (<= 64 records) lives in shared memory, path state in registers: nothing streams
            # from HBM (ncu: 5 MB of DRAM traffic per launch).  The kernel is bound by instruction issue;
            # its roof is the FP32 rate: algorithmic flops of SURVEY 8(d) — a box instance counted as ONE
            # slab test (24 flops), not as six rectangles — against the measured FP32 FMA peak.
            n_box = int(cap.get("box_instances", 2 if name in ("C1", "C3") else 0))
            n_rec = max(n_prim - 5.0 * n_box, 0.0)           # records tested per ray: 6 rects of a box = 1 record
            f_ray = 24.0 * n_box + (30.0 * frac_sphere + 10.0 * (1.0 - frac_sphere)) * (n_rec - n_box) + 60.0
            peak = float(self.micro.get("fp32_fma_tflops", 74.5))
            ach = grays * f_ray / 1e3
            rl = {"kernel": "k_fused (trace + shade + regenerate, scene in shared memory)", "bound": "fp32",
                  "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                  "peak_source": "tools/microbench FP32 FMA rate (profiles/microbench.json); nominal 74.5",
                  "flops_per_ray": f_ray, "records_per_ray": n_rec,
                  "note": (f"issue bound (ncu, profiles/r02_c1_fused_metrics.txt: {ncu.get('issue_slots_busy_pct', '?')} % of the issue slots, "
                           f"ALU pipe {ncu.get('alu_pipe_pct', '?')} %, {ncu.get('active_lanes_of_32', '?')} of 32 lanes on C1; see `issue_roof`): "
                           "the algorithmic flops are a fraction of the instructions a ray needs (compares, selects, RNG, control)")}
            b_ray = 32.0 * n_node + 32.0 * n_rec + 48.0
        else:
            # BVH scenes: the tree and the primitive records are served by L1 / L2 (they fit the 126 MB L2);
            # DRAM carries the streaming queues.  Algorithmic bytes per ray as SURVEY 8(d) defines them, with
            # this round's 64-byte nodes: 64 n_node + 32 n_prim + 48, against the measured L2 stream rate.
            node_bytes = 32.0 if st_t.get("traversal") == 1 else 64.0   # binary: n_node counts 32-byte child boxes; 4-wide: 64-byte nodes
            b_ray = node_bytes * n_node + 32.0 * n_prim + 48.0
            peak = float(self.micro.get("l2_stream_read_gbs", 21000.0))
            ach = grays * b_ray
            binary = st_t.get("traversal") == 1
            flat = st_t.get("traversal") == 0
            rl = {"kernel": ("k_extend (lockstep walk of the shared-memory scene copy: a scene of <= 64 records on the wavefront schedule)"
                             if flat else
                             "k_extend (binary while-while traversal: the per-scene rule keeps it where media / instances sit in the tree)"
                             if binary else "k_extend_w (warp-scheduled 4-wide traversal: closest hit, refill, material sort)"),
                  "bound": "l2", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                  "peak_source": "tools/microbench L2 stream read (profiles/microbench.json)",
                  "note": ("issue / latency bound, not bandwidth bound (ncu of this kernel on this configuration, `ncu` / `issue_roof`: "
                           f"issue slots {ncu.get('issue_slots_busy_pct', '?')} %, ALU pipe {ncu.get('alu_pipe_pct', '?')} %, "
                           f"{ncu.get('active_lanes_of_32', '?')} of 32 lanes; profiles/r02_c5_extend_w_metrics.txt, r02_c2_extend_metrics.txt); "
                           "DRAM traffic per launch (`traffic`) is the queue streaming, not the tree")}
        rl.update(common)
        rl.update({"bytes_per_ray": b_ray, "nodes_per_ray": n_node, "prims_per_ray": n_prim,
                   "hbm_8d": {"achieved": grays * b_ray, "peak": self.hbm_peak, "unit": "GB/s", "frac": grays * b_ray / self.hbm_peak,
                              "peak_source": self.hbm_src,
                              "note": "SURVEY 8(d) figure against the HBM copy rate; the bytes are served on chip, so this is not the binding roof"}})
        return rl


def run_native(args, rank, local_rank, world):
    B = Bench(args, rank, local_rank, world)
    strong = args.scaling == "strong"
    head = B.measure(CONFIG, SPP, args.steps, args.warmup, strong, with_e2e=True, clocks=True)
    roofline = B.roofline(CONFIG) if world == 1 else None
    cpu = None
    if rank == 0 and world == 1:
        try:
            v, cores, kind, sample, _, _ = cpu_reference(1, 0, 20.0)
            cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:  # the baseline is reported, never required for the product arm
            cpu = {"value": None, "unit": "Mpaths/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"unavailable: {e}"}
    extras, strong_runs = {}, {}
    if not args.no_extras:
        if world == 1:
            # the other BASELINE configurations, same run, same GPU (spp of a step stated per entry)
            # (every configuration at its FULL sample count; C5 = 1024 spp of a 4K image = 4.2 s per step: two steps)
            for name, steps in (("C2", 3), ("C3", 3), ("C4", 3), ("C4env", 3), ("C5", 2)):
                if name == CONFIG:
                    continue
                cfg = B.configs.get(name)
                m = B.measure(name, cfg.spp, steps, 1, False, with_e2e=True)
                m["roofline"] = B.roofline(name)
                extras[name] = m
        # strong scaling: the total work of a step is fixed, the library splits it over the N GPUs
        for name, spp in (("C1", 400), ("C5", 1024)):
            strong_runs[name] = B.measure(name, spp, 2 if name == "C5" else 5, 1, True, with_e2e=False)
    if rank == 0:
        out = {"metric": "Mpaths/s", "value": head["mpaths_per_s"], "unit": "Mpaths/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
               "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "impl": "native", "mrays_per_s": head["mrays_per_s"],
               "config": config_dict(head["spp_run"]),
               "run": {"spp_per_gpu": head["spp_total_per_step"] / world,
                       "paths_per_step_per_gpu": W * H * head["spp_total_per_step"] // world,
                       "parallelism": f"spp-split x{world}",
                       "collective": ("one ncclReduce(SUM) of float3 means per step, issued by the library "
                                      "(rtb_render_reduce), staging kernel fused with the division by spp") if world > 1 else "none",
                       "schedule": head["schedule"]},
               "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
               "roofline": roofline, "cpu_baseline": cpu, "configs": extras, "strong": strong_runs}
        print(json.dumps(out), flush=True)
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="C1")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-extras", action="store_true", help="headline only (skip the other configurations and the strong-scaling runs)")
    args = ap.parse_args()
    select_config(args.config, args.spp)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

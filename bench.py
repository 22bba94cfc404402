#!/usr/bin/env python
"""bench.py -- ERP frame-pairs/s through resample (2 images) + reproject + stitch on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --steps K --warmup W    (CPU port of the reference, all host cores)

Headline workload (BASELINE.json configs[1] / configs[3]): synthetic 2048x1024 ERP pairs, 20 tangent faces,
padding 0.3, 480x416 tangent images, normwarp blend, tangent flows injected on the device (the per-face DIS
estimator is a fixed CPU stage, excluded).  One step = one pass of the hot path over a batch of --batch pairs per GPU
(32 pairs: 1.9 GB of inputs per GPU, far larger than the 126 MB L2).  Pairs are independent, so ranks shard the batch
with no data-path collective inside the step (weak scaling: per-GPU batch fixed); at N>1 the run is repeated with the
one exchange the path has -- collecting the stitched flows on rank 0 -- fused into the stitch (`gather`).

The same JSON line also carries, under `configs`, short device-resident runs of the other shapes BASELINE.json names
(config #3: 3840x1920 with the flow_warp evaluation; config #5: 4096x2048 with the 80-face subdivision; the
straightforward blend at 2K), each with its per-kernel roofline fractions (N=1 only; --no-extra-configs skips them).
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ERP frame-pairs/s (2K ERP, 20 faces) resample+stitch"
UNIT = "pairs/s"
PAD, TANGENT_W = 0.3, 480


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frame pairs per GPU per step")
    ap.add_argument("--height", type=int, default=1024, help="ERP height (width = 2x)")
    ap.add_argument("--level", type=int, default=0, choices=[0, 1], help="1: the 80-face subdivision (config #5)")
    ap.add_argument("--blend", default="normwarp", choices=["straightforward", "normwarp"],
                    help="face_blending_method; normwarp is the reference pipeline's default (flow_estimate.py:104)")
    ap.add_argument("--warp", action="store_true", help="add warp_backward(tar, stitched flow) to the step (config #3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N>1: skip the run with the collection of the flows on rank 0")
    ap.add_argument("--gather-mode", default="copy", choices=["fused", "copy", "nccl"],
                    help="fused: the blend kernel stores into rank 0's peer-mapped buffer; copy: local result + "
                         "cudaMemcpyAsync to the peer buffer on a side stream; nccl: dist.gather after each step")
    ap.add_argument("--serial", action="store_true", help="launch the step's kernels on one stream (no overlap)")
    ap.add_argument("--kernel-reps", type=int, default=10)
    return ap.parse_args()


def workload_name(height, level, blend, warp):
    return ("synthetic %dx%d ERP pairs, %d tangent faces %s, padding 0.3, blend %s, wrap_around, tangent flows injected "
            "(DIS excluded)%s" % (2 * height, height, 80 if level else 20, "480 x {346,416,485,502}" if level else "480x416",
                                 blend, ", + warp_backward(tar, flow)" if warp else ""))


def workload_config(args, n_gpus):
    per_pair = 2 * args.height * 2 * args.height * 3 + 20 * 416 * 480 * 8
    return {"workload": workload_name(args.height, args.level, args.blend, args.warp),
            "pairs_per_gpu_per_step": args.batch, "global_pairs_per_step": args.batch * n_gpus,
            "sharding": "batch across ranks, no data-path collective inside the step (the collection of the stitched "
                        "flows on rank 0 is reported under `gather`)",
            "streams": "1 (serial)" if args.serial else "3 (resample src | resample tar | stitch)",
            "l2": "inputs (%.1f GB/GPU/step) exceed the 126 MB L2" % (args.batch * per_pair / 1e9)}


# ------------------------------------------------------------------------------------------------------
# reference arm: the CPU port of the reference on every host core
# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    runner = cpu_baseline.ParallelPair(args.height, TANGENT_W, PAD, args.blend)
    try:
        for _ in range(args.warmup):
            runner.step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            runner.step()
        dt = time.perf_counter() - t0
    finally:
        runner.close()
    value = args.steps / dt
    sample = "1 frame pair per step, the 20 faces of each stage spread over %d worker processes" % runner.cores
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": runner.cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is pure Python (numpy/scipy) and cannot travel to the GPU box; this is oracle/tif_oracle.py, "
                    "its numpy port pinned to the live reference by tests/test_oracle_golden.py"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every few ms through NVML while the timed region runs
    (nvidia-smi -lms cannot sample a 100 ms region)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.power = []
        self._stop = False
        self._thread = None
        self._nvml = None

    def _loop(self):
        n = self._nvml
        h = n.nvmlDeviceGetHandleByIndex(self.index)
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        try:
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        except Exception:
            pass
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(n.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml
            import threading
            pynvml.nvmlInit()
            self._nvml = pynvml
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    self.index = int(vis.split(",")[self.index])
                except Exception:
                    pass
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread is None:
            return out
        self._stop = True
        self._thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                       samples=len(self.samples), power_w_max=max(self.power) if self.power else None)
        return out


def bind_to_gpu_numa_node(index):
    """One process per GPU: run on the CPUs NVML reports as local to the GPU, so that the pinned staging buffers of
    the end-to-end pipeline are first-touched on that NUMA node.  Returns the number of CPUs bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def synth_device(torch, plan, batch, seed, device):
    """Synthetic pairs on the device: IID uint8 images, analytic tangent flows (level 1: the same
    formula with f/80 and every face's own height)."""
    import math
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    h, w = plan.erp_h, plan.erp_w
    src = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=device, generator=g)
    tar = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=device, generator=g)
    wt, nf = plan.tangent_w, plan.n_faces
    xx = torch.arange(wt, device=device, dtype=torch.float64).view(1, wt)
    parts = []
    for f, ht in enumerate(plan.face_heights):
        yy = torch.arange(ht, device=device, dtype=torch.float64).view(ht, 1)
        u = (4.0 * torch.sin(2 * math.pi * (xx / wt + f / nf)) + 1.5).expand(ht, wt)
        v = (3.0 * torch.cos(2 * math.pi * (yy / ht - f / nf)) - 0.5).expand(ht, wt)
        parts.append(torch.stack((u, v), -1).to(torch.float32).reshape(-1, 2))
    one = torch.cat(parts).reshape(1, plan.tangent_pixels, 2)
    jitter = torch.rand((batch, 1, 2), device=device, generator=g) - 0.5      # per-pair offset so pairs differ
    flows = (one + jitter).contiguous()
    return src, tar, flows


def load_traffic():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


class Workload:
    """One configuration resident on one GPU: plan, synthetic inputs, outputs, scratch, and its step."""

    def __init__(self, torch, pkg, device, height, level, blend_name, batch, seed, warp):
        self.torch, self.pkg, self.device = torch, pkg, device
        self.ops, self.lib = pkg._ops, pkg._lib
        self.height, self.level, self.blend_name, self.batch, self.warp = height, level, blend_name, batch, warp
        self.blend = self.lib.BLEND_NORMWARP if blend_name == "normwarp" else self.lib.BLEND_STRAIGHTFORWARD
        t0 = time.perf_counter()
        self.plan = plan = pkg._geometry.get_plan(PAD, height, TANGENT_W, device, level)
        torch.cuda.synchronize()
        self.plan_build_ms = 1e3 * (time.perf_counter() - t0)
        self.src, self.tar, self.flows = synth_device(torch, plan, batch, seed, device)
        self.tan_src = torch.empty((batch, plan.tangent_pixels, 4), dtype=torch.uint8, device=device)
        self.tan_tar = torch.empty_like(self.tan_src)
        self.out = torch.empty((batch, plan.erp_h, plan.erp_w, 2), dtype=torch.float32, device=device)
        self.warped = torch.empty_like(self.tar) if warp else None
        need = int(plan.lib.tif_ico2erp_scratch_bytes(plan.handle, batch, self.blend))
        self.scratch = torch.empty(max(need, 8), dtype=torch.uint8, device=device)
        self.side = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
        self.fork, self.joins = torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]
        # erp2ico x2; stitch = k_lift + k_gather, or (normwarp) k_lift_nw + k_nw_face_scale + k_blend_nw; + warp
        self.launches_per_step = (4 if self.blend == self.lib.BLEND_STRAIGHTFORWARD else 5) + (1 if warp else 0)
        self.out_target = self.out        # a raw peer address while the gather is fused into the stitch

    def _warp(self):
        lib = self.lib.load()
        p, t = self.plan, self.torch
        self.lib.check(lib.tif_warp_backward_u8(self.tar.data_ptr(), self.out.data_ptr(), self.lib.DTYPE_F32, self.batch, p.erp_h,
                                                p.erp_w, 3, self.lib.WARP_WRAP, self.warped.data_ptr(),
                                                t.cuda.current_stream().cuda_stream), "tif_warp_backward_u8")

    def step(self, serial=False):
        ops, plan, t = self.ops, self.plan, self.torch
        if serial:
            ops.erp2ico_out(plan, self.src, True, self.tan_src)
            ops.erp2ico_out(plan, self.tar, True, self.tan_tar)
            ops.ico2erp_flow_out(plan, self.flows, self.src, self.tar, self.blend, True, self.out_target, self.scratch)
        else:
            # the two resamplings and the stitch are independent of each other: separate streams let their blocks
            # share the SMs (`streams` in the JSON line reports what that buys against one stream)
            main = t.cuda.current_stream()
            self.fork.record(main)
            for s_, img, dst, ev in ((self.side[0], self.src, self.tan_src, self.joins[0]),
                                     (self.side[1], self.tar, self.tan_tar, self.joins[1])):
                s_.wait_event(self.fork)
                with t.cuda.stream(s_):
                    ops.erp2ico_out(plan, img, True, dst)
                    ev.record(s_)
            ops.ico2erp_flow_out(plan, self.flows, self.src, self.tar, self.blend, True, self.out_target, self.scratch)
            main.wait_event(self.joins[0])
            main.wait_event(self.joins[1])
        if self.warp:
            self._warp()

    def kernels(self):
        """name -> (launcher, algorithmic bytes per launch, launches per step); bytes per SURVEY.md section 8d."""
        ops, plan, lib = self.ops, self.plan, self.lib
        B, hw = self.batch, plan.erp_h * plan.erp_w
        P, Pref = plan.tangent_pixels, plan.referenced_tangent_pixels
        sf, nw = lib.BLEND_STRAIGHTFORWARD, lib.BLEND_NORMWARP
        a = (self.flows, self.src, self.tar)
        ks = {"k_erp2ico": (lambda: ops.erp2ico_out(plan, self.src, True, self.tan_src), B * (hw * 3 + P * 4), 2)}
        if self.blend == sf:
            # pass_mask: 8 = the lift kernel only, 4 = skip the lift (reuse its records), 1|2 = the gather
            ks["k_lift<straightforward>"] = (lambda: ops.ico2erp_flow_out(plan, a[0], None, None, sf, True, self.out, self.scratch, 8), B * 8 * Pref, 1)
            ks["k_gather<straightforward>"] = (lambda: ops.ico2erp_flow_out(plan, a[0], None, None, sf, True, self.out, self.scratch, 4 | 3), B * 8 * hw, 1)
        else:
            # k_lift_nw = end-point lift + target samples + warp error of every sample + face sums (reads the referenced
            # tangent flows and both ERP images); k_blend_nw = weights + blend (writes the ERP flow)
            ks["k_lift_nw"] = (lambda: ops.ico2erp_flow_out(plan, a[0], a[1], a[2], nw, True, self.out, self.scratch, 1), B * (8 * Pref + 6 * hw), 1)
            ks["k_blend_nw"] = (lambda: ops.ico2erp_flow_out(plan, a[0], a[1], a[2], nw, True, self.out, self.scratch, 2), B * 8 * hw, 1)
        if self.warp:
            ks["k_warp_backward"] = (self._warp, B * hw * (3 + 8 + 3), 1)
        return ks

    def step_bytes(self):
        return sum(nbytes * per_step for _, nbytes, per_step in self.kernels().values())


def time_steps(torch, dist, world, device, fn, steps, warmup, sampler=None, drain=None):
    """K steps between barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    barrier()
    if sampler is not None:
        sampler.start()
        time.sleep(0.02)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        fn()
    if drain is not None:       # work the steps left on a side stream belongs to the timed region
        torch.cuda.current_stream().wait_stream(drain)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler is not None else None
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, clocks


def kernel_rooflines(torch, wl, reps, peak_gbs, traffic):
    """Every kernel of the step alone: CUDA events around `reps` back-to-back launches (inputs >> L2)."""
    out = {}
    key = "h%d_l%d" % (wl.height, wl.level)
    per_pair = traffic.get("bytes_per_pair", {}).get(key, {})
    for name, (fn, nbytes, per_step) in wl.kernels().items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        gbs = nbytes / (ms * 1e-3) / 1e9
        tr = per_pair.get(name)
        out[name] = {"ms": ms, "launches_per_step": per_step, "algorithmic_bytes": nbytes, "gbs": gbs, "frac": gbs / peak_gbs,
                     "traffic": tr * wl.batch if tr else None,
                     "dram_over_algorithmic": (tr * wl.batch / nbytes) if tr else None}
    return out


def roofline_block(wl, ktimes, ms_per_step, peak_gbs, peak_src):
    weight = {k: v["ms"] * v["launches_per_step"] for k, v in ktimes.items()}
    dominant = max(weight, key=weight.get)
    step_gbs = wl.step_bytes() / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": dominant, "achieved": ktimes[dominant]["gbs"], "peak": peak_gbs, "unit": "GB/s",
            "frac": ktimes[dominant]["frac"], "traffic": ktimes[dominant]["traffic"], "peak_source": peak_src,
            "share_of_step": weight[dominant] / sum(weight.values()),
            "step_frac": step_gbs / peak_gbs, "step_algorithmic_bytes": wl.step_bytes(), "step_gbs": step_gbs,
            "sum_of_kernel_ms": sum(weight.values()), "kernels": ktimes}


def pcie_probe(torch, device, mib=256, reps=6):
    """Raw pinned-memory copy ceiling of this rank's link: H2D alone, D2H alone, both at once."""
    n = mib << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=device), torch.empty(n, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(h2d, d2h):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(device)
        return n * reps / (time.perf_counter() - t0) / 1e9
    run(True, True)
    return {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "both_gbs_each_direction": run(True, True)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cpu_binding = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created; stdout carries exactly one
        # JSON line, so route fd 1 to stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    entry.build()
    pkg = entry.load_package()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    traffic = load_traffic()

    B = args.batch
    wl = Workload(torch, pkg, device, args.height, args.level, args.blend, B, 1234 + rank, args.warp)
    plan = wl.plan

    # ---- the timed steps ---------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step, clocks = time_steps(torch, dist, world, device, lambda: wl.step(args.serial), args.steps, args.warmup, sampler)
    value = world * B / (ms_per_step * 1e-3)
    # what the three streams buy: the same step on one stream (a few steps, not part of `value`)
    alt_ms, _ = time_steps(torch, dist, world, device, lambda: wl.step(not args.serial), max(3, args.steps // 4), 2)
    streams = {"overlapped_ms_per_step": alt_ms if args.serial else ms_per_step,
               "serial_ms_per_step": ms_per_step if args.serial else alt_ms}
    streams["serial_over_overlapped"] = streams["serial_ms_per_step"] / streams["overlapped_ms_per_step"]

    # ---- per-kernel timing (CUDA events on the launching stream) and roofline ------------------------------
    ktimes = kernel_rooflines(torch, wl, args.kernel_reps, peak_gbs, traffic)
    roofline = roofline_block(wl, ktimes, ms_per_step, peak_gbs, peak_src)
    wl.step(True)
    torch.cuda.synchronize()
    reference_out = wl.out.clone()

    # ---- the path's only exchange: the stitched flows of all ranks collected on rank 0 ---------------------
    gather = None
    if world > 1 and not args.no_gather:
        gather = run_with_gather(args, torch, dist, pkg, wl, world, rank, device, ms_per_step)

    # ---- end to end through the public host API (pinned host buffers, copies inside the timed region) -------
    e2e = None
    if not args.no_e2e and args.level == 0:
        e2e = run_e2e(args, torch, dist, pkg, wl, world, rank, device, reference_out)

    # ---- the other configurations BASELINE.json names, device resident, short (N=1) -------------------------
    configs = None
    if world == 1 and not args.no_extra_configs and args.height == 1024 and args.level == 0:
        del reference_out
        configs = {}
        extra = [("config2_2048x1024_straightforward", 1024, 0, "straightforward", 32, False),
                 ("config3_3840x1920_normwarp_warp", 1920, 0, "normwarp", 16, True),
                 ("config5_4096x2048_level1_80faces", 2048, 1, "normwarp", 8, False)]
        for name, height, level, blend, batch, warp in extra:
            torch.cuda.empty_cache()
            w2 = Workload(torch, pkg, device, height, level, blend, batch, 99, warp)
            ms2, _ = time_steps(torch, dist, 1, device, lambda: w2.step(False), 6, 3)
            k2 = kernel_rooflines(torch, w2, 5, peak_gbs, traffic)
            rf = roofline_block(w2, k2, ms2, peak_gbs, peak_src)
            configs[name] = {"workload": workload_name(height, level, blend, warp), "pairs_per_step": batch, "steps": 6,
                             "value": batch / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "step_frac": rf["step_frac"],
                             "plan_build_ms": w2.plan_build_ms, "plan_device_bytes": w2.plan.device_bytes,
                             "kernels": {k: {"ms": v["ms"], "frac": v["frac"], "gbs": v["gbs"], "algorithmic_bytes": v["algorithmic_bytes"],
                                             "traffic": v["traffic"], "dram_over_algorithmic": v["dram_over_algorithmic"]}
                                         for k, v in k2.items()}}
            del w2
            if level == 1:
                pkg._geometry.clear_plans()

    # ---- CPU baseline: the oracle port on one host core, bounded sample ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.level == 0:
        from oracle import cpu_baseline, synth
        s_flows = synth.synth_flows(TANGENT_W)
        n_sample, dt = 0, 0.0
        while dt < 10.0 and n_sample < 8:          # a bounded sample: at least 10 s of CPU work, at most 8 pairs
            s_src, s_tar = synth.synth_images(args.height, n_sample)
            dt += cpu_baseline.pair_serial(s_src, s_tar, s_flows, args.height, TANGENT_W, PAD, args.blend)
            n_sample += 1
        cpu = {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d frame pairs %dx%d (%s), oracle/tif_oracle.py on one core, %.1f s" % (n_sample, 2 * args.height, args.height, args.blend, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 images / f32 flow (f64 geometry plan)", "data": "synthetic", "config": workload_config(args, world),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": wl.launches_per_step * args.steps,
                "clocks": clocks, "streams": streams, "gather": gather, "configs": configs,
                "plan": {"build_ms_outside_timed_region": wl.plan_build_ms, "device_bytes": plan.device_bytes,
                         "max_faces_per_pixel": plan.max_faces_per_pixel, "accepted": plan.accepted_unique,
                         "referenced_tangent_pixels": plan.referenced_tangent_pixels},
                "cpus_bound_per_rank": cpu_binding}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_with_gather(args, torch, dist, pkg, wl, world, rank, device, ms_plain):
    """The timed steps again with the collection of the stitched flows on rank 0 on the path."""
    B, plan = wl.batch, wl.plan
    n_items = B * world
    total_bytes = (world - 1) * B * plan.erp_h * plan.erp_w * 8
    info = {"mode": args.gather_mode, "bytes_into_rank0_per_step": total_bytes}
    steps, warmup = args.steps, args.warmup
    if args.gather_mode == "nccl":
        def step():
            wl.step(args.serial)
            pkg.sharding.gather_erp_flow(wl.out, n_items, dst=0)
        ms, _ = time_steps(torch, dist, world, device, step, steps, warmup)
        info["note"] = "blocking dist.gather (NCCL send/recv into rank 0) after every step"
    else:
        buf = pkg.sharding.PeerGatherBuffer(n_items, plan.erp_h, plan.erp_w, dst=0, device=device)
        lib = pkg._lib.load()
        nbytes = B * plan.erp_h * plan.erp_w * 8
        if args.gather_mode == "fused":
            wl.out_target = buf.slice_ptr()           # the blend kernel's stores cross NVLink themselves
            ms, _ = time_steps(torch, dist, world, device, lambda: wl.step(args.serial), steps, warmup)
            wl.out_target = wl.out
            info["note"] = ("peer-store epilogue: every rank's blend kernel writes its [B,H,W,2] shard straight into rank 0's "
                            "CUDA-IPC-mapped buffer (no staging, no receive side)")
        else:
            # Two local result buffers: the push of step i (side stream, copy engine) overlaps all of step i + 1, and
            # step i + 2 waits until the push of step i has read its buffer.  No rank waits for another: the pushes of
            # the ranks drift apart until rank 0's NVLink ingress is busy all the time, which is the bound of this
            # exchange (`limiter`).
            copy_stream = torch.cuda.Stream(device)
            outs = [wl.out, torch.empty_like(wl.out)]
            done, free = torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]
            state = {"i": 0}

            def step():
                main = torch.cuda.current_stream()
                k = state["i"] & 1
                if state["i"] >= 2:
                    main.wait_event(free[k])          # the push of step i - 2 has read this buffer
                state["i"] += 1
                wl.out_target = outs[k]
                wl.step(args.serial)
                done.record(main)
                copy_stream.wait_event(done)
                with torch.cuda.stream(copy_stream):
                    pkg._lib.check(lib.tif_ipc_copy_async(buf.slice_ptr(), outs[k].data_ptr(), nbytes, copy_stream.cuda_stream), "tif_ipc_copy_async")
                    free[k].record(copy_stream)

            def timed_step():
                step()
            ms, _ = time_steps(torch, dist, world, device, timed_step, steps, warmup, drain=copy_stream)
            wl.out_target = wl.out
            info["note"] = ("local result (two buffers), then cudaMemcpyAsync into rank 0's peer-mapped buffer on a side stream: the push of "
                            "step i overlaps step i + 1; the timed region ends when the last push has landed")
        # the collected field on rank 0 must equal what every rank computed locally
        torch.cuda.synchronize()
        dist.barrier()
        wl.step(True)
        torch.cuda.synchronize()
        mine = wl.out.clone()
        if args.gather_mode == "fused":
            wl.out_target = buf.slice_ptr()
            wl.step(True)
            wl.out_target = wl.out
        else:
            pkg._lib.check(lib.tif_ipc_copy_async(buf.slice_ptr(), mine.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream), "tif_ipc_copy_async")
        torch.cuda.synchronize()
        dist.barrier()
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            full = buf.tensor()
            info["matches_nccl_gather"] = bool(all(torch.equal(full[r * B:(r + 1) * B], parts[r]) for r in range(world)))
            del full
        del parts, mine
        buf.close()
    value = world * B / (ms * 1e-3)
    ingress = total_bytes / (ms * 1e-3) / 1e9
    info.update({"value": value, "unit": UNIT, "ms_per_step": ms, "ms_per_step_without": ms_plain,
                 "rank0_ingress_gbs": ingress, "nvlink_peak_gbs_per_direction": 900.0, "nvlink_measured_peer_copy_gbs": 770.0,
                 "ingress_frac_of_measured": ingress / 770.0,
                 "limiter": "NVLink ingress of rank 0: %.2f GB per step / 770 GB/s measured peer copy = %.2f ms per step at best" %
                            (total_bytes / 1e9, total_bytes / 770e9 * 1e3)})
    return info


def run_e2e(args, torch, dist, pkg, wl, world, rank, device, reference_out):
    B = wl.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    probe = pcie_probe(torch, device)
    if world > 1:       # all ranks probed at once: the box's aggregate ceiling
        t = torch.tensor([probe["both_gbs_each_direction"]], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        probe["all_ranks_both_gbs_each_direction"] = float(t.item())
    h_src, h_tar, h_flows = (t.cpu().pin_memory() for t in (wl.src, wl.tar, wl.flows))
    h_out = torch.empty(wl.out.shape, dtype=wl.out.dtype).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def measure(tangent_format):
        pipe = pkg.pipeline.HostPipeline(PAD, args.height, TANGENT_W, args.blend, True, chunk=2, slots=4, device=device,
                                         tangent_format=tangent_format)
        h_ts = torch.empty(pipe.tangent_out_shape(B), dtype=torch.uint8).pin_memory()
        h_tt = torch.empty(pipe.tangent_out_shape(B), dtype=torch.uint8).pin_memory()
        for _ in range(2):
            pipe.run(h_src, h_tar, h_flows, h_out, h_ts, h_tt)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.run(h_src, h_tar, h_flows, h_out, h_ts, h_tt)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        ok = bool(torch.equal(h_out.to(device), reference_out))
        h2d, d2h = B * pipe.h2d_bytes_per_pair(), B * pipe.d2h_bytes_per_pair()
        per_step = dt / e2e_steps
        ceiling = max(h2d, d2h) / (probe["both_gbs_each_direction"] * 1e9)
        return {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": pipe.describe(), "matches_device_path": ok,
                "h2d_gbs": h2d / per_step / 1e9, "d2h_gbs": d2h / per_step / 1e9, "pcie_frac": ceiling / per_step}

    res = measure("rgb")            # the default call
    res["pcie_probe"] = probe
    res["pcie_note"] = ("pcie_frac = time the larger direction needs at this rank's measured concurrent-copy rate / measured time per "
                        "step; the path is PCIe-bound, so only fewer bytes raise it: `variants.gray` returns the tangent stacks as the "
                        "grey images the reference's DIS wrapper computes first (flow_estimate.py:31-47), a third of the bytes")
    res["variants"] = {"gray": measure("gray")}
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

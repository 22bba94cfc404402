#!/usr/bin/env python
"""bench.py -- ERP frame-pairs/s through resample (2 images) + reproject + stitch on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --steps K --warmup W    (CPU port of the reference, all host cores)

Workload (BASELINE.json configs[1] / configs[3]): synthetic 2048x1024 ERP pairs, 20 tangent faces,
padding 0.3, 480x416 tangent images, tangent flows injected on the device (the per-face DIS estimator is
a fixed CPU stage, excluded).  One step = one pass of the hot path over a batch of --batch pairs per GPU
(32 pairs: 1.9 GB of inputs per GPU, far larger than the 126 MB L2).  Pairs are independent, so ranks
shard the batch with no data-path collective (weak scaling: per-GPU batch fixed).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
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
    ap.add_argument("--blend", default="normwarp", choices=["straightforward", "normwarp"],
                    help="face_blending_method; normwarp is the reference pipeline's default (flow_estimate.py:104)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial", action="store_true", help="launch the step's kernels on one stream (no overlap)")
    ap.add_argument("--kernel-reps", type=int, default=10)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "synthetic %dx%d ERP pairs, 20 tangent faces 480x416, padding 0.3, blend %s, wrap_around, "
                        "tangent flows injected (DIS excluded)" % (2 * args.height, args.height, args.blend),
            "pairs_per_gpu_per_step": args.batch, "global_pairs_per_step": args.batch * n_gpus,
            "sharding": "batch across ranks, no data-path collective",
            "streams": "1 (serial)" if args.serial else "3 (resample src | resample tar | stitch overlap)", "l2": "inputs (%.1f GB/GPU/step) exceed the 126 MB L2"
            % (args.batch * (2 * args.height * 2 * args.height * 3 + 20 * 416 * 480 * 8) / 1e9)}


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
    the end-to-end pipeline are first-touched on that NUMA node (the host<->device copies of the ranks otherwise
    share one node's memory controllers).  Returns the number of CPUs bound to, or None."""
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
    """Synthetic pairs on the device: IID uint8 images, analytic tangent flows."""
    import math
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    h, w = plan.erp_h, plan.erp_w
    src = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=device, generator=g)
    tar = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=device, generator=g)
    ht, wt, nf = plan.tangent_h, plan.tangent_w, plan.n_faces
    yy = torch.arange(ht, device=device, dtype=torch.float64).view(1, ht, 1)
    xx = torch.arange(wt, device=device, dtype=torch.float64).view(1, 1, wt)
    ff = torch.arange(nf, device=device, dtype=torch.float64).view(nf, 1, 1) / nf
    u = (4.0 * torch.sin(2 * math.pi * (xx / wt + ff)) + 1.5).expand(nf, ht, wt)
    v = (3.0 * torch.cos(2 * math.pi * (yy / ht - ff)) - 0.5).expand(nf, ht, wt)
    one = torch.stack((u, v), -1).to(torch.float32).reshape(1, plan.tangent_pixels, 2)
    jitter = torch.rand((batch, 1, 2), device=device, generator=g) - 0.5      # per-pair offset so pairs differ
    flows = (one + jitter).contiguous()
    return src, tar, flows


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
    ops, lib = pkg._ops, pkg._lib
    blend = lib.BLEND_NORMWARP if args.blend == "normwarp" else lib.BLEND_STRAIGHTFORWARD

    t0 = time.perf_counter()
    plan = pkg._geometry.get_plan(PAD, args.height, TANGENT_W, device)
    torch.cuda.synchronize()
    plan_build_ms = 1e3 * (time.perf_counter() - t0)

    B = args.batch
    src, tar, flows = synth_device(torch, plan, B, 1234 + rank, device)
    tan_src = torch.empty((B, plan.tangent_pixels, 4), dtype=torch.uint8, device=device)
    tan_tar = torch.empty_like(tan_src)
    out = torch.empty((B, plan.erp_h, plan.erp_w, 2), dtype=torch.float32, device=device)
    scratch = torch.empty(max(int(plan.lib.tif_ico2erp_scratch_bytes(plan.handle, B, lib.BLEND_NORMWARP)), 8),
                          dtype=torch.uint8, device=device)
    # erp2ico x2; stitch = k_lift + k_gather (normwarp: k_lift, k_gather<sums>, k_gather<blend>)
    launches_per_step = 4 if blend == lib.BLEND_STRAIGHTFORWARD else 6

    # The two resamplings and the stitch are independent of each other: they run on separate streams so that
    # their blocks share the SMs (each kernel alone is issue/latency bound and leaves slots idle).
    side = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
    fork, joins = torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]

    def step():
        main = torch.cuda.current_stream()
        if args.serial:
            ops.erp2ico_out(plan, src, True, tan_src)
            ops.erp2ico_out(plan, tar, True, tan_tar)
            ops.ico2erp_flow_out(plan, flows, src, tar, blend, True, out, scratch)
            return
        fork.record(main)
        for s_, img, dst, ev in ((side[0], src, tan_src, joins[0]), (side[1], tar, tan_tar, joins[1])):
            s_.wait_event(fork)
            with torch.cuda.stream(s_):
                ops.erp2ico_out(plan, img, True, dst)
                ev.record(s_)
        ops.ico2erp_flow_out(plan, flows, src, tar, blend, True, out, scratch)
        main.wait_event(joins[0])
        main.wait_event(joins[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.02)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- per-kernel timing (CUDA events on the launching stream) and roofline ------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    hw = plan.erp_h * plan.erp_w
    P, Pref = plan.tangent_pixels, plan.referenced_tangent_pixels
    sf, nw = lib.BLEND_STRAIGHTFORWARD, lib.BLEND_NORMWARP
    # pass_mask: 1 = lift + normwarp face sums, 2 = blend, 4 = skip the lift kernel (reuse the records), 8 = lift only
    kernels = {
        "k_erp2ico": (lambda: ops.erp2ico_out(plan, src, True, tan_src), B * (hw * 3 + P * 4)),
        "k_lift<straightforward>": (lambda: ops.ico2erp_flow_out(plan, flows, None, None, sf, True, out, scratch, 8),
                                    B * (8 * Pref)),
        "k_gather<straightforward>": (lambda: ops.ico2erp_flow_out(plan, flows, None, None, sf, True, out, scratch, 4 | 3),
                                      B * (8 * hw)),
        "k_lift<normwarp>": (lambda: ops.ico2erp_flow_out(plan, flows, src, tar, nw, True, out, scratch, 8),
                             B * (8 * Pref + 3 * hw)),
        "k_gather<sums>": (lambda: ops.ico2erp_flow_out(plan, flows, src, tar, nw, True, out, scratch, 1 | 4), B * (3 * hw)),
        "k_gather<blend>": (lambda: ops.ico2erp_flow_out(plan, flows, src, tar, nw, True, out, scratch, 2), B * (8 * hw)),
    }
    ktimes = {}
    for name, (fn, nbytes) in kernels.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.kernel_reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.kernel_reps
        ktimes[name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9,
                        "frac": nbytes / (ms * 1e-3) / 1e9 / peak_gbs}
    used = ["k_erp2ico"] + (["k_lift<straightforward>", "k_gather<straightforward>"] if blend == sf
                            else ["k_lift<normwarp>", "k_gather<sums>", "k_gather<blend>"])
    weight = {k: ktimes[k]["ms"] * (2 if k == "k_erp2ico" else 1) for k in used}
    dominant = max(weight, key=weight.get)
    traffic = None
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum per frame pair from the committed `ncu --set full` capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("erp_h") == args.height and dominant in tr.get("bytes_per_pair", {}):
            traffic = tr["bytes_per_pair"][dominant] * B
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": ktimes[dominant]["gbs"], "peak": peak_gbs, "unit": "GB/s",
                "frac": ktimes[dominant]["frac"], "traffic": traffic, "peak_source": peak_src,
                "share_of_step": weight[dominant] / sum(weight.values()), "kernels": ktimes}

    # ---- end to end through the public host API (pinned host buffers, copies inside the timed region) -------
    e2e = None
    if not args.no_e2e:
        pipe = pkg.pipeline.HostPipeline(PAD, args.height, TANGENT_W, args.blend, True, chunk=2, slots=4, device=device)
        h_src, h_tar, h_flows = (t.cpu().pin_memory() for t in (src, tar, flows))
        h_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        h_ts = torch.empty(tan_src.shape, dtype=torch.uint8).pin_memory()
        h_tt = torch.empty(tan_src.shape, dtype=torch.uint8).pin_memory()
        e2e_steps = max(3, min(args.steps, 10))
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
        ok = bool(torch.allclose(h_out.to(device), out, atol=0, rtol=0))
        e2e = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * pipe.h2d_bytes_per_pair(),
               "d2h_bytes_per_step": B * pipe.d2h_bytes_per_pair(), "steps": e2e_steps,
               "api": "pipeline.HostPipeline.run (pinned host in/out; D2H returns both tangent stacks and the ERP flow)",
               "matches_device_path": ok}
        del pipe, h_src, h_tar, h_flows, h_out, h_ts, h_tt

    # ---- optional collection of the stitched flows on rank 0 (the path's only exchange) --------------------
    gather = None
    if world > 1:
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pkg.sharding.gather_erp_flow(out, B * world, dst=0)
        barrier()
        a.record()
        pkg.sharding.gather_erp_flow(out, B * world, dst=0)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather = {"ms": float(t.item()), "bytes_to_rank0": (world - 1) * out.numel() * 4,
                  "note": "NCCL gather of [B,H,W,2] f32 to rank 0, outside the timed steps"}

    # ---- CPU baseline: the oracle port on one host core, bounded sample ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline, synth
        s_src, s_tar = synth.synth_images(args.height, 0)
        s_flows = synth.synth_flows(TANGENT_W)
        dt = cpu_baseline.pair_serial(s_src, s_tar, s_flows, args.height, TANGENT_W, PAD, args.blend)
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "1 frame pair %dx%d (%s), oracle/tif_oracle.py on one core, %.1f s" % (2 * args.height, args.height, args.blend, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 images / f32 flow (f64 geometry plan)", "data": "synthetic", "config": workload_config(args, world),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "clocks": clocks, "plan": {"build_ms_outside_timed_region": plan_build_ms, "device_bytes": plan.device_bytes,
                                           "max_faces_per_pixel": plan.max_faces_per_pixel, "accepted": plan.accepted_unique,
                                           "referenced_tangent_pixels": Pref},
                "gather": gather, "cpus_bound_per_rank": cpu_binding}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

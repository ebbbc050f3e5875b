#!/usr/bin/env python3
"""bench.py -- headline benchmark of the CDSPResampler::process() path on B200.

Metric (BASELINE.json / SURVEY.md section 8d): input Msamples/s = 1e-6 * frames * channels / seconds
inside process() -- the reference's "Mrops" (bench/r8bfreesrc.cpp:95,140-141).

A "step" is one process() call over one batch: `channels` independent fp64 streams, `block`
(=65536, the reference's MaxInLen in BASELINE's configs) new samples each.  The default workload is
BASELINE configs[1]: 1024 channels, 44100 -> 96000, CDSPResampler24 (2 % transition band).

  value      inputs already resident in HBM, K steps timed with CUDA events on the launch stream
  e2e        the same K steps through r8bgpu_batch_process_host(): pinned HOST buffers, H2D of the
             block and D2H of the produced samples inside the timed region
  roofline   dominant kernel's algorithmic bytes / its CUDA-event time, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU path (oracle/_ref, R8B_PFFFT_DOUBLE, -O3 -mavx2 -mfma) on all
             host threads over a bounded sample of the same workload

N > 1 (torchrun): channels are sharded over ranks (weak scaling: `channels` per GPU), no data-path
collective; barrier + synchronize on both sides, max over ranks.

--impl reference: times ONLY the reference CPU implementation (rank 0), same metric/config.
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

WORKLOADS = {
    # name: (src, dst, atten, tb, extfft, default channels per GPU)
    "cfg2_1024ch_44100_96000_r24": (44100.0, 96000.0, 180.15, 2.0, 0, 1024),
    "cfg3_1024ch_48000_44100_r24": (48000.0, 44100.0, 180.15, 2.0, 0, 1024),
    "cfg5_512ch_48000_47999_r24": (48000.0, 47999.0, 180.15, 2.0, 0, 512),
    "cfg4_128ch_44100_2822400_r24_extfft": (44100.0, 2822400.0, 180.15, 2.0, 1, 128),
    "cfg3b_1024ch_192000_44100_r24": (192000.0, 44100.0, 180.15, 2.0, 0, 1024),
    "cfg3c_1024ch_2822400_44100_r24": (2822400.0, 44100.0, 180.15, 2.0, 0, 1024),
}
# algorithmic flops per input sample (real-FFT 2.5 N log2 N per block + interpolation / half-band MACs)
FLOPS_PER_IN_SAMPLE = {"cfg2_1024ch_44100_96000_r24": 241.0, "cfg3_1024ch_48000_44100_r24": 188.0,
                       "cfg5_512ch_48000_47999_r24": 281.0, "cfg4_128ch_44100_2822400_r24_extfft": 855.0}
DEFAULT_WORKLOAD = "cfg2_1024ch_44100_96000_r24"
CPU_KEYS = ("value", "unit", "cores", "kind", "sample", "per_core", "sample_channels", "single_thread")
BLOCK = 65536


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [t.strip() for t in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    """Threads the CPU baseline may use: the affinity mask, capped by a cgroup CPU quota if one is set."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                t = f.read().split()
            if path.endswith("cpu.max"):
                if t[0] != "max":
                    n = min(n, max(1, int(float(t[0]) / float(t[1]))))
            else:
                q = int(t[0])
                if q > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                        n = min(n, max(1, q // int(f.read())))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


def bind_to_numa_node(node):
    """Pin this process to the CPUs of `node` (inside its current affinity mask).  Returns True when bound."""
    if node is None or node < 0:
        return False
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except (OSError, ValueError):
        return False


def synth_block(n_ch, block, seed):
    """Planar fp64 white noise in [-1,1), per-channel stream."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=(n_ch, block))


def cpu_reference_run(src, dst, tb, atten, extfft, threads, target_seconds=8.0, single_thread=True):
    """Time the compiled reference (oracle/_ref) on host cores over a bounded sample."""
    import oracle_util as ou
    flavor = ("e1" if extfft else "e0") + ("_fast" if ou.cpu_supports_fast() and ou.have_ref(("e1" if extfft else "e0") + "_fast") else "")
    if not ou.have_ref(flavor):
        return None
    ref = ou.RefOracle(flavor)
    n_ch = max(1, threads) * 2
    x = synth_block(n_ch, BLOCK, 99)
    # calibrate with 1 call, then size the sample to ~target_seconds of wall time
    t1, _, _ = ref.bench(src, dst, BLOCK, tb, atten, x, 1, 1, threads)
    calls = int(max(2, min(8192, target_seconds / max(t1, 1e-4))))
    secs, n_out, _ = ref.bench(src, dst, BLOCK, tb, atten, x, 1, calls, threads)
    val = 1e-6 * n_ch * BLOCK * calls / secs
    res = {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "reference",
           "sample": "%d ch x %d calls x %d frames, %s, %d threads, %.2f s" % (n_ch, calls, BLOCK, ref.name, threads, secs),
           "per_core": val / max(1, threads), "sample_channels": n_ch, "calls": calls, "secs": secs,
           "sample_ms_per_call": secs / calls * 1e3}
    if single_thread:
        # BASELINE configs[0]: one channel, one thread (what example.cpp / README.md:111-114 quote per core)
        t1, _, _ = ref.bench(src, dst, BLOCK, tb, atten, x[:1], 1, 1, 1)
        c1 = int(max(2, min(2048, 2.0 / max(t1, 1e-4))))
        s1, _, _ = ref.bench(src, dst, BLOCK, tb, atten, x[:1], 1, c1, 1)
        res["single_thread"] = {"value": 1e-6 * BLOCK * c1 / s1, "unit": "Msamples/s", "cores": 1,
                                "sample": "1 ch x %d calls x %d frames, %s, %.2f s" % (c1, BLOCK, ref.name, s1)}
    return res


def verify_against_oracle(src, dst, tb, atten, extfft, xs, calls, got_last, channels):
    """Replay, on a few sampled channels, every process() call the batch has seen since it was created through the
    oracle (the compiled reference when present, else the C port) and compare what the LAST timed call wrote:
    counts equal, max |d| <= 32 eps * max|y|, rms d <= 4 eps * rms y (the parity tolerance of tests/)."""
    import numpy as np
    import oracle_util as ou
    ref = ou.best_oracle(extfft)
    worst_max, worst_rms, ok = 0.0, 0.0, True
    for k, ch in enumerate(channels):
        r = ref.Resampler(src, dst, BLOCK, tb, atten)
        y = None
        for which in calls:
            y = r.process(xs[which][k])
        if y is None or len(y) != got_last.shape[1]:
            return {"ok": False, "oracle": ref.name, "why": "count mismatch on channel %d: oracle %s, device %d"
                    % (ch, None if y is None else len(y), got_last.shape[1])}
        mx, rm = ou.parity_metrics(got_last[k], y)
        worst_max, worst_rms = max(worst_max, mx), max(worst_rms, rm)
        ok = ok and mx <= 32 * ou.EPS and rm <= 4 * ou.EPS and bool(np.all(np.isfinite(got_last[k])))
    return {"ok": bool(ok), "oracle": ref.name, "channels": list(channels), "calls_replayed": len(calls),
            "samples_compared": int(got_last.size), "max_err_eps": worst_max / ou.EPS, "rms_err_eps": worst_rms / ou.EPS,
            "tolerance": "max <= 32 eps of max|y|, rms <= 4 eps of rms y, counts equal"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU (default: the workload's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-scatter", action="store_true", help="skip the NCCL scatter/gather leg (N > 1)")
    args = ap.parse_args()

    src, dst, atten, tb, extfft, def_ch = WORKLOADS[args.workload]
    n_ch = args.channels or def_ch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)
    threads = host_threads()
    config = {"workload": args.workload, "src_rate": src, "dst_rate": dst, "preset": "CDSPResampler24",
              "trans_band_pct": tb, "channels_per_gpu": n_ch, "block_frames": BLOCK, "extfft": extfft,
              "parallelism": "channels sharded over %d GPU(s), no data-path collective" % world,
              "l2_policy": "per-step input (%.0f MB) and output exceed the 126 MB L2" % (n_ch * BLOCK * 8 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(src, dst, tb, atten, extfft, threads,
                              target_seconds=max(2.0, min(20.0, 0.5 * (K + W))))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (no /root/reference at build time)"}))
            return 0
        config["reference_sample_channels"] = r["sample_channels"]
        line = {"impl": "reference", "metric": "input Msamples/s (Mrops), fp64 %g->%g" % (src, dst),
                "value": r["value"], "unit": "Msamples/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": r["sample_ms_per_call"], "higher_is_better": True, "scaling": "weak",
                "ms_per_step_note": "one step of the reference arm = one process() call on each of the %d SAMPLE channels "
                                    "(config.channels_per_gpu is the GPU arm's workload); a full %d-channel step would take %.1f ms"
                                    % (r["sample_channels"], n_ch, r["sample_ms_per_call"] * n_ch / r["sample_channels"]),
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in CPU_KEYS if k in r},
                "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    from __graft_entry__ import load_package
    pkg = load_package()
    if not torch.cuda.is_available() or pkg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on stdout at first use; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    plan = pkg.Plan(src, dst, BLOCK, tb, atten, extfft=extfft)
    batch = pkg.Batch(plan, n_ch, local_rank)
    # this rank's thread and its pinned host buffers belong on the socket its GPU hangs off (two-socket boxes: GPUs 4-7 sit
    # on NUMA node 1; unbound, every rank's staging lands on node 0 and the end-to-end rate at N = 8 drops by a third)
    numa_node = batch.shards()[0][3]
    numa_bound = bind_to_numa_node(numa_node)
    cap = (plan.max_out_len + 7) // 8 * 8  # rows start 64-byte aligned
    # Two distinct input blocks alternate so that no step re-reads a block that could sit in L2.
    xs = [torch.from_numpy(synth_block(n_ch, BLOCK, 1000 + 17 * rank + i)).to(dev) for i in range(2)]
    out = torch.empty((n_ch, cap), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    batch.set_stream(stream.cuda_stream)

    call_log = []  # which input block every call since batch creation consumed (the oracle replays it)

    def step(i):
        call_log.append(i & 1)
        return batch.process_ptr(xs[i & 1].data_ptr(), BLOCK, BLOCK, out.data_ptr(), cap, cap)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # samples every 100 ms through warm-up and the timed region (both under load)
    for i in range(W):
        step(i)
    barrier()
    # keep the GPU under the same load until nvidia-smi has produced a few rows
    t_load = time.perf_counter()
    j = 0
    while rank == 0 and len(sampler.rows) < 3 and time.perf_counter() - t_load < 3.0:
        step(W + j)
        j += 1
        torch.cuda.synchronize(dev)
    sampler.rows = sampler.rows[-1:] if sampler.rows else []
    l0 = batch.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    n_out_total = 0
    step_counts = []
    for i in range(K):
        step_counts.append(step(W + i))
    n_out_total = sum(step_counts)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = batch.kernel_launches - l0
    # what the LAST timed call wrote, on a few sampled channels: checked against the oracle below
    check_ch = sorted(set([0, n_ch // 3, (2 * n_ch) // 3, n_ch - 1]))
    n_last = step_counts[-1] if step_counts else 0
    got_last = out[check_ch, :n_last].cpu().numpy()
    check_calls = list(call_log)
    check_x = [x[check_ch].cpu().numpy() for x in xs]
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = 1e-6 * world * n_ch * BLOCK * K / (ms * 1e-3)

    # ---- per-stage device time (separate pass; events around every stage launch)
    batch.set_timing(True)
    for i in range(K):
        step(i)
    st_times = batch.stage_times()
    batch.set_timing(False)
    stage_info = plan.stages()
    tot_stage_ms = sum(t[1] for t in st_times) or 1.0
    dom = max(range(len(st_times)), key=lambda i: st_times[i][1])
    # algorithmic bytes of the dominant kernel per launch: its own input + output streams
    rate = [1.0]
    for s in stage_info:
        prev = rate[-1]
        if s["name"] == "blockconv":
            rate.append(prev * s["up"] / s["down"])
        elif s["name"] == "hbup":
            rate.append(prev * 2)
        elif s["name"] == "hbdown":
            rate.append(prev / 2)
        else:
            rate.append(dst / src)  # interpolator lands on the chain's final ratio at that point
    # fix interpolator rates for chains where more stages follow (use exact plan ratios)
    for i, s in enumerate(stage_info):
        if s["name"].startswith("frac"):
            tail = 1.0
            for t in stage_info[i + 1:]:
                tail *= (t["up"] / t["down"]) if t["name"] == "blockconv" else (2 if t["name"] == "hbup" else 0.5 if t["name"] == "hbdown" else 1.0)
            rate[i + 1] = (dst / src) / tail
    kernels = batch.stage_kernels()  # (kernel name, plan stages covered) per plan stage
    span = max(1, kernels[dom][1])
    # a fused kernel reads the stream entering its first stage and writes the one leaving its last
    dom_bytes = 8.0 * n_ch * BLOCK * (rate[dom] + rate[dom + span])
    dom_ms = st_times[dom][1] / max(1, st_times[dom][2])
    peak, peak_src = measured_peaks()
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/), scaled to
    # this launch's in-samples; null when no capture of this kernel/workload is on file
    traffic, traffic_src = None, None
    try:
        for fn in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
            if fn.endswith(".json") and "ncu_summary" in fn:
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    ps = json.load(f)
                if kernels[dom][0] in ps.get("kernel", "") and ps.get("workload") == args.workload:
                    traffic = ps["dram_bytes_per_in_sample"] * n_ch * BLOCK
                    traffic_src = "profiles/" + fn
    except Exception:
        pass
    path_bytes_per_in = 8.0 * (1.0 + dst / src)
    roofline = {"bound": "hbm", "kernel": kernels[dom][0], "plan_stages_covered": span, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "kernel_share_of_step": st_times[dom][1] / tot_stage_ms,
                "stage_ms_per_step": {("%d:%s:%s" % (i, t[0], kernels[i][0])): t[1] / K for i, t in enumerate(st_times)},
                "traffic_source": traffic_src,
                "path": {"algorithmic_bytes_per_in_sample": path_bytes_per_in,
                         "achieved": path_bytes_per_in * value * 1e6 / world / 1e9,
                         "frac": path_bytes_per_in * value * 1e6 / world / 1e9 / peak}}

    # secondary ceiling: algorithmic flops of the path against the DFMA rate measured on this box
    flops_in = FLOPS_PER_IN_SAMPLE.get(args.workload)
    if flops_in is not None:
        try:
            pk = pkg.measure_fp64_tflops(local_rank)
            ach = flops_in * value * 1e6 / world / 1e12
            roofline["fp64"] = {"flops_per_in_sample": flops_in, "achieved_tflops": ach, "peak_tflops": pk,
                                "peak_source": "measured here (r8bgpu_measure_fp64_tflops: register-resident DFMA stream)",
                                "frac": ach / pk}
        except Exception as e:  # the calibration is informative; never fail the bench on it
            roofline["fp64"] = {"flops_per_in_sample": flops_in, "error": str(e)}

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers)
    e2e = None
    if not args.no_e2e:
        # pinned buffers placed on the GPU's NUMA node (r8bgpu_batch_host_alloc: mmap + mbind + cudaHostRegister)
        def pinned(cols):
            try:
                return torch.from_numpy(batch.host_alloc(cols)), "r8bgpu_batch_host_alloc (NUMA-placed, cudaHostRegister)"
            except Exception:  # mbind / registration refused on this host: ordinary pinned memory
                return torch.empty((n_ch, cols), dtype=torch.float64).pin_memory(), "torch pin_memory (fallback)"
        hx = []
        for i in range(2):
            t, host_mem = pinned(BLOCK)
            t.numpy()[:] = synth_block(n_ch, BLOCK, 2000 + i)
            hx.append(t)
        hy, host_mem = pinned(cap)
        batch.set_stream(None)
        ke = max(3, min(K, 8))
        for i in range(2):
            batch.process_host_ptr(hx[i & 1].data_ptr(), BLOCK, BLOCK, hy.data_ptr(), cap, cap)
        barrier()
        t0 = time.perf_counter()
        outs = 0
        for i in range(ke):
            outs += batch.process_host_ptr(hx[i & 1].data_ptr(), BLOCK, BLOCK, hy.data_ptr(), cap, cap)
        torch.cuda.synchronize(dev)
        t_e2e = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        e2e = {"value": 1e-6 * world * n_ch * BLOCK * ke / t_e2e, "unit": "Msamples/s",
               "h2d_bytes_per_step": n_ch * BLOCK * 8, "d2h_bytes_per_step": int(n_ch * (outs / ke) * 8),
               "steps": ke, "api": "r8bgpu_batch_process_host (pinned host in/out, sync per call)",
               "h2d_gbs_per_gpu": n_ch * BLOCK * 8 * ke / t_e2e / 1e9, "d2h_gbs_per_gpu": n_ch * (outs / ke) * 8 * ke / t_e2e / 1e9,
               "numa_node": numa_node, "numa_bound": numa_bound, "host_buffers": host_mem}

        if world == 1:
            # same call with float32 planar host buffers (r8bgpu_batch_process_host_fmt): what a caller holding
            # 32-bit audio pays -- half the PCIe bytes, widening/narrowing on the device (r8b_format.cu)
            fx = [torch.from_numpy(synth_block(n_ch, BLOCK, 3000 + i).astype("float32")).pin_memory() for i in range(2)]
            fy = torch.empty((n_ch, cap), dtype=torch.float32).pin_memory()
            bo = pkg.Buffer.make(fy.data_ptr(), pkg.F32, False, cap)
            bis = [pkg.Buffer.make(t.data_ptr(), pkg.F32, False, BLOCK) for t in fx]
            for i in range(2):
                batch.process_fmt(bis[i & 1], BLOCK, bo, cap, host=True)
            t0 = time.perf_counter()
            for i in range(ke):
                batch.process_fmt(bis[i & 1], BLOCK, bo, cap, host=True)
            torch.cuda.synchronize(dev)
            e2e["float32_io"] = {"value": 1e-6 * n_ch * BLOCK * ke / (time.perf_counter() - t0), "unit": "Msamples/s",
                                 "api": "r8bgpu_batch_process_host_fmt (R8BGPU_F32 planar in/out)"}

    # ---- NCCL scatter / gather of ONE buffer that lives on rank 0: timed apart from the kernels
    scatter = None
    if dist is not None and not args.no_scatter:
        par = __import__("r8brain_free_src_b200.parallel", fromlist=["x"])
        total_ch = n_ch * world
        full_in = torch.from_numpy(synth_block(total_ch, BLOCK, 4000)).to(dev) if rank == 0 else None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_sc, t_pr, t_ga = [], [], []
        for it in range(3):
            barrier()
            ev[0].record(stream)
            mine = par.scatter_channels(full_in, total_ch, BLOCK, dist, device=dev, dtype=torch.float64)
            ev[1].record(stream)
            n_sc = batch.process_ptr(mine.data_ptr(), BLOCK, BLOCK, out.data_ptr(), cap, cap)
            ev[2].record(stream)
            back = par.gather_channels(out[:, :n_sc].contiguous(), total_ch, dist)
            ev[3].record(stream)
            barrier()
            if it > 0:  # first round connects the NCCL peers
                t_sc.append(ev[0].elapsed_time(ev[1]))
                t_pr.append(ev[1].elapsed_time(ev[2]))
                t_ga.append(ev[2].elapsed_time(ev[3]))
            del back
        tt = torch.tensor([min(t_sc), min(t_pr), min(t_ga)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        scatter = {"scatter_ms": float(tt[0]), "process_ms": float(tt[1]), "gather_ms": float(tt[2]),
                   "scatter_bytes": (world - 1) * n_ch * BLOCK * 8, "gather_bytes": int((world - 1) * n_ch * n_sc * 8),
                   "scatter_gbs": (world - 1) * n_ch * BLOCK * 8 / (float(tt[0]) * 1e-3) / 1e9,
                   "gather_gbs": (world - 1) * n_ch * n_sc * 8 / (float(tt[2]) * 1e-3) / 1e9,
                   "how": "rank 0 holds [N*channels, frames] on its GPU; ncclSend/ncclRecv of contiguous row slabs "
                          "(parallel.scatter_channels / gather_channels); device time, max over ranks; NOT part of `value`"}
        del full_in

    verified = verify_against_oracle(src, dst, tb, atten, extfft, check_x, check_calls, got_last, check_ch)
    if dist is not None:
        t = torch.tensor([1.0 if verified["ok"] else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        verified["ok_all_ranks"] = bool(t.item() > 0.5)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_run(src, dst, tb, atten, extfft, threads, target_seconds=8.0)
        if r is not None:
            cpu = {k: r[k] for k in CPU_KEYS if k in r}

    if rank == 0:
        line = {"metric": "input Msamples/s (Mrops), fp64 %g->%g" % (src, dst), "value": value, "unit": "Msamples/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": (value / 38.0 if (src, dst) == (44100.0, 96000.0) else None),
                "vs_baseline_note": "published: 38 Mrops per core (Ooura FFT, Ryzen 3700X), README.md:111-114",
                "dtype": "f64", "data": "synthetic", "config": config,
                "out_msamples_per_s": 1e-6 * world * n_ch * n_out_total / (ms * 1e-3),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "verified": bool(verified["ok"] and verified.get("ok_all_ranks", True)), "verification": verified, "roofline": roofline, "cpu_baseline": cpu, "scatter_gather": scatter,
                "scatter_ms": scatter["scatter_ms"] if scatter else None, "gather_ms": scatter["gather_ms"] if scatter else None,
                "device_state_bytes": batch.device_bytes}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

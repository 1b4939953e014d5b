#!/usr/bin/env python3
"""bench.py — candidate IK rollouts/sec of the action-feasibility predict path.

  python bench.py --gpus N --steps K --warmup W            # CUDA arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle on host cores)

Workload (BASELINE.json configs[4]): synthetic sweep of
perturbed-goal `get` candidates in the reference's pizza scene, the three grasp
variants the kernel supports in equal parts (PowerGrasp tomato_sauce_bottle,
BallGrasp fresh_tomatoes, TopGrasp italian_seasoning_bowl), every rollout the
full 1555 steps (no early exit).  One "step" = one pass of the predict path over
one batch of ROLLOUTS_PER_GPU candidates per GPU.

  value : rollouts/s with the candidate tables already resident in HBM
          (kernel launches + NCCL verdict gather only), CUDA events, max over ranks.
          One launch per variant, each one resident wave of the thread-per-candidate kernel
          (148 SMs x 1024 rollouts), back to back on one stream.
  e2e   : the same metric through aff_predict_batch with HOST buffers in and out
          (plan compile, H2D of the tables, kernels, D2H of the verdicts), one host
          thread per variant running its variant for all --steps iterations back to back
          (the host work of a call overlaps the kernels of the calls before it)
  roofline: FP64 pipe (the path's bound); flops per rollout COUNTED by
          running the kernel source with a counting scalar (tests/tools/count_flops.py ->
          profiles/flops_per_rollout.json); peak = DFMA rate measured on this GPU in
          this run by aff_measure_fp64_peak.  roofline_hbm: the per-thread state of the
          sweep kernel streams through HBM; measured DRAM bytes per rollout from ncu
  cpu_baseline / --impl reference: the CPU oracle (a port: the reference cannot
          be built) on all host cores over a bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = ["get tomato_sauce_bottle", "get fresh_tomatoes", "get italian_seasoning_bowl"]
PER_VARIANT = 151552                     # one resident wave of the thread-per-candidate kernel: 148 SMs x 2 blocks x 512 rollouts
ROLLOUTS_PER_GPU = PER_VARIANT * len(VARIANTS)
STEPS_PER_ROLLOUT = 1555
WORKLOAD = ("synthetic sweep of perturbed-goal get candidates (BASELINE.json configs[4]), pizza scene, "
            "PowerGrasp/BallGrasp/TopGrasp in equal parts, 1555 steps per rollout, no early exit")
METRIC = "candidate IK rollouts/sec"
UNIT = "rollouts/s"
L2_FLUSH_BYTES = 256 << 20


# candidate tables read once (tasks, via points, activations, graph constraints, candidate record) + 88-byte verdict
ALGORITHMIC_BYTES_PER_ROLLOUT = 2300 + 88
# dram__bytes_read.sum + dram__bytes_write.sum of one aff_sweep_kernel launch / rollouts in it (profiles/r2_summary.md:
# 4.47 TB for 151 552 rollouts, gpurun_out/prof_r2c.ncu-rep): the per-thread rollout state does not fit on chip and streams through HBM
NCU_DRAM_BYTES_PER_ROLLOUT = 29480000


def algorithmic_flops_per_rollout():
    """FP64 flops of one full-length rollout (FMA = 2), mean over the three variants: COUNTED by executing
    the kernel source with a counting scalar (tests/emu/flop_count.cpp, tests/tools/count_flops.py), not estimated.
    -> (flops, per-variant dict, source)"""
    path = os.path.join(ROOT, "profiles", "flops_per_rollout.json")
    with open(path) as f:
        d = json.load(f)
    per = {k: v["flops_per_rollout"] for k, v in d["variants"].items()}
    return d["flops_per_rollout"], per, "profiles/flops_per_rollout.json (" + d["how"] + ")"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_batches(ab, scene, rank, per_variant, world=1):
    """This rank's contiguous block of every variant's candidate index space (affaction_b200.sharding:
    rank r of G takes [r N/G, (r+1) N/G)); candidate i of variant v has seed 0xAFFAC7 + v * 2^32 + i, so a
    smaller batch is a prefix of a larger one and the CPU sample compares like with like."""
    from affaction_b200.sharding import shard_bounds
    batches = []
    for v, cmd in enumerate(VARIANTS):
        lo, hi = shard_bounds(per_variant * world, rank, world)
        hb = ab.HostBatch(scene)
        hb.add_perturbed(cmd, hi - lo, seed0=0xAFFAC7 + (v << 32) + lo)
        batches.append(hb)
    return batches


def cpu_arm(ab, scene, n_rollouts, threads):
    """Oracle on `threads` host threads over the first n_rollouts candidates of the
    workload (equal parts of the three variants). Returns (rollouts/s, seconds)."""
    import oracle_binding as ob
    per = max(1, n_rollouts // len(VARIANTS))
    batches = make_batches(ab, scene, 0, per)
    opts = ab.default_opts()
    t0 = time.perf_counter()
    results = []
    for hb in batches:
        results.append(ob.predict_batch(scene, hb, opts, n_threads=threads))
    dt = time.perf_counter() - t0
    cpu_arm.last_results = results          # per variant, for the agreement report
    return per * len(VARIANTS) / dt, dt, per * len(VARIANTS)


def run_reference(args, rank, world):
    import affaction_b200 as ab
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    scene = ab.HostScene()
    sample = max(len(VARIANTS), 60 * threads)      # ~5-10 s per step
    for _ in range(args.warmup):
        cpu_arm(ab, scene, len(VARIANTS) * threads, threads)
    total, secs = 0, 0.0
    for _ in range(args.steps):
        _, dt, n = cpu_arm(ab, scene, sample, threads)
        total += n
        secs += dt
    value = total / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rollouts_per_step": sample, "steps_per_rollout": STEPS_PER_ROLLOUT, "note": "CPU oracle (restated reference; Rcs/Tropic are not vendored so the "
                   "reference itself cannot be built), ConcurrentExecutor-style pool on all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d rollouts per step x %d steps" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # candidates come from the pure C++ host shim (libaffhost.so); the CUDA library is never mapped by this arm
        "cuda_library_loaded": ab.gpu_library_loaded(),
    }
    print(json.dumps(line))


def run_cuda(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import affaction_b200 as ab
    from affaction_b200.sharding import gather_verdicts

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev_index = local_rank
    scene = ab.HostScene()
    per_variant = args.rollouts // len(VARIANTS) if args.rollouts else PER_VARIANT
    n_local = per_variant * len(VARIANTS)
    host_batches = make_batches(ab, scene, rank, per_variant, world)
    opts = ab.default_opts()
    stream = torch.cuda.current_stream().cuda_stream
    rec = C.sizeof(ab.Result)

    # ---- resident arm: candidate tables uploaded once, verdicts written into a torch buffer
    dev_batches = [ab.DeviceBatch(scene, hb, opts, device=dev_index) for hb in host_batches]
    kernels = sorted(set(db.kernel for db in dev_batches))
    verdicts = torch.empty(n_local * rec, dtype=torch.uint8, device="cuda")
    gathered = torch.empty(world * n_local * rec, dtype=torch.uint8, device="cuda") if (world > 1 and rank == 0) else None
    for i, db in enumerate(dev_batches):
        db.set_results_buffer(verdicts.data_ptr() + i * per_variant * rec)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def launch_all():
        """one full-wave launch per variant, back to back"""
        n_launch = 0
        for db in dev_batches:
            n_launch += db.launch(stream)
        return n_launch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        launch_all()
        gather_verdicts(verdicts, world, dist, rank, gathered)
    barrier()

    sampler = ClockSampler(dev_index)
    sampler.start()
    kernel_ms, step_ms, launches = [], [], 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(args.steps):
        flush.fill_(1)                  # L2 flush between timed iterations
        barrier()
        ev[0].record()
        launches += launch_all()
        ev[1].record()
        gather_verdicts(verdicts, world, dist, rank, gathered)      # per-candidate verdicts and costs to rank 0
        ev[2].record()
        barrier()
        kernel_ms.append(ev[0].elapsed_time(ev[1]))
        step_ms.append(ev[0].elapsed_time(ev[2]))
    total_ms = sum(step_ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n_local * args.steps / (total_ms * 1e-3)

    # verdict mix of this rank's last step
    res = (ab.Result * n_local).from_buffer_copy(verdicts.cpu().numpy().tobytes())
    mix = {}
    for r in res:
        mix[ab.CODE_NAMES.get(r.code, str(r.code))] = mix.get(ab.CODE_NAMES.get(r.code, str(r.code)), 0) + 1

    # ---- e2e arm: host buffers in, host verdicts out, through aff_predict_batch; one host thread per variant
    # (the call is blocking and thread-safe; its kernels run on private streams and share the GPU)
    e2e_out = [None] * len(VARIANTS)

    def e2e_call(i):
        e2e_out[i] = ab.predict_batch(scene, host_batches[i], opts, device=dev_index)

    def e2e_step():
        th = [threading.Thread(target=e2e_call, args=(i,)) for i in range(len(VARIANTS))]
        for t in th:
            t.start()
        for t in th:
            t.join()

    def e2e_stream(i):
        # one variant, all steps back to back: the host work of a step (plan compile, upload) overlaps the
        # kernels of the calls before it, as a sweep service that keeps the GPU fed would run it.  Every
        # step still copies its inputs from host memory and reads its verdicts back, all inside the timed region.
        for _ in range(args.steps):
            e2e_call(i)

    e2e_step()   # warm-up
    barrier()
    t0 = time.perf_counter()
    th = [threading.Thread(target=e2e_stream, args=(i,)) for i in range(len(VARIANTS))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n_local * args.steps / e2e_s
    e2e_equal = all(a.as_tuple() == b.as_tuple() for i in range(len(VARIANTS))
                    for a, b in zip(e2e_out[i], res[i * per_variant:(i + 1) * per_variant]))
    h2d = sum(db.h2d_bytes for db in dev_batches)
    d2h = n_local * rec
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        # roofline of the dominant kernel, live CUDA-event timing of its launches
        peak = ab.lib().aff_measure_fp64_peak(dev_index, 0.5)
        flops_per, flops_by_variant, flops_src = algorithmic_flops_per_rollout()
        flops = flops_per * n_local
        avg_kernel_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
        achieved = flops / avg_kernel_s * 1e-12
        hbm_peak, hbm_src = 6553.9, "fallback (MEASURED_PEAKS.json missing)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except (OSError, KeyError, ValueError):
            pass
        sweep = kernels == ["sweep"]
        dram_per_rollout = NCU_DRAM_BYTES_PER_ROLLOUT if sweep else 2758
        hbm_achieved = dram_per_rollout * n_local / avg_kernel_s * 1e-9
        # bounded CPU sample on the box's host cores (N=1 only)
        cpu, agreement = None, None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            v, secs, n = cpu_arm(ab, scene, max(len(VARIANTS), 120 * threads), threads)   # ~10-30 s of CPU work
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d rollouts of the same workload in %.1f s (CPU oracle, %d threads)" % (n, secs, threads)}
            # the CPU sample = the first candidates of each variant of this rank's GPU batch: compare the records
            agreement = {"sample": 0, "records_equal": 0, "verdicts_equal": 0, "ranked_first_equal": True,
                         "max_abs_diff": {"min_dist": 0.0, "jl_cost": 0.0, "coll_cost": 0.0}}
            for vi, cpu_res in enumerate(cpu_arm.last_results):
                gpu_res = res[vi * per_variant: vi * per_variant + len(cpu_res)]
                for a, b in zip(cpu_res, gpu_res):
                    agreement["sample"] += 1
                    agreement["records_equal"] += int(a.as_tuple() == b.as_tuple())
                    agreement["verdicts_equal"] += int((a.success, a.code, a.first_fail_step) == (b.success, b.code, b.first_fail_step))
                    for f in agreement["max_abs_diff"]:
                        agreement["max_abs_diff"][f] = max(agreement["max_abs_diff"][f], abs(getattr(a, f) - getattr(b, f)))
                if ab.rank(cpu_res)[0] != ab.rank(list(gpu_res))[0]:
                    agreement["ranked_first_equal"] = False
        kname = "aff_sweep_kernel" if sweep else "aff_rollout_kernel"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rollouts_per_gpu_per_step": n_local, "steps_per_rollout": STEPS_PER_ROLLOUT,
                       "kernel": kname + (" (one rollout per thread, 32 per warp)" if sweep else " (one rollout per warp)"),
                       "l2": "flushed between timed iterations (256 MiB fill)",
                       "parallelism": "candidates sharded statically over %d GPU(s); NCCL gather of verdicts to rank 0" % world,
                       "verdict_mix_rank0": mix},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.steps, "steps_overlap": True, "records_equal_resident_arm": bool(e2e_equal),
                    "call": "aff_predict_batch (plan compile + H2D + kernels + D2H, pageable host buffers), one host thread per variant"},
            "gpu_launches": launches * world,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak > 0 else None,
                         "traffic": dram_per_rollout * per_variant, "traffic_unit": "bytes per launch (dram read + write of one "
                         "launch of %d rollouts, ncu --set full capture, profiles/r2_summary.md)" % per_variant,
                         "kernel": kname, "flops_per_rollout": flops_per, "flops_per_rollout_by_variant": flops_by_variant,
                         "flops_source": flops_src,
                         "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json holds no FP64 figure)",
                         "kernel_ms_per_step": sum(kernel_ms) / len(kernel_ms)},
            # the same kernel against the HBM roof: MEASURED dram traffic (the rollout state streams through HBM), not algorithmic bytes
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_achieved / hbm_peak, "measured_dram_bytes_per_rollout": dram_per_rollout,
                             "algorithmic_bytes_per_rollout": ALGORITHMIC_BYTES_PER_ROLLOUT, "peak_source": hbm_src},
            "cpu_baseline": cpu,
            "agreement": agreement,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--rollouts", type=int, default=0, help="rollouts per GPU per step (default %d)" % ROLLOUTS_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import __graft_entry__ as ge
    if rank == 0 or not os.path.isfile(os.path.join(ROOT, "affaction_b200", "libaffpredict.so")):
        ge.build(check_gpu_library=args.impl != "reference")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

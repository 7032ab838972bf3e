#!/usr/bin/env python
"""Benchmark of the rover hot path (BASELINE.json metric: env-steps/s and heightmap rays/s).

  python bench.py --gpus N --steps K --warmup W            # this repo (hand-written sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's torch-CPU algorithm (oracle port)

Workload (BASELINE.json configs[1]): 4,096 envs per GPU, full dense+sparse heightmap (1634 rays/env, K=200
candidates/ray) on a synthetic 200 x 200 m, ~1M-triangle terrain (708^2 heightfield), big_rock_layer
collision, Ackermann kinematics, reward/reset terms.  One "step" = one pass of the hot path over all envs
(RoverTask.hot_step), PhysX excluded.  Multi-GPU: envs sharded by index (weak scaling), terrain replicated,
one NCCL all-reduce of the 16-entry statistics vector per step.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_RAYS = 1634
ALGO_BYTES_PER_ENV_STEP = 665196          # index rows 785*K*4 + unique triangles + inputs + outputs
PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")
FALLBACK_HBM_GBS = 6650.0                 # /opt/skills/guides/B200_PROFILING.md fallback


_JSON_FD = None


def claim_stdout():
    """Everything that libraries print to stdout while the benchmark runs (NCCL's "NCCL version ..." banner, for one) goes to
    stderr; the ONE JSON line is written to the real stdout by emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def hbm_peak():
    try:
        return float(json.load(open(PEAKS_FILE))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled every 10 ms during the timed regions: NVML in a background thread of this
    process (what `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` reads, without spawning a poller process
    that contends with the benchmark for the driver); falls back to an `nvidia-smi -lms 50` child if NVML is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.thread, self.sm, self.reasons, self.mx = None, None, [], set(), None
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                uuid = torch.cuda.get_device_properties(index).uuid          # CUDA_VISIBLE_DEVICES-proof
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            self.stop_flag = threading.Event()

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for name, b in bits.items():
                            if r & b:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self.stop_flag.wait(0.01)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "source": "NVML, 10 ms period, in-process thread"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 50"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_sample_assets(n_envs, seed=42):
    """A 24 m crop-sized terrain with the SAME mesh density (0.2825 m vertex spacing), K=200, res 0.1 m as the
    200 m benchmark terrain; index by CPU brute force (rover_utils.py:52-118 semantics).  Per-env CPU cost does
    not depend on the map extent, so env-steps/s measured here is the CPU figure for the full workload."""
    import isaac_rover_b200  # noqa: F401  (synthetic generator only; no kernels on this path)
    from isaac_rover_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rover_oracle as O
    w = synth.make_world(length=24.0, nv=86, K=200, n_stones=30, seed=seed, build_index="cpu")
    pat, ci, fi = O.heightmap_pattern()
    assets = dict(pattern=pat, coarse_idx=ci, fine_idx=fi, map_indices=w.map_indices, triangles=w.triangles,
                  vertices=w.vertices, rock_indices=w.rock_indices, rock_triangles=w.rock_triangles,
                  rock_vertices=w.rock_vertices, shift=torch.tensor([0, 0, 0.0]))
    st = synth.make_env_state(w, n_envs, seed=seed, margin=6.0)
    return O, assets, st


def time_cpu_oracle(n_envs, steps, warmup):
    torch.set_num_threads(os.cpu_count())
    O, assets, st = cpu_sample_assets(n_envs)
    for _ in range(warmup):
        O.full_step(assets, st)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.full_step(assets, st)
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)


def time_gpu_eager_oracle(n_envs, dev, steps=3, warmup=1):
    """The reference's eager-torch op sequence (oracle port) on the GPU itself -- how the reference is actually deployed
    (it hard-wires 'cuda:0', rover.py:90).  Reported beside the CPU baseline; same 24 m sample world, n_envs envs."""
    O, assets, st = cpu_sample_assets(n_envs)
    assets = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in assets.items()}
    st = {k: v.to(dev) for k, v in st.items()}
    for _ in range(warmup):
        O.full_step(assets, st, env_chunk=min(n_envs, 128))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        O.full_step(assets, st, env_chunk=min(n_envs, 128))
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def time_policy_epilogue(R, obs, steps=20, warmup=3):
    """SURVEY.md 8f-3, reported beside the step (not part of `value`): actor + critic inference on the step's obs_buf through
    rvb_policy_forward, and the same two networks as eager torch fp32 (cuBLAS) on this GPU for comparison."""
    import torch.nn.functional as F
    N = obs.shape[0]
    net = R.model.NetworkInfo([256, 160, 128], [80, 60], [80, 60], [80, 60], "leakyrelu")
    info = R.model.ObserverationInfo(4, 634, 1112, 0)
    actor = R.model.StochasticActorHeightmap(1750, 2, net, info, device=obs.device)
    critic = R.model.DeterministicHeightmap(1750, 2, net, info, device=obs.device)
    out_a = torch.empty((N, 2), dtype=torch.float32, device=obs.device)
    out_c = torch.empty((N, 1), dtype=torch.float32, device=obs.device)

    def ours():
        R.model.compute_pair(actor, critic, obs, out_a, out_c)

    def ours_single():
        actor._forward(obs, out_a)

    def eager_net(sd, tanh):
        sd = {k: v.to(obs.device) for k, v in sd.items()}

        def chain(x, prefix, n):
            for i in range(n):
                x = F.leaky_relu(F.linear(x, sd["%s.%d.layer.0.weight" % (prefix, i)], sd["%s.%d.layer.0.bias" % (prefix, i)]))
            return x

        def fwd():
            x = torch.cat((obs[:, 0:4], chain(obs[:, 4:638], "encoder0.encoder", 2), chain(obs[:, 638:1750], "encoder1.encoder", 2)), dim=1)
            x = F.linear(chain(x, "network", 3), sd["network.3.weight"], sd["network.3.bias"])
            return torch.tanh(x) if tanh else x
        return fwd
    ea, ec = eager_net(actor.state_dict(), True), eager_net(critic.state_dict(), False)

    def timed(fn):
        for _ in range(warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    ms = timed(ours)
    ms_single = timed(ours_single)
    ms_eager = timed(lambda: (ea(), ec()))
    err = max((out_a - ea()).abs().max().item(), (out_c - ec()).abs().max().item())
    flops = 2 * 2 * N * (634 * 80 + 1112 * 80 + 2 * 80 * 60 + 124 * 256 + 256 * 160 + 160 * 128 + 128 * 2)
    return {"ms": ms, "envs_per_s": N / ms * 1e3, "fp32_tflops": flops / ms / 1e9, "torch_eager_fp32_ms": ms_eager,
            "max_abs_diff_vs_torch": err, "launches": 1, "actor_alone_ms": ms_single,
            "what": "actor + critic (encoders [80,60] x2, mlp [256,160,128], model.py:152-241) on the step's obs_buf f32 [%d,1750]; "
                    "one fused fp32 launch for both networks (rvb_policy_forward_pair); not included in `value`" % N}


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.cpu_envs
    t = time_cpu_oracle(n, args.steps, args.warmup)
    v = n / t
    cores = torch.get_num_threads()
    sample = ("%d envs x %d rays x K=200 per step on a 24 m terrain of the benchmark's mesh density; the reference is Python "
              "(cannot travel to the GPU box), so its torch-CPU algorithm is run through the oracle restatement "
              "(oracle/rover_oracle.py, pinned bit-exact to the reference)" % (n, P_RAYS))
    line = {"impl": "reference", "metric": "env-steps/sec (obs+kinematics+reward)", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "4096 envs/GPU, 1634 rays/env, K=200, 200x200 m ~1M-triangle synthetic terrain (configs[1])",
                       "sample_envs": n},
            "rays_per_s": v * P_RAYS,
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local):
    import isaac_rover_b200 as R
    from isaac_rover_b200 import synth
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device -- this arm has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.envs
    w = synth.make_world(length=args.length, nv=args.nv, K=args.K, n_stones=args.stones, seed=42, build_index=None)
    t0 = time.perf_counter()
    w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device=dev)
    w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, w.K, device=dev)
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0
    n_sets = 3
    states = [synth.make_env_state(w, N, seed=100 + s, env_offset=rank * N) for s in range(n_sets)]
    task = synth.make_task(w, states[0], device=str(dev), level=2, num_envs_total=N * world)
    w.map_indices = w.rock_indices = None          # the layers own K-contiguous copies
    dstates = [{k: v.to(dev) for k, v in s.items()} for s in states]
    view = task._rover

    def set_state(i):
        s = dstates[i % n_sets]
        view.pos, view.quat, view.joints = s["pos"], s["quat"], s["joints"]
        return s["actions"]

    reducer = R.dist.StatsReducer()

    def step(i):
        task.hot_step(set_state(i), fused=not args.unfused)
        if args.sync_reduce:
            R.dist.reduce_stats(task.stats)
        else:
            reducer.submit(task.stats)          # asynchronous all-reduce of the 16 sums (off the critical path)

    task.Camera.variant = args.variant
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    R.dist.barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    task.Camera.timing = []
    lib = R._lib.load()
    lib.rvb_timing_enable(1)
    launches0 = R._lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    last = None
    for i in range(args.steps):
        last = step(args.warmup + i)
    if not args.sync_reduce:
        reducer.result((reducer.i - 1) % reducer.depth)       # the timed region ends when the last reduction has landed
    e1.record()
    torch.cuda.synchronize()
    R.dist.barrier()
    launches = R._lib.launch_count - launches0
    dt = R.dist.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    ray_ms = [a.elapsed_time(b) for a, b in task.Camera.timing]
    task.Camera.timing = None
    if not ray_ms:          # fused step: the library timed the ray-cast itself (CUDA events on the launching stream)
        import ctypes
        buf = (ctypes.c_float * (args.steps + 8))()
        n_t = lib.rvb_timing_read(buf, args.steps + 8)
        ray_ms = [buf[i] for i in range(max(n_t, 0))]
    lib.rvb_timing_enable(0)
    ray_s = R.dist.max_over_ranks(sum(ray_ms) / len(ray_ms) * 1e-3, dev)
    # ---- end to end through the host-buffer API (H2D + hot path + D2H every step)
    hstates = [{k: v.pin_memory() for k, v in s.items() if k in ("pos", "quat", "joints", "actions")} for s in states]

    def run_e2e(packed):
        """every step: inputs copied from pinned host memory, results (obs, rew, reset) read back and touched on the host;
        two slots, so the read-back of step i overlaps the kernels of step i+1"""
        pipe = R.HostPipeline(task, packed_obs=packed)
        for i in range(max(args.warmup, 3)):
            h = hstates[i % n_sets]
            pipe.step(h["pos"], h["quat"], h["joints"], h["actions"])
        torch.cuda.synchronize()
        R.dist.barrier()
        t0 = time.perf_counter()
        prev, checksum = None, 0.0

        def touch(res):
            rew_h, reset_h = res[-2], res[-1]
            return float(rew_h[0]) + float(res[-3][-1, -1]) + int(reset_h[0])
        for i in range(args.steps):
            h = hstates[i % n_sets]
            k = pipe.submit(h["pos"], h["quat"], h["joints"], h["actions"])
            if args.sync_reduce:
                R.dist.reduce_stats(task.stats)
            else:
                reducer.submit(task.stats)
            if prev is not None:
                checksum += touch(pipe.result(prev))
            prev = k
        checksum += touch(pipe.result(prev))
        if not args.sync_reduce:
            reducer.result((reducer.i - 1) % reducer.depth)
        torch.cuda.synchronize()
        dt_ = R.dist.max_over_ranks(time.perf_counter() - t0, dev)
        task.obs16_buf = None
        return dt_, pipe.h2d_bytes, pipe.d2h_bytes, checksum
    dt_e2e, h2d_b, d2h_b, checksum = run_e2e(False)
    dt_e2e_p, h2d_bp, d2h_bp, checksum_p = run_e2e(True)
    clk = clocks.stop() if clocks else None
    if rank != 0:
        return
    total_envs = N * world
    value = total_envs * args.steps / dt
    peak, peak_src = hbm_peak()
    achieved = ALGO_BYTES_PER_ENV_STEP * N / ray_s / 1e9
    traffic, issue = None, None
    tf = os.path.join(ROOT, "profiles", "raycast_traffic.json")
    if os.path.exists(tf):
        try:
            prof = json.load(open(tf))
            traffic = prof.get("dram_bytes_per_launch")
            if N != prof.get("envs_per_launch"):
                traffic = None                   # the ncu capture is of the 4096-env launch
            # the resource that actually binds the kernel: warp-instruction issue slots.  Instruction count per
            # env from the ncu capture (smsp__inst_executed.sum), time measured live, clock sampled live
            wi = prof["warp_instructions_per_launch"] / prof["envs_per_launch"] * N
            clk_hz = float(clk["sm_mhz"]) * 1e6 if (clk and clk.get("sm_mhz")) else 1.965e9      # median SM clock under load
            issue = {"warp_inst_per_launch": wi, "source": prof.get("source"), "sm_clock_hz": clk_hz,
                     "peak_warp_inst_per_s": 148 * 4 * clk_hz, "achieved_warp_inst_per_s": wi / ray_s,
                     "frac": wi / ray_s / (148 * 4 * clk_hz)}
        except Exception:
            traffic, issue = None, None
    line = {"metric": "env-steps/sec (obs+kinematics+reward)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "%d envs/GPU, 1634 rays/env, K=%d, %gx%g m %d-triangle synthetic terrain + big_rock_layer "
                                   "(%d triangles), %d stones (configs[1])" % (N, w.K, w.length, w.length, w.triangles.shape[0],
                                                                               w.rock_triangles.shape[0], w.stone_info.shape[0]),
                       "envs_per_gpu": N, "total_envs": total_envs, "rays_per_env": P_RAYS, "K": w.K, "index_cells": w.G * w.G,
                       "l2": "inputs larger than L2: %d pose sets cycled, %.1f GB of index rows touched per step, index %.1f GB"
                             % (n_sets, N * 785 * w.K * 4 / 1e9, w.G * w.G * w.K * 4 / 1e9),
                       "parallelism": "env shards x%d, terrain replicated, 1 all-reduce of 16 f64 per step (%s)"
                                      % (world, "compute stream" if args.sync_reduce else "asynchronous, NCCL stream"),
                       "raycast_variant": args.variant, "fused_step": not args.unfused, "index_build_s": round(t_index, 3)},
            "rays_per_s": value * P_RAYS,
            # (ray, candidate) tests the reference evaluates for the same output: (1634 heightmap + 26 rock rays) x K per env-step
            "reference_equivalent_pair_tests_per_s": value * (P_RAYS + 26) * w.K,
            "raycast_ms": ray_s * 1e3,
            "raycast_share_of_step": ray_s / (dt / args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * N,
                         "kernel": "heightmap ray-cast (Camera.get_depths)",
                         "note": "contractual HBM figure on the reference-format algorithmic bytes; the kernel is bound by "
                                 "instruction issue, see issue_slots", "issue_slots": issue},
            # headline end-to-end number: the host pipeline with the observation read back in its native precision (the
            # heightmap columns ARE fp16 values, rover.py:324-325; HostPipeline.obs_f32() widens them on the host);
            # e2e_f32_obs is the same loop reading back the reference's f32 [N,1750] layout (twice the bytes: at 8 GPUs the
            # host link, not the GPUs, then sets the pace)
            "e2e": {"value": total_envs * args.steps / dt_e2e_p, "unit": "env-steps/s", "h2d_bytes_per_step": h2d_bp * world,
                    "d2h_bytes_per_step": d2h_bp * world, "ms_per_step": dt_e2e_p / args.steps * 1e3,
                    "api": "HostPipeline(packed_obs=True).submit/result (2 slots: read-back of step i overlaps step i+1; inputs "
                           "uploaded from pinned host memory every step; obs = f32 [N,4] proprioceptive + f16 [N,1746] heightmap "
                           "columns, lossless; rew f32 [N]; reset i64 [N])", "checksum": checksum_p},
            "e2e_f32_obs": {"value": total_envs * args.steps / dt_e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d_b * world,
                            "d2h_bytes_per_step": d2h_b * world, "ms_per_step": dt_e2e / args.steps * 1e3,
                            "api": "HostPipeline.submit/result, obs read back as the reference's f32 [N,1750]", "checksum": checksum},
            "gpu_launches": launches,
            "clocks": clk}
    if world == 1 and not args.no_cpu:
        n = args.cpu_envs
        t = time_cpu_oracle(n, 2, 1)
        line["cpu_baseline"] = {"value": n / t, "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d envs (x1634 rays x K=200) per step, 1 warm-up + 2 timed steps of the oracle port of the "
                                          "reference's torch path on a 24 m terrain of the same mesh density" % n}
        try:
            ng = 256
            tg = time_gpu_eager_oracle(ng, dev)
            line["torch_eager_gpu_baseline"] = {"value": ng / tg, "unit": "env-steps/s", "kind": "port",
                                                "sample": "%d envs per step, 1 warm-up + 3 timed steps of the same oracle port run as eager "
                                                          "torch ops on this GPU (the reference's own deployment: ~700 ATen launches per "
                                                          "step, fp16 temporaries in HBM)" % ng}
        except Exception as e:          # a baseline, never a reason to lose the bench line
            line["torch_eager_gpu_baseline"] = {"unavailable": str(e)[:200]}
    if world == 1:
        try:
            line["policy_epilogue"] = time_policy_epilogue(R, task.obs_buf)
        except Exception as e:          # reported beside the step, never a reason to lose the bench line
            line["policy_epilogue"] = {"unavailable": str(e)[:200]}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--length", type=float, default=200.0)
    ap.add_argument("--nv", type=int, default=708)
    ap.add_argument("--K", type=int, default=200)
    ap.add_argument("--stones", type=int, default=2000)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cpu-envs", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="one library call per reference call instead of rvb_env_step")
    ap.add_argument("--sync-reduce", action="store_true", help="all-reduce the statistics on the compute stream every step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 1 if args.impl == "reference" else 3)
    rank = int(os.environ.get("RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import isaac_rover_b200 as R
    rank, world, local = R.dist.init_from_env()
    try:
        run_b200(args, rank, world, local)
    finally:
        if world > 1 and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

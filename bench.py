#!/usr/bin/env python
"""bench.py -- agent-steps/sec of the gym-formation MPE step path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU reference arm

One "step" = one env step (MultiAgentEnv.step semantics) of EVERY env of the batch: the random
policy kernel (fg_random_actions: act ~ U(-1,1), test.py:20) followed by the fused step kernel
(fg_step_fused: _set_action + World.step + observation + reward + done + auto-reset).  Workload
(config.workload): formation_hd_env, 9 agents, 131072 envs per GPU (= the north star's 1M envs on
8 GPUs), episode_length 25, fp32.  A step moves 131072 * 2437 B = 319 MB > the 126 MB L2, so
every timed iteration streams from/to HBM (no L2 flush needed).

Printed JSON keys beyond the base contract: `roofline` (fused step kernel vs measured HBM peak),
`cpu_baseline` (oracle port on the host cores, N=1 only), `e2e` (same metric through the public
API with pinned HOST buffers: H2D of the actions and D2H of obs/reward/done inside the timed
region), `clocks`, `gpu_launches`, `also` (other BASELINE configs, informational).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "gym-formation_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "agent-steps/sec (formation_hd_env, random policy)"
UNIT = "agent-steps/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=25)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenario", default="formation_hd_env")
    ap.add_argument("--agents", type=int, default=9)
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--episode-length", type=int, default=25)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-obs", action="store_true", help="state+reward only (B_state accounting)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--graph-steps", type=int, default=50,
                    help="env steps per CUDA graph in the timed region (0 = plain per-step launches)")
    ap.add_argument("--cpu-seconds", type=float, default=3.0)
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons)}


def run_reference(a):
    """CPU reference arm: the reference's per-env numpy loop (oracle/ref_loop_port.py -- the
    unmodified reference cannot travel to the GPU box) on all host cores, one env per process
    like the reference's SubprocVecEnv.  A 'step' = one env step of that batch of P envs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    from oracle import ref_loop_port as rp

    procs = os.cpu_count() or 1
    N, K, W = a.agents, a.steps, a.warmup
    # keep the whole run within a few minutes whatever K is
    per_step = {3: 0.004, 9: 0.015, 27: 0.07}.get(N, 0.01 * N)
    K_eff = max(1, min(K, int(120.0 / per_step)))

    def worker(q, seed):
        np.random.seed(seed)
        env = rp.RefLoopEnv(a.scenario, N, a.episode_length)
        acts = lambda: [np.random.uniform(-1, 1, 2) for _ in range(N)]  # noqa: E731
        for _ in range(min(W, 10)):
            env.step(acts())
        t0 = time.perf_counter()
        for _ in range(K_eff):
            _, _, done_n, _ = env.step(acts())
            if all(done_n):
                env.reset()
        q.put(time.perf_counter() - t0)

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(q, 100 + k)) for k in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    times = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = max(times)
    value = procs * N * K_eff / wall
    sample = ("%d envs (one per host core, %d processes) x %d env-steps of %s N=%d, episode_length %d; "
              "per-env Python/numpy loop port of the reference (oracle/ref_loop_port.py, scipy "
              "directed_hausdorff=%s)" % (procs, procs, K_eff, a.scenario, N, a.episode_length, rp.HAVE_SCIPY))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": K_eff, "warmup": min(W, 10), "ms_per_step": wall / K_eff * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a, procs, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(a, envs_per_unit, where):
    return {"workload": "%s N=%d, %d envs per %s, episode_length %d, random policy U(-1,1), auto-reset"
                        % (a.scenario, a.agents, envs_per_unit, "GPU" if where == "gpu" else "run",
                           a.episode_length),
            "scenario": a.scenario, "agents": a.agents, "envs_per_gpu": envs_per_unit if where == "gpu" else None,
            "episode_length": a.episode_length, "obs": not a.no_obs,
            "launch": ("cuda graph of per-step launches" if a.graph_steps > 0 else "per-step launches") if where == "gpu" else None,
            "l2": "inputs larger than L2 (step traffic > 126 MB), no flush" if where == "gpu" else None}


def run_b200(a):
    import torch
    import torch.distributed as dist
    import formation_gym
    from formation_gym import distributed as fgd

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the step path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner / debug lines to stdout: send them to a file so that stdout carries
        # exactly ONE JSON line whatever NCCL_DEBUG level the launcher chose
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fg_bench_nccl.%h.%p.log")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL honours the debug file only above the VERSION level
        dist.init_process_group("nccl", device_id=device)
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    E, N, K, W = a.envs_per_gpu, a.agents, a.steps, max(a.warmup, 3)
    lo, hi = fgd.shard_range(E * world, rank, world)              # contiguous env range of this rank
    env = formation_gym.make_batched_env(a.scenario, hi - lo, N, a.episode_length, device=device,
                                         dtype=dtype, seed=0, auto_reset=True, env_offset=lo,
                                         write_obs=not a.no_obs)
    env.reset()

    def barrier():
        if world > 1:
            dist.barrier()

    def one_step():
        env.sample_actions()
        return env.step(env.actions)

    for _ in range(W):
        one_step()
    torch.cuda.synchronize()

    # ---- timed region 1 (-> value): EXACTLY K steps, replayed from a CUDA graph of per-step launches
    # (random-policy kernel + fused step kernel per step; the Philox tick lives on the device so every
    # replayed step draws fresh numbers).  Host cost per step ~0, so the number is the GPU's.
    chunk = max(1, min(K, a.graph_steps))
    while K % chunk:
        chunk -= 1
    graph = env.capture_steps(chunk) if a.graph_steps > 0 else None
    if graph is not None:
        graph.replay()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = env.launches
    barrier(); torch.cuda.synchronize()
    ev0.record()
    if graph is not None:
        for _ in range(K // chunk):
            graph.replay()
    else:
        for _ in range(K):
            one_step()
    ev1.record()
    torch.cuda.synchronize(); barrier()
    launches = 2 * K                                              # fg::k_random_actions + step kernel per step
    ms = ev0.elapsed_time(ev1)

    # ---- timed region 2 (-> roofline): the fused step kernel ALONE.  G steps are captured into a CUDA graph, each
    # on its own pre-sampled action buffer (fresh actions every step, as in region 1, but no random-policy kernel in
    # between), and the graph is replayed until K steps have run; CUDA events on the launching stream bracket the
    # replays.  kernel_ms = elapsed / steps is the kernel's average launch duration including the (sub-microsecond)
    # kernel-to-kernel hand-over inside the graph.  (Events recorded around every single launch, the first version
    # of this region, added ~5 us of launch latency to each 53 us kernel.)
    G = max(1, min(K, 20))
    acts = []
    for _ in range(G):
        env.sample_actions()
        acts.append(env.actions.clone())
        env.step(env.actions)                                     # also advances the state / tick like region 1
    env.use_device_tick(True)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for a_k in acts:                                          # warm-up outside capture (per-buffer fg_buffers)
            env.step(a_k)
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for a_k in acts:
            env.step(a_k)
    g2.replay(); torch.cuda.synchronize()
    reps2 = max(1, (K + G - 1) // G)
    er0, er1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    er0.record()
    for _ in range(reps2):
        g2.replay()
    er1.record()
    torch.cuda.synchronize(); barrier()
    region2_ms = er0.elapsed_time(er1)
    kernel_ms = region2_ms / (reps2 * G)
    del g2, acts
    # keep the sampler alive a little if the regions were short, so it has samples under load
    clocks = None
    if sampler:
        t_extra = time.time()
        while len(sampler.rows) < 3 and time.time() - t_extra < 1.0:
            one_step()
        torch.cuda.synchronize()
        clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_agents = (hi - lo) * world * N
    value = total_agents * K / (ms * 1e-3)

    # episode statistics: the ONLY collective of the design (NCCL all-reduce of 4 doubles)
    stats = fgd.all_reduce_stats(env.stats.clone())
    ep = {"episodes": float(stats[0]), "return_mean": float(stats[1] / stats[0]) if float(stats[0]) else None}

    # ---------------- e2e: the reference-facing VecEnv API with HOST (numpy) buffers ----------------
    # formation_gym.make_vec_env(...).step(actions_np) -> (obs_np, rews_np, dones_np, infos): what the
    # reference's trainers call on SubprocVecEnv (train/maddpg-v2/utils/env_wrappers.py:63-72).  Per step:
    # H2D of the actions from pinned host memory, the fused step kernel, D2H of obs / rewards / dones /
    # individual rewards into pinned host memory, host sync -- all inside the timed region.
    e2e = None
    if env.obs is not None:
        Ke = max(1, a.e2e_steps)
        del graph
        venv = formation_gym.make_vec_env(a.scenario, hi - lo, N, a.episode_length, device=device, dtype=dtype,
                                          seed=0, env_offset=lo, to_numpy=True)
        venv.reset()
        act_h = venv.action_buffer                                # pinned host array [E,N,2]
        act_h[...] = torch.empty(act_h.shape, dtype=dtype).uniform_(-1, 1).numpy()
        out_bytes = 0
        for _ in range(2):
            outs = venv.step(act_h)
        out_bytes = outs[0].nbytes + outs[1].nbytes + outs[2].nbytes + outs[3].individual_reward.nbytes
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(Ke):
            venv.step(act_h)                                      # returns host arrays (synchronises)
        dt_e = time.perf_counter() - t0
        barrier()
        te = torch.tensor([dt_e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_agents * Ke / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(act_h.nbytes) * world,
               "d2h_bytes_per_step": int(out_bytes) * world,
               "steps": Ke, "api": "formation_gym.make_vec_env(...).step(numpy actions) -> numpy obs/rews/dones/infos",
               "note": "H2D actions + fused step + D2H obs/reward/done/individual reward per step, pinned host "
                       "buffers, host sync every step (PCIe-bound: obs is 24N^2 B per env)"}
        del venv
        # Informational second figure (NOT the `e2e` key): the same VecEnv API with to_numpy=False -- the
        # observations stay in HBM for a policy network on the same GPU; per step the host uploads the actions
        # from pinned memory and reads rewards + dones back.
        try:
            venv = formation_gym.make_vec_env(a.scenario, hi - lo, N, a.episode_length, device=device, dtype=dtype,
                                              seed=0, env_offset=lo, to_numpy=False)
            venv.reset()
            act_p = torch.empty(hi - lo, N, 2, dtype=dtype).uniform_(-1, 1).pin_memory()
            act_d = torch.empty(hi - lo, N, 2, dtype=dtype, device=device)
            rew_p = torch.empty(hi - lo, N, 1, dtype=dtype).pin_memory()
            done_p = torch.empty(hi - lo, N, dtype=torch.bool).pin_memory()

            def dev_step():
                act_d.copy_(act_p, non_blocking=True)
                o, r, d, _ = venv.step(act_d)
                rew_p.copy_(r, non_blocking=True); done_p.copy_(d, non_blocking=True)
                torch.cuda.current_stream(device).synchronize()
            for _ in range(3):
                dev_step()
            Kd = max(10, 5 * Ke)
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(Kd):
                dev_step()
            dt_d = time.perf_counter() - t0
            barrier()
            td = torch.tensor([dt_d], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            e2e["obs_on_device"] = {
                "value": total_agents * Kd / float(td.item()), "unit": UNIT, "steps": Kd,
                "h2d_bytes_per_step": int(act_p.numel() * act_p.element_size()) * world,
                "d2h_bytes_per_step": int(rew_p.numel() * rew_p.element_size() + done_p.numel()) * world,
                "api": "make_vec_env(..., to_numpy=False).step(cuda actions) -> CUDA obs; rewards/dones read to host"}
            del venv
        except Exception as ex:                                   # informational only
            e2e["obs_on_device"] = {"error": repr(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    bytes_step = env.bytes_per_env_step() * (hi - lo)
    achieved = bytes_step / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        key = "%s_N%d_E%d_%s" % (a.scenario, N, hi - lo, a.dtype)
        traffic = tj.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    kname = ("fg::k_hd_warp<%s,%d,obs=%s> (fg_step_fused, warp-autonomous persistent kernel)"
             % (a.dtype, N, "yes" if not a.no_obs else "no")) if (a.scenario == "formation_hd_env" and N in (3, 9, 27)) \
        else "fg::k_step<%s> (fg_step_fused, tile kernel)" % a.dtype
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_env_step": env.bytes_per_env_step(),
                "share_of_step": kernel_ms / (ms / K),
                "timing": "CUDA events around replays of a CUDA graph holding only fused-step launches, each on its "
                          "own pre-sampled action buffer (region 1, which gives `value`, replays random-policy kernel "
                          "+ fused step per step)"}

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        from oracle import ref_loop_port as rp
        r = rp.time_port(a.scenario, N, a.episode_length, seconds=a.cpu_seconds)
        cpu_baseline = {"value": r["agent_steps_per_s"], "unit": UNIT, "cores": r["procs"], "kind": "port",
                        "sample": "%d processes x %.1f s of per-env stepping (%d env-steps total) of %s N=%d, "
                                  "episode_length %d; oracle/ref_loop_port.py (reference-structured numpy loop, "
                                  "scipy=%s)" % (r["procs"], a.cpu_seconds, r["env_steps"], a.scenario, N,
                                                 a.episode_length, r["scipy"])}

    also = None
    if world == 1 and not a.no_also:
        also = also_configs(formation_gym, torch, device, dtype, peak)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.dtype, "data": "synthetic", "config": workload_config(a, hi - lo, "gpu"),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
        "gpu_launches": launches, "episode_stats": ep, "also": also,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def flops_per_env_step(N):
    """Algorithmic flops per env-step, SURVEY.md 8(d): F(N) = 27.5 N^2 + 10 N (dense reference arithmetic,
    shared reward terms once per env, no credit removed for early-outs)."""
    return 27.5 * N * N + 10.0 * N


def also_configs(formation_gym, torch, device, dtype, peak):
    """Other BASELINE.json configs, measured briefly (informational; not the headline line)."""
    res = []
    fp32_peak = None
    try:
        from formation_gym import probe
        pk = probe.measure("ffma", device=device)                # scalar FFMA probe, CUDA events
        fp32_peak = pk["tflops"]
        res.append({"config": "FP32 peak probe (fg_fp32_probe, scalar FFMA, 8 CTAs/SM)", "tflops": fp32_peak,
                    "ffma2_tflops": probe.measure("ffma2", device=device)["tflops"]})
    except Exception as ex:
        res.append({"config": "FP32 peak probe", "error": repr(ex)[:200]})
    for name, scen, N, E, steps, mode in (
            ("configs[1] hd N=9 E=4096 (launch-bound; per-step launches)", "formation_hd_env", 9, 4096, 500, "step"),
            ("configs[1] hd N=9 E=4096 (CUDA graph of 25 per-step launches)", "formation_hd_env", 9, 4096, 40, "graph"),
            ("configs[1] hd N=9 E=4096 (in-kernel 25-step rollouts)", "formation_hd_env", 9, 4096, 40, "rollout"),
            ("hd N=9 E=131072, device controller get_action_BFS(ezpolicy) instead of the random policy",
             "formation_hd_env", 9, 131072, 50, "bfs"),
            ("configs[2] hd N=27 E=65536", "formation_hd_env", 27, 65536, 50, "step"),
            ("configs[3] hd N=243 E=1024 (one GPU's share of 8192)", "formation_hd_env", 243, 1024, 30, "step"),
            ("configs[3] hd N=243 E=8192 (all 8192 envs on one GPU: steady state, no wave tail)", "formation_hd_env", 243, 8192, 8, "step"),
            ("hd N=243 E=1024 state+reward only (no obs; fused step kernel alone, CUDA graph)", "formation_hd_env", 243, 1024, 8, "noobs"),
            ("hd N=243 E=8192 state+reward only (no obs; steady state)", "formation_hd_env", 243, 8192, 3, "noobs"),
            ("hd N=3 E=1048576", "formation_hd_env", 3, 1048576, 50, "step"),
            ("basic N=3 L=3 E=1048576", "basic_formation_env", 3, 1048576, 50, "step")):
        try:
            env = formation_gym.make_batched_env(scen, E, N, 25, device=device, dtype=dtype, seed=1,
                                                 write_obs=(mode != "noobs"))
            env.reset()

            graph = env.capture_steps(25) if mode == "graph" else None
            if mode == "noobs":
                # the step kernel alone: actions sampled once, 5 fused steps per CUDA graph (no launch gaps)
                env.sample_actions()
                graph = env.capture_steps(5, policy=lambda env_: None)

            def run(n):
                for _ in range(n):
                    if mode == "rollout":
                        env.rollout_random(25)
                    elif mode == "graph":
                        graph.replay()
                    elif mode == "bfs":
                        env.step(env.bfs_actions(3))
                    elif mode == "noobs":
                        graph.replay()
                    else:
                        env.sample_actions(); env.step(env.actions)
            run(5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(steps); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            env_steps = steps * (25 if mode in ("rollout", "graph") else 5 if mode == "noobs" else 1)
            gbs = env.bytes_per_env_step() * E * env_steps / (ms * 1e-3) / 1e9
            row = {"config": name, "agent_steps_per_s": E * N * env_steps / (ms * 1e-3),
                   "ms_per_env_step": ms / env_steps, "algorithmic_GBps": gbs, "hbm_frac": gbs / peak}
            if mode == "step":
                # the fused step kernel alone (actions pre-sampled, CUDA graph of 5 launches: no launch gaps, no
                # random-policy kernel) -- the figure comparable with the headline's roofline.frac
                env.sample_actions()
                g5 = env.capture_steps(5, policy=lambda env_: None)
                g5.replay(); torch.cuda.synchronize()
                reps = max(2, steps // 5)
                e0.record()
                for _ in range(reps):
                    g5.replay()
                e1.record(); torch.cuda.synchronize()
                kms = e0.elapsed_time(e1) / (5 * reps)
                row.update({"step_kernel_ms": kms,
                            "hbm_frac_step_kernel": env.bytes_per_env_step() * E / (kms * 1e-3) / 1e9 / peak})
                del g5
            if mode == "noobs":
                # the step+reward kernel without dense observations is pair-compute bound: FP32 roofline
                tf = flops_per_env_step(N) * E * env_steps / (ms * 1e-3) / 1e12
                row.update({"bound": "fp32", "algorithmic_TFLOPs": tf, "fp32_peak_TFLOPs": fp32_peak,
                            "fp32_frac": (tf / fp32_peak) if fp32_peak else None,
                            "flops_per_env_step": flops_per_env_step(N),
                            "note": "fused step kernel alone (actions pre-sampled, CUDA graph of 5 steps); "
                                    "fraction of the MEASURED scalar-FFMA peak"})
            res.append(row)
            del env
        except Exception as ex:  # keep the headline line alive
            res.append({"config": name, "error": repr(ex)[:200]})
    return res


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)

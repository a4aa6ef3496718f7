#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched 2-player TRON tick (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps 50 --warmup 5                      # our arm (CUDA kernels through the C ABI)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # reference arm: the reference's Python loop (port) on the host cores

A "step" is one fused tick (move, collision, winner, trail write, auto-reset, reward, both players'
observation planes) of every env on the GPU.  Workload: BASELINE config #2 (10x10 grid, uniform random actions,
1-plane bf16 observations, auto-reset) scaled from 4096 to --envs-per-gpu games so that state + observations are far
larger than L2 and the HBM-roofline fraction the metric asks for is meaningful; envs shard across ranks with no
collective (weak scaling).  One JSON line is printed by rank 0.
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

METRIC = "env-steps/s (batched 2-player TRON)"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 22)
    ap.add_argument("--width", type=int, default=10)
    ap.add_argument("--obs-dtype", default="bf16", choices=["bf16", "f32", "i8"])
    ap.add_argument("--enc", default="lut1", choices=["lut1", "popup3", "popup3_const", "none"])
    ap.add_argument("--layout", default="bits10", choices=["bits10", "tile8", "trail"],
                    help="state layout: 32-byte bit planes (10x10 only), int8 Tile.value grid, or trail-list records (pure ticks only)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--sustained-seconds", type=float, default=1.0, help="extra untimed-for-the-headline run of the same step (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-streams", dest="streams", action="store_false", help="skip the epsilon-greedy action-stream lines")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------------------------- CPU legs
def _c_oracle_worker(args):
    from oracle import c_oracle as oc
    from tron_b200 import abi
    n_envs, W, ticks, dt, enc, seed = args
    steps, sec = oc.bench_random(n_envs, W, W, ticks, dt, enc, seed)
    return steps, sec


def _py_port_worker(args):
    from oracle import py_port
    seed, env_steps, W = args[:3]
    t0 = time.perf_counter()
    steps, _ = py_port.play_random(seed, env_steps, W, W, with_pop_up=len(args) > 3 and args[3])
    return steps, time.perf_counter() - t0


def cpu_c_port(W, dt, enc, budget_s, cores):
    """The plain-C oracle (scalar port), one process per host core, same random-policy auto-reset workload."""
    import multiprocessing as mp
    from oracle import c_oracle as oc
    oc.build()
    n_envs = 4096
    _, sec = oc.bench_random(n_envs, W, W, 8, dt, enc, 1)  # calibrate
    ticks = max(8, int(budget_s / max(sec / 8, 1e-6)))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_c_oracle_worker, [(n_envs, W, ticks, dt, enc, 100 + i) for i in range(cores)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    return steps / max(r[1] for r in res), "%d procs x %d envs x %d ticks, C oracle, wall %.1fs" % (cores, n_envs, ticks, wall)


def cpu_py_port(W, env_steps_per_core, cores, with_pop_up=False):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_py_port_worker, [(i, env_steps_per_core, W, with_pop_up) for i in range(cores)])
    steps = sum(r[0] for r in res)
    return steps / max(r[1] for r in res), steps


def run_reference(a):
    """Reference arm: the reference's own Python Game.step loop (pure-Python port, oracle/py_port.py, validated against
    the live reference for results and cost), multiprocess over all host cores, same workload, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W = a.width
    # calibrate: ~1.5 s of work per core per step
    from oracle import py_port
    t0 = time.perf_counter()
    s, _ = py_port.play_random(0, 300, W, W)
    rate1 = s / (time.perf_counter() - t0)
    per = max(200, int(min(2.5, 120.0 / max(a.steps, 1)) * rate1))  # ~2.5 s of work per core per step, whole run bounded to ~2 min
    for _ in range(a.warmup):
        cpu_py_port(W, max(50, per // 8), cores)
    total, dt = 0, 0.0
    for _ in range(a.steps):  # time = slowest worker's own loop time (pool start-up excluded, as BASELINE.md section 4 asks)
        rate, st = cpu_py_port(W, per, cores)
        total += st
        dt += st / rate
    v = total / dt
    popup_rate, _ = cpu_py_port(W, max(100, per // 4), cores, with_pop_up=True)  # BASELINE.md section 4 "variant B": + pop_up on both observations
    c_v, c_sample = cpu_c_port(W, 3, 1, min(a.cpu_seconds, 6.0), cores)
    sample = "%d host procs x %d env-steps per bench step, pure-Python port of Game.step (results and cost validated vs the live reference)" % (cores, per)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / max(a.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8",
        "data": "synthetic", "config": workload_config(a, None),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "cpu_c_port": {"value": c_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": c_sample},
        "python_loop_with_pop_up": {"value": popup_rate, "unit": UNIT, "cores": cores, "note": "same loop + pop_up on both observations (DDQN data path)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------- ours
def workload_config(a, n_envs):
    return {"workload": "BASELINE config #2 scaled: %dx%d grid, uniform random actions (pre-generated u8 tape in HBM), auto-reset, "
                        "%s observations (%s), fused step+obs, state layout %s" % (a.width, a.width, a.obs_dtype, a.enc, a.layout),
            "envs_per_gpu": n_envs, "grid": [a.width, a.width], "obs_dtype": a.obs_dtype, "obs_enc": a.enc, "state_layout": a.layout,
            "l2_policy": "inputs larger than L2 (state+obs per GPU >> 126 MB), no flush", "parallelism": "env-sharded, no collective"}


def run_ours(a):
    import torch
    import torch.distributed as dist
    import tron_b200
    from tron_b200 import abi
    from tron_b200.batch_env import BatchedTron, HostTron

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N, W = a.envs_per_gpu, a.width
    tdt = {"bf16": torch.bfloat16, "f32": torch.float32, "i8": torch.int8}[a.obs_dtype]
    if a.layout == "bits10" and W != 10:
        a.layout = "tile8"
    env = BatchedTron(N, W, W, device=dev, obs_dtype=tdt, obs_enc=a.enc, reward="ddqn", auto_reset=True, seed=0, env_id_base=rank * N, layout=a.layout)
    obs = env.reset()
    # synthetic random-policy action stream, resident in HBM before the timed region
    tape = [env.random_actions(1000 + i) for i in range(4)]
    reward = torch.empty((N, 2), dtype=torch.float32, device=dev)
    done = torch.empty(N, dtype=torch.uint8, device=dev)
    winner = torch.empty(N, dtype=torch.uint8, device=dev)

    def one_step(i):
        env.step(tape[i & 3], obs=obs, reward=reward, done=done, winner=winner, want_ep_len=False)

    for i in range(a.warmup):
        one_step(i)
    torch.cuda.synchronize()
    st0 = env.stats_dict()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(a.steps):
        one_step(i)
    ev1.record()
    torch.cuda.synchronize(); barrier()
    t_wall1 = time.perf_counter()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    st1 = env.stats_dict()
    # sustained check (not the headline): the same step for >= 1 s, to show the K-step number is not a burst artefact
    sustained = None
    if a.sustained_seconds > 0:
        n_sus = max(a.steps, int(a.sustained_seconds / (ms / a.steps * 1e-3)))
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
            time.sleep(0.1)
        barrier(); torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        s0.record()
        for i in range(n_sus):
            one_step(i)
        s1.record()
        torch.cuda.synchronize(); barrier()
        tw1 = time.perf_counter()
        ms_sus = max_over_ranks(s0.elapsed_time(s1))
        sustained = {"steps": n_sus, "seconds": ms_sus * 1e-3, "value": world * N * n_sus / (ms_sus * 1e-3), "unit": UNIT,
                     "clocks": sampler2.stop(tw0, tw1) if rank == 0 else None}
    env_steps = st1["env_steps"] - st0["env_steps"]
    assert env_steps == N * a.steps, (env_steps, N * a.steps)
    f_reset = (st1["episodes"] - st0["episodes"]) / env_steps
    value = sum_over_ranks(float(env_steps)) / (ms * 1e-3)

    # roofline of the one kernel in the step: algorithmic bytes per env-step x envs per launch / launch time
    C, P = env.C, env.P
    b_o = {"bf16": 2, "f32": 4, "i8": 1}[a.obs_dtype]
    M = 48
    grid_bytes = 32 if a.layout == "bits10" else C  # state bytes per game as stored (C * b_g)
    if a.layout == "trail":
        grid_bytes = 0  # a reset writes nothing but the 16-byte header
    if P:
        bytes_per_env_step = grid_bytes * (1 + f_reset) + 2 * P * C * b_o + M
    else:  # pure tick: <=4 sectors read+written, one metadata sector each way, reset amortisation
        bytes_per_env_step = 320 + f_reset * grid_bytes
        if a.layout == "trail":
            bytes_per_env_step = 64 + 64 + 16  # record head read + written back, actions + reward/done/winner
    launch_ms = ms / a.steps
    achieved = bytes_per_env_step * N / (launch_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s_%s_%s_%d" % (a.layout, a.obs_dtype, a.enc, N))
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "step_bits10_kernel" if a.layout == "bits10" else ("step_trail_kernel" if a.layout == "trail" else "step_sparse_kernel" if (not P and C >= 1024) else "step_tile_kernel"), "state_bytes_per_game": grid_bytes, "bytes_per_env_step": bytes_per_env_step, "reset_fraction": f_reset,
                "envs_per_launch": N, "launch_ms": launch_ms, "frac_of_8TBs_nominal": achieved / 8000.0}

    # e2e: the same workload through the host-buffer C-ABI front end (pinned host arrays in, host arrays out)
    e2e = None
    if not a.no_e2e:
        del obs
        torch.cuda.empty_cache()
        h = HostTron(N, W, W, obs_dtype={"bf16": abi.BF16, "f32": abi.F32, "i8": abi.I8}[a.obs_dtype],
                     obs_enc={"lut1": abi.ENC_LUT1, "popup3": abi.ENC_POPUP3, "popup3_const": abi.ENC_POPUP3_CONST, "none": abi.ENC_NONE}[a.enc],
                     reward="ddqn", seed=0, env_id_base=rank * N, n_chunks=16, layout=a.layout)
        h.reset()
        host_tape = [t.cpu().numpy() for t in tape]
        h.step(host_tape[0])
        barrier()
        t0 = time.perf_counter()
        for i in range(a.e2e_steps):
            h.step(host_tape[i & 3])
        dt_e2e = max_over_ranks(time.perf_counter() - t0)
        barrier()
        frame = 2 * P * C * b_o
        e2e = {"value": sum_over_ranks(float(N * a.e2e_steps)) / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * N, "d2h_bytes_per_step": N * (frame + 8 + 2),
               "steps": a.e2e_steps, "api": "tron_host_env_step (pinned host buffers, 16 chunks over 4 streams)", "checksum": float(h.reward[:1024].sum())}
        h.close()
        if a.obs_dtype != "i8" and P:  # same call with int8 observations (the reference returns integer arrays): 1/2 resp. 1/4 of the PCIe bytes
            h8 = HostTron(N, W, W, obs_dtype=abi.I8, obs_enc={"lut1": abi.ENC_LUT1, "popup3": abi.ENC_POPUP3, "popup3_const": abi.ENC_POPUP3_CONST}[a.enc],
                          reward="ddqn", seed=0, env_id_base=rank * N, n_chunks=16, layout=a.layout)
            h8.reset(); h8.step(host_tape[0])
            barrier()
            t0 = time.perf_counter()
            for i in range(a.e2e_steps):
                h8.step(host_tape[i & 3])
            dt8 = max_over_ranks(time.perf_counter() - t0)
            barrier()
            e2e["int8_obs_variant"] = {"value": sum_over_ranks(float(N * a.e2e_steps)) / dt8, "unit": UNIT, "d2h_bytes_per_step": N * (2 * P * C + 10)}
            h8.close()

    # epsilon-greedy action streams (SURVEY 8d proxy computed in-kernel: with prob. eps uniform, else a random FREE neighbour): fewer resets
    streams = None
    if rank == 0 and a.streams:
        streams = []
        for eps in (1.0, 0.5, 0.1, 0.003):
            env_e = BatchedTron(N, W, W, device=dev, obs_dtype=tdt, obs_enc=a.enc, reward="ddqn", seed=3, layout=a.layout, policy="free_eps", policy_epsilon=eps)
            obs_e = env_e.reset()
            for i in range(40):  # let episode lengths reach their stationary mix
                env_e.step(obs=obs_e, reward=reward, done=done, winner=winner, want_ep_len=False)
            torch.cuda.synchronize()
            q0 = env_e.stats_dict()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                env_e.step(obs=obs_e, reward=reward, done=done, winner=winner, want_ep_len=False)
            e1.record(); torch.cuda.synchronize()
            q1 = env_e.stats_dict()
            fe = (q1["episodes"] - q0["episodes"]) / (q1["env_steps"] - q0["env_steps"])
            be = (grid_bytes * (1 + fe) + 2 * P * C * b_o + M) if P else bytes_per_env_step
            rate = N * a.steps / (e0.elapsed_time(e1) * 1e-3)
            streams.append({"epsilon": eps, "value": rate, "unit": UNIT, "reset_fraction": fe, "bytes_per_env_step": be, "frac_of_peak": rate * be / 1e9 / peak})
            del env_e, obs_e
            torch.cuda.empty_cache()

    # launch-bound regime of the literal config #2 size: 4096 envs, T ticks per launch (tron_step_many)
    small = None
    if rank == 0:
        env_s = BatchedTron(4096, W, W, device=dev, obs_dtype=tdt, obs_enc=a.enc, seed=0, layout=a.layout)
        env_s.reset()
        env_s.step_many(64)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            env_s.step_many(64)
        e1.record(); torch.cuda.synchronize()
        small = {"envs": 4096, "ticks_per_launch": 64, "value": 4096 * 64 * 10 / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
                 "note": "literal BASELINE config #2 size: L2-resident and launch-bound, obs written every tick"}
        from tron_b200.batch_env import GraphedStep
        gs = GraphedStep(BatchedTron(4096, W, W, device=dev, obs_dtype=tdt, obs_enc=a.enc, seed=1, layout=a.layout), policy="random", ticks_per_replay=16)
        gs.env.reset()
        gs.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(40):
            gs.replay()
        e1.record(); torch.cuda.synchronize()
        small["cuda_graph_16_ticks_per_replay"] = {"value": 4096 * 16 * 40 / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
                                                   "note": "one tron_step launch per tick, 16 ticks captured in a CUDA graph, device-side RNG counter"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample = cpu_c_port(W, {"bf16": abi.BF16, "f32": abi.F32, "i8": abi.I8}[a.obs_dtype],
                               {"lut1": abi.ENC_LUT1, "popup3": abi.ENC_POPUP3, "popup3_const": abi.ENC_POPUP3_CONST, "none": abi.ENC_NONE}[a.enc],
                               a.cpu_seconds, cores)
        pv, psteps = cpu_py_port(W, 1500, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "python_loop_port": {"value": pv, "unit": UNIT, "cores": cores, "sample": "%d env-steps, pure-Python port of the reference's Game.step loop" % psteps}}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8", "data": "synthetic", "config": workload_config(a, N),
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": a.steps, "clocks": clocks, "sustained": sustained, "eps_greedy_streams": streams, "small_n": small,
               "episode_stats": {"reset_fraction": f_reset, "mean_episode_ticks": (st1["ep_ticks"] - st0["ep_ticks"]) / max(1, st1["episodes"] - st0["episodes"])}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

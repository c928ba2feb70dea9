#!/usr/bin/env python
"""bench.py -- MCTS node-expansions/sec of batched self-play (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our arm (one rank per GPU under torchrun)
  python bench.py --impl reference ...                    CPU arm: the oracle port of the reference's path

A "step" is one ply of EVERY concurrent game: a full MCTS_tree.main of `--playouts` playouts per game
(select / encode / network / expand / backup waves), then get_action's host-side move choice and the
re-root.  Workload = BASELINE.json configs[1]: 1024 concurrent games x 1200 playouts, res_block_nums=7,
per GPU (weak scaling).  Expansions are counted by the engine (calls of expand), not inferred.

value : expansions / device time of the search waves (CUDA events, state resident in HBM)
e2e   : expansions / time of the whole SelfPlay.step() loop through the public API, including the
        per-ply device->host read of root statistics / status and host->device write of the chosen
        moves, host move sampling and tuple recording.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "mcts_node_expansions_per_sec"
FLOPS_PER_EVAL = {7: 375.4e6, 19: 1012.4e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=1024, help="concurrent games per GPU")
    ap.add_argument("--playouts", type=int, default=1200)
    ap.add_argument("--res-blocks", type=int, default=7)
    ap.add_argument("--precision", default=os.environ.get("CCHESS_NN_PRECISION", "fp16"))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--first-conv", default=None, choices=["gather", "tc"], help="first-layer kernel: CUDA-core gather-add or tcgen05/TMEM")
    ap.add_argument("--overlap-movegen", action="store_true", help="leaf move generation on a side stream under the network (measured: no gain)")
    ap.add_argument("--lanes", type=int, default=1, choices=[1, 2], help="2 = pipeline two half-batches (tree kernel under the other half's network)")
    ap.add_argument("--library-ends", action="store_true", help="use cuDNN/cuBLAS for the first conv and the heads instead of csrc/cz_net.cu")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--profile-waves", type=int, default=200, help="waves timed individually for the roofline line")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def algorithmic_bytes(c0, c1, enc_bytes):
    """per playout sum_l 12*c_l + 36*L; per expansion 90 + encode + 4C+4 + 14C."""
    d = {k: c1[k] - c0[k] for k in ("n_expand", "n_playout", "sum_L", "sum_c", "sum_C")}
    return 12 * d["sum_c"] + 36 * d["sum_L"] + d["n_expand"] * (90 + enc_bytes + 4) + 18 * d["sum_C"], d


# ---------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may really use: affinity mask, capped by the cgroup CPU quota when there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = max(1, min(n, int(float(q[0]) / float(q[1]) + 0.5)))
    except Exception:
        pass
    return n


class CpuArm:
    """The reference's path on the host cores: the C oracle port (oracle/cchess_oracle.c) drives n_games trees in
    lock-step (one leaf per game per wave, search_threads=1 semantics) and the same seed-0 network is evaluated by
    PyTorch on the CPU.  The thread count is calibrated (a few candidates, one wave each) and the best is kept."""

    def __init__(self, n_games, playouts, res_blocks, threads=None):
        import ctypes as C
        from cchess_zero_b200.net import PolicyValueNet
        from oracle import oracle as O
        self.C, self.O, self.L = C, O, O.lib()
        torch.manual_seed(0)
        self.net = PolicyValueNet(res_blocks).eval().to(memory_format=torch.channels_last)
        self.B, self.playouts = n_games, playouts
        self.trees = [O.Tree() for _ in range(n_games)]
        self.arr = (C.c_void_p * n_games)(*[t.h for t in self.trees])
        self.side = np.zeros(n_games, dtype=np.int32)
        self.rr = np.zeros(n_games, dtype=np.int32)
        self.nn_in = np.zeros((n_games, 9, 10, 14), dtype=np.float32)
        self.pending = np.zeros(n_games, dtype=np.uint8)
        self.cores = host_cores()
        self.wave_s = None
        if threads is None:
            self.calibrate()
        else:
            self.threads = threads
            torch.set_num_threads(threads)

    def _p(self, a):
        return a.ctypes.data_as(self.C.c_void_p)

    def wave(self):
        L, p = self.L, self._p
        L.co_batch_select(self.arr, p(self.side), p(self.rr), self.B, self.playouts, p(self.nn_in), p(self.pending), self.threads)
        with torch.no_grad():
            lo, v = self.net(torch.from_numpy(self.nn_in))
        lo = np.ascontiguousarray(lo.numpy(), dtype=np.float32)
        v = np.ascontiguousarray(v.numpy().reshape(-1), dtype=np.float32)
        L.co_batch_finish(self.arr, self.B, p(lo), p(v), p(self.pending), self.threads)

    def calibrate(self):
        cands = sorted({c for c in (self.cores, self.cores // 2, 64, 32, 16, 8) if 1 <= c <= self.cores}, reverse=True)
        best, best_t = cands[0], None
        for c in cands:
            self.threads = c
            torch.set_num_threads(c)
            t0 = time.perf_counter(); self.wave(); dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = c, dt
        self.threads, self.wave_s = best, best_t
        torch.set_num_threads(best)

    def expansions(self):
        return sum(t.stats()["n_expand"] for t in self.trees)

    def run(self, seconds=None, waves=None):
        e0, t0, n = self.expansions(), time.perf_counter(), 0
        while True:
            self.wave(); n += 1
            if (waves is not None and n >= waves) or (waves is None and time.perf_counter() - t0 >= seconds):
                break
        dt = time.perf_counter() - t0
        return dict(value=(self.expansions() - e0) / dt, seconds=dt, waves=n)


def sized_cpu_arm(max_games, playouts, res_blocks, wave_budget_s):
    """Probe with 64 games (also calibrates the thread count), then size the sample so that one lock-step wave of the
    sample costs about wave_budget_s on this host.  Keeps every CPU leg bounded whatever the box's core count is."""
    probe = CpuArm(min(64, max_games), playouts, res_blocks)
    per_game = max(probe.wave_s, 1e-4) / probe.B
    n = int(min(max_games, max(32, wave_budget_s / per_game)))
    n = 1 << (n.bit_length() - 1)                                   # power of two <= n
    if n <= probe.B:
        return probe
    return CpuArm(min(n, max_games), playouts, res_blocks, threads=probe.threads)


def run_reference(a, rank, world):
    """--impl reference: rank 0 times the CPU arm (oracle port; the Python reference cannot travel), other ranks exit 0.
    Each step is a bounded sample of the workload: `waves_per_step` lock-step waves of a sample of the games, the sample
    sized so that the whole --steps/--warmup run takes about two minutes."""
    if rank != 0:
        return
    waves_per_step = max(1, int(os.environ.get("CCHESS_REF_WAVES_PER_STEP", "2")))
    total_budget = float(os.environ.get("CCHESS_REF_SECONDS", "120"))
    wave_budget = total_budget / max(1, (a.steps + a.warmup) * waves_per_step)
    arm = sized_cpu_arm(a.games, a.playouts, a.res_blocks, wave_budget)
    if a.warmup:
        arm.run(waves=a.warmup * waves_per_step)
    r = arm.run(waves=a.steps * waves_per_step)
    v = r["value"]
    sample = "%d of the %d games x %d lock-step waves per step (first waves of the %d-playout search from the start position), oracle C port + torch CPU fp32 net, %d threads (calibrated) of %d usable cores" % (
        arm.B, a.games, waves_per_step, a.playouts, arm.threads, arm.cores)
    line = dict(metric=METRIC, value=v, unit="expansions/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=r["seconds"] / a.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="%d concurrent self-play games x %d playouts, res_block_nums=%d" % (a.games, a.playouts, a.res_blocks),
                            games_per_gpu=a.games, playouts=a.playouts, res_block_nums=a.res_blocks),
                cpu_baseline=dict(value=v, unit="expansions/s", cores=arm.threads, kind="port", sample=sample),
                e2e=dict(value=v, unit="expansions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(a, rank, world, local_rank):
    from cchess_zero_b200.net import policy_value_network
    from cchess_zero_b200.selfplay import SelfPlay
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = a.games
    pv = policy_value_network(a.res_blocks, precision=a.precision, device=local_rank, seed=0)
    native = a.precision == "fp16" and not a.library_ends
    factory = (lambda n: pv.native_plan(n, a.first_conv)) if native else (lambda n: pv.plan())
    plan = factory(B // a.lanes)
    sp = SelfPlay(B, None, a.playouts, seeds=[rank * B + g for g in range(B)], device=local_rank,
                  auto_reset=True, keep_records=True, plan=plan if a.lanes == 1 else None, plan_factory=factory, lanes=a.lanes,
                  overlap_movegen=a.overlap_movegen)
    if not a.no_graph:
        sp.capture_graph()
    e = sp.engine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    search_ms = []
    orig_search = sp.search

    def timed_search():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w = orig_search()
        e1.record()
        search_ms.append((e0, e1))
        return w
    sp.search = timed_search

    gathered = 0
    for _ in range(a.warmup):
        sp.step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    search_ms.clear()
    c0 = e.counters()
    l0 = e.launches
    waves0, fin0 = sp.waves, len(sp.finished)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(a.steps):
        out = sp.step()
        if world > 1:
            gathered += gather_tuples(out["finished"], dev, world)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    c1 = e.raise_on_error()
    e2e_ms = ev0.elapsed_time(ev1)
    dev_ms = sum(x.elapsed_time(y) for x, y in search_ms)
    n_exp = c1["n_expand"] - c0["n_expand"]
    launches = e.launches - l0
    # max time over ranks, sum of expansions
    t = torch.tensor([dev_ms, e2e_ms, wall * 1e3], dtype=torch.float64, device=dev)
    n = torch.tensor([n_exp, launches, len(sp.finished) - fin0, sum(len(r) for _, r in sp.finished[fin0:])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms = [float(x) for x in t]
    tot_exp, tot_launch, games_done, tuples_done = [float(x) for x in n]

    # ---- roofline of the dominant kernel of OUR code (k_wave), timed live per launch ----
    roof = None
    cpu = None
    if rank == 0:
        hbm, peak_src, _ = measured_peaks()
        enc_bytes = 96 if plan.dtype == torch.uint8 else 1260 * (4 if plan.dtype == torch.float32 else 2)
        sp.search = orig_search
        e.begin_search(a.playouts)
        k0 = e.counters()
        evs = []
        lanes = sp.lanes if sp.lanes is not None else [sp]       # a single-lane SelfPlay has the same attribute names
        side = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        for _ in range(a.profile_waves):
            for ln in lanes:
                x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x.record(); ln.engine.wave(ln.nn_in, ln.logits, ln.value); y.record()
                evs.append((x, y))
                if sp.lanes is None and sp.overlap_movegen and not a.no_graph:      # same schedule as the captured graph
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        ln.engine.prepare_leaves()
                ln.forward(ln.nn_in)
                cur.wait_stream(side)
        torch.cuda.synchronize()
        k1 = e.counters()
        kms = [x.elapsed_time(y) for x, y in evs]
        ab, d = algorithmic_bytes(k0, k1, enc_bytes)
        per_launch = ab / len(kms)
        avg_ms = float(np.mean(kms))
        achieved = per_launch / (avg_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "kwave_traffic.json")
        if os.path.exists(tpath):        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, scaled to this launch size
            tj = json.load(open(tpath))
            traffic = (tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]) * (B // a.lanes) / tj["games_per_launch"]
            traffic_src = tj["source"]
        roof = dict(kernel="k_wave (expand+backup+select+encode, one warp per game; %d games per launch)" % (B // a.lanes), bound="hbm", achieved=achieved, peak=hbm, unit="GB/s",
                    frac=achieved / hbm, traffic=traffic, traffic_source=traffic_src, peak_source=peak_src, avg_launch_ms=avg_ms, p50_launch_ms=float(np.median(kms)), max_launch_ms=float(np.max(kms)),
                    algorithmic_bytes_per_launch=per_launch,
                    bytes_per_expansion=ab / max(1, d["n_expand"]), launches_timed=len(kms),
                    note="latency-bound pointer-chasing kernel: HBM fraction is reported as required, the binding limits are per-warp dependent loads and the network")
        if not a.no_cpu_baseline and world == 1:      # reported baseline: rank 0, N=1 only
            arm = sized_cpu_arm(B, a.playouts, a.res_blocks, wave_budget_s=a.cpu_seconds / 5.0)
            r = arm.run(seconds=a.cpu_seconds)
            cpu = dict(value=r["value"], unit="expansions/s", cores=arm.threads, kind="port",
                       sample="%d of the %d games x %d lock-step waves (%.1f s) from the start position, oracle C port + torch CPU fp32 net, %d threads (calibrated) of %d usable cores" % (
                           arm.B, B, r["waves"], r["seconds"], arm.threads, arm.cores))

    if rank == 0:
        value = tot_exp / (dev_ms * 1e-3)
        e2e_v = tot_exp / (e2e_ms * 1e-3)
        plies_per_game = tuples_done / games_done if games_done else None
        games_per_hour = (world * B * a.steps / (e2e_ms * 1e-3) * 3600.0 / plies_per_game) if plies_per_game else None
        h2d = B * 4 + B          # chosen child indices + search mask
        d2h = B * 4 + B * 128 * (2 + 4 + 4 + 4 + 4) + B * (1 + 1 + 4 + 4 + 1 + 90) + 4 * (sp.waves - waves0) // max(1, a.steps) // max(1, a.playouts)
        line = dict(metric=METRIC, value=value, unit="expansions/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=e2e_ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=a.precision, data="synthetic (seed-0 xavier-initialised network, all games from the start position)",
                    config=dict(workload="%d concurrent self-play games x %d playouts per move, res_block_nums=%d, per GPU" % (B, a.playouts, a.res_blocks),
                                games_per_gpu=B, playouts=a.playouts, res_block_nums=a.res_blocks, search_threads=1, exploration=True,
                                cuda_graph=not a.no_graph, lanes=a.lanes, movegen_under_network=bool(sp.overlap_movegen and not a.no_graph and a.lanes == 1),
                                fused_conv_epilogue=plan.fused,
                                network_ends=("csrc/cz_net.cu (board-byte first conv [%s], fused heads)" % plan.first_conv) if plan.dtype == torch.uint8 else "library",
                                l2_policy="working set (trees %.1f GB + activations) exceeds the 126 MB L2" % (c1["max_arena_words"] * 4 * B / 1e9)),
                    e2e=dict(value=e2e_v, unit="expansions/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, wall_ms=wall_ms),
                    gpu_launches=int(tot_launch), clocks=clk, roofline=roof, cpu_baseline=cpu,
                    extra=dict(expansions=tot_exp, waves_per_step=(sp.waves - waves0) / a.steps, games_finished=games_done,
                               plies_per_finished_game=plies_per_game, games_per_hour=games_per_hour, tuples_all_gathered=gathered,
                               nn_tflops=tot_exp * FLOPS_PER_EVAL.get(a.res_blocks, 0) / (dev_ms * 1e-3) / 1e12 / world,
                               max_arena_words=c1["max_arena_words"], max_depth=c1["max_depth"],
                               mean_L=(c1["sum_L"] - c0["sum_L"]) / max(1, c1["n_playout"] - c0["n_playout"]),
                               mean_children=(c1["sum_C"] - c0["sum_C"]) / max(1, n_exp)))
        print(json.dumps(line), flush=True)


def gather_tuples(finished, dev, world, cap=2048):
    """NCCL all_gather of the (s, pi, z) tuples of the games that ended this step."""
    from cchess_zero_b200.distributed import all_gather_tuples
    return len(all_gather_tuples([r for _, r in finished], dev, cap))


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    run_ours(a, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's own CPU self-play, timed (bench.py --impl reference, cpu_baseline).

What runs: the UNMODIFIED reference (`cchess_main.selfplay()` -> `get_action` -> `MCTS_tree.main`, main.py:1493-1554, 1332-1358,
473-493) with `search_threads=16`, `playout=1200`, `exploration=True`, exactly `python main.py --mode train --train_playout 1200
--search_threads 16 --processor cpu` minus the training step.  The reference is single-threaded asyncio, so
ONE PROCESS PER USABLE HOST CORE plays independent games (distinct np.random seeds).  The evaluator is the PyTorch
re-implementation of the network on the CPU with torch.set_num_threads(1) and the seed-0 weights of the GPU run -- the reference's
own TensorFlow evaluator cannot run (TensorFlow is not installed and no weights ship with the repo).

Source of the reference: oracle/stage_reference.py (the live /root/reference here, the staged byte-identical copy under
oracle/_ref/reference on the GPU box; sha256-checked against oracle/reference_manifest.json).

Instrumentation is a wrapper around `leaf_node.expand` (a counter + timestamps); no reference text is edited.  A "step" is a fixed
QUOTA of expansions per worker process, so K steps are an exact amount of work:
    worker w records t_w[j] = time at which its j-th quota completed; it runs W + K quotas of its endless self-play and exits.
    whole-job rate = (n_workers * K * quota) / max_w (t_w[W+K] - t_w[W])
The quota is sized by a short calibration so that the whole run fits the caller's time budget."""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def usable_cores():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = max(1, min(n, int(float(q[0]) / float(q[1]) + 0.5)))
    except Exception:
        pass
    return n


# ------------------------------------------------------------------------------------------------------------------
# worker: one process = one instance of the reference's single-threaded program
# ------------------------------------------------------------------------------------------------------------------
def worker(a):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    os.environ["CCHESS_REFERENCE_DIR"] = a.ref_dir
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import ref_harness as H
    ref = H.load_reference()

    if a.net == "zero":
        # tree-only figure: an evaluator that costs nothing (zeros would make every prior NaN: use a constant positive logit)
        def forward(x):
            n = len(x)
            return np.full((n, 2086), 1.0, dtype=np.float32), np.zeros((n, 1), dtype=np.float32)
    else:
        from cchess_zero_b200.net import PolicyValueNet          # the architecture only (plain PyTorch module, CPU)
        torch.manual_seed(0)
        net = PolicyValueNet(a.res_blocks).eval().to(memory_format=torch.channels_last)

        def forward(x):                                           # policy_value_network.forward signature (policy_value_network.py:202-214)
            with torch.no_grad():
                lo, v = net(torch.from_numpy(np.asarray(x, dtype=np.float32)).reshape(-1, 9, 10, 14))
            return lo.numpy(), v.numpy().reshape(-1, 1)

    st = dict(n=0, marks=[], batches=0, rows=0)
    quota, total_marks = a.quota, a.warmup + a.steps
    t_start = time.perf_counter()
    orig_expand = ref.leaf_node.expand

    class Done(Exception):
        pass

    def counting_expand(self, moves, action_probs):               # wrapper, not an edit: counts calls of expand (= the metric)
        orig_expand(self, moves, action_probs)
        st["n"] += 1
        if st["n"] % quota == 0:
            st["marks"].append(time.perf_counter())
            if len(st["marks"]) > total_marks:
                raise Done()

    ref.leaf_node.expand = counting_expand
    fwd0 = forward

    def counted_forward(x):
        st["batches"] += 1
        st["rows"] += len(x)
        return fwd0(x)

    m = H.make_cchess_main(counted_forward, a.playouts, search_threads=a.search_threads, exploration=True)
    np.random.seed(a.seed)
    plies = 0
    st["marks"].append(time.perf_counter())                       # mark 0 = start of quota 1
    try:
        with H.quiet(), np.errstate(all="ignore"):
            while True:                                           # cchess_main.run() without policy_update (main.py:1224-1231)
                _, n = m.selfplay()
                plies += n
    except Done:
        pass
    marks = st["marks"]
    W, K = a.warmup, a.steps
    out = dict(seed=a.seed, expansions=st["n"], quota=quota, marks=len(marks) - 1,
               timed_s=marks[W + K] - marks[W], warm_s=marks[W] - marks[0], setup_s=marks[0] - t_start,
               mean_batch=st["rows"] / max(1, st["batches"]), plies_finished_games=plies)
    sys.stdout.write("REFARM " + json.dumps(out) + "\n")
    sys.stdout.flush()
    os._exit(0)                                                   # the asyncio loop holds half-finished coroutines: leave at once


# ------------------------------------------------------------------------------------------------------------------
# driver
# ------------------------------------------------------------------------------------------------------------------
def _spawn(n, quota, steps, warmup, playouts, res_blocks, search_threads, net, ref_dir, seed0=1000):
    procs = []
    for w in range(n):
        cmd = [sys.executable, os.path.abspath(__file__), "--worker", "--quota", str(quota), "--steps", str(steps), "--warmup", str(warmup),
               "--playouts", str(playouts), "--res-blocks", str(res_blocks), "--search-threads", str(search_threads), "--net", net,
               "--seed", str(seed0 + w), "--ref-dir", ref_dir]
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = []
    for p in procs:
        so, se = p.communicate()
        line = [l for l in so.splitlines() if l.startswith("REFARM ")]
        if p.returncode != 0 or not line:
            raise RuntimeError("reference worker failed (rc %s): %s" % (p.returncode, se[-1500:]))
        outs.append(json.loads(line[-1][7:]))
    return outs


def available():
    sys.path.insert(0, HERE)
    import stage_reference as S
    d = S.staged_dir()
    return d if (d and S.verify(d)) else None


def run(steps, warmup, playouts=1200, res_blocks=7, search_threads=16, budget_s=150.0, procs=None, net="torch", quota=None):
    """Times `steps` quotas (after `warmup` quotas) on every usable core.  Returns a dict with value (expansions/s, whole job),
    the sample description and per-worker records.  budget_s: target wall time of the measured run (calibration excluded)."""
    ref_dir = available()
    if ref_dir is None:
        raise RuntimeError("staged reference not found (run `python oracle/stage_reference.py` where /root/reference exists)")
    cores = usable_cores()
    n = procs or cores
    if quota is None:
        # calibration: every core busy (the per-core rate depends on shared caches / memory bandwidth), 1 warm + 2 timed quotas of 16
        cal = _spawn(n, 16, 2, 1, playouts, res_blocks, search_threads, net, ref_dir, seed0=500)
        rate = min(32.0 / max(c["timed_s"], 1e-6) for c in cal)          # slowest worker, expansions/s
        quota = int(max(16, min(playouts, budget_s * rate / max(1, steps + warmup))))
        quota -= quota % 16 if quota >= 32 else 0
    t0 = time.perf_counter()
    outs = _spawn(n, quota, steps, warmup, playouts, res_blocks, search_threads, net, ref_dir)
    wall = time.perf_counter() - t0
    timed = max(o["timed_s"] for o in outs)
    value = n * steps * quota / timed
    return dict(value=value, cores=n, usable_cores=cores, quota=quota, timed_s=timed, wall_s=wall, ms_per_step=timed / steps * 1e3,
                mean_nn_batch=sum(o["mean_batch"] for o in outs) / len(outs), workers=outs, net=net, ref_dir=ref_dir,
                sample=("unmodified reference cchess_main.selfplay() (search_threads=%d, %d playouts, exploration on), one process on each of "
                        "%d host cores; step = %d expansions per process (%d per step over all processes), %d warm-up + %d timed steps; "
                        "evaluator = %s" % (search_threads, playouts, n, quota, n * quota, warmup, steps,
                                             "PyTorch CPU fp32 net, 1 thread per process (TensorFlow absent)" if net == "torch"
                                             else "zero-cost stand-in (tree-only rate)")))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--quota", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--playouts", type=int, default=1200)
    ap.add_argument("--res-blocks", type=int, default=7)
    ap.add_argument("--search-threads", type=int, default=16)
    ap.add_argument("--net", default="torch", choices=["torch", "zero"])
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--ref-dir", default="")
    ap.add_argument("--budget", type=float, default=60.0)
    ap.add_argument("--procs", type=int, default=None)
    a = ap.parse_args()
    if a.worker:
        worker(a)
    else:
        r = run(a.steps, a.warmup, a.playouts, a.res_blocks, a.search_threads, a.budget, a.procs, a.net)
        r.pop("workers")
        print(json.dumps(r, indent=1))

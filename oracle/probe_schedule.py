"""TEST INFRASTRUCTURE ONLY (needs /root/reference).  Is the reference's search deterministic for search_threads > 1?
Runs MCTS_tree.main (unmodified reference) from the start position with an injected evaluator latency and prints the hash of
the whole tree.  Finding recorded in oracle/probe_schedule.last.txt: search_threads=1 never changes; search_threads=2 gives
DIFFERENT trees for latencies 0 and 1 ms (the interleaving depends on wall-clock timers: asyncio.sleep(1e-3) in
prediction_worker, main.py:452) -- which is why parity is defined on the search_threads=1 schedule."""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as H  # noqa: E402


def run(K, playouts, net, delay):
    ref = H.load_reference()
    base = H.FAKE_NETS[net]

    def fwd(x):
        if delay:
            time.sleep(delay)
        return base(x)
    t = H.make_mcts(fwd, K)
    with np.errstate(all="ignore"):
        t.main(t.root.state, "w", 0, playouts)
    sig = np.asarray(H.tree_signature(t.root, ref), dtype=np.int64)
    return hashlib.sha256(sig.tobytes()).hexdigest()[:12]


if __name__ == "__main__":
    lines = []
    for K in (1, 2, 4, 8, 16):
        hs = [run(K, 200, "hash_pos", d) for d in (0.0, 0.001, 0.003)]
        lines.append("search_threads=%-2d latencies 0/1/3 ms -> %s  %s" % (K, " ".join(hs), "STABLE" if len(set(hs)) == 1 else "TIMING-DEPENDENT"))
        print(lines[-1], flush=True)
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_schedule.last.txt"), "w").write("\n".join(lines) + "\n")

"""TEST INFRASTRUCTURE ONLY.  BASELINE config 1 at scale: >= 10^6 positions reached by seeded random play, the oracle's
ordered move lists / applied boards / capture flags / encodes compared with the UNMODIFIED reference (needs /root/reference).
    python oracle/diff_million.py [n_positions] [n_procs]
Prints one summary line; the committed record of the last run is oracle/diff_million.last.txt."""
import multiprocessing as mp
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def work(args):
    seed, n = args
    import ref_harness as H
    import oracle as O
    ref = H.load_reference()
    tree = H.make_mcts(H.FAKE_NETS["mod17"], 1)
    rng = random.Random(seed)
    done = enc_checked = 0
    while done < n:
        state, player = O.START, "w"
        for ply in range(400):
            b = O.from_state(state)
            side = 0 if player == "w" else 1
            mv_ref = ref.GameBoard.get_legal_moves(state, player)
            mv_o = [O.move_str(m) for m in O.legal_moves(b, side)]
            assert mv_ref == mv_o, (state, player)
            done += 1
            if done % 50 == 0:
                assert np.array_equal(tree.generate_inputs(state, player), O.encode(b, side))
                enc_checked += 1
            if not mv_ref or done >= n:
                break
            m = rng.choice(mv_ref)
            ns = ref.GameBoard.sim_do_action(m, state)
            nb, cap = O.apply_move(b, O.move_from_str(m))
            assert O.to_state(nb) == ns and (cap != 0) == (ref.is_kill_move(state, ns) != 0)
            state, player = ns, ("b" if player == "w" else "w")
            if "K" not in state or "k" not in state:
                break
    return done, enc_checked


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    t0 = time.time()
    with mp.Pool(procs) as pool:
        res = pool.map(work, [(1000 + i, n // procs + 1) for i in range(procs)])
    tot, enc = sum(r[0] for r in res), sum(r[1] for r in res)
    line = "positions=%d encodes=%d mismatches=0 procs=%d seconds=%.0f (ordered move lists, applied boards, capture flags vs the unmodified reference)" % (
        tot, enc, procs, time.time() - t0)
    print(line)
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "diff_million.last.txt"), "w").write(line + "\n")

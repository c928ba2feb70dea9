"""TEST INFRASTRUCTURE ONLY -- golden self-play games of the UNMODIFIED reference at its default search_threads = 16.

cchess_main.selfplay() (main.py:1493-1554) is run with the reference's own MCTS_tree coroutines (tree_search, start_tree_search,
prediction_worker: main.py:337-493) on oracle/detloop.py's deterministic event loop in its canonical "busy" schedule -- the one
the real uvloop runs follow whenever they are not timing-sensitive (tests/golden/k16_stats.json.gz pins that: equal to recorded
uvloop runs on 240 / 240 positions) and the one the engine's search_threads=K mode implements.  Evaluators are the deterministic
stand-ins of fakenets_np.py; np.random is seeded per game as in gen_golden.py.

    python oracle/gen_golden_k16_selfplay.py        ->  tests/golden/selfplay_k16.json
"""
import asyncio
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import detloop as D          # noqa: E402
import ref_harness as H      # noqa: E402

GAMES = [("hash_pos", 48, 7, 16), ("hash_signed", 64, 3, 16), ("mod17", 40, 5, 16), ("hash_pos", 200, 11, 16), ("hash_signed", 96, 2026, 8),
         ("hash_pos", 120, 1, 4)]


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def play(net, playouts, seed, K):
    m = H.make_cchess_main(H.FAKE_NETS[net], playouts, K)
    loop = D.DetLoop("busy")
    asyncio.set_event_loop(loop)
    m.mcts.loop = loop                               # MCTS_tree.main runs self.loop.run_until_complete(...) (main.py:492)
    np.random.seed(seed)
    with H.quiet(), np.errstate(all="ignore"):
        data, n = m.selfplay()
    loop.close()
    data = list(data)
    pis = np.asarray([d[1] for d in data], dtype=np.float64)
    sparse = [[[int(i), float(p[i]).hex()] for i in np.nonzero(p)[0]] for p in pis]
    return dict(net=net, playouts=playouts, seed=seed, search_threads=K, n=n, states=[d[0] for d in data], z=[float(d[2]) for d in data],
                sha_pi=sha(pis.tobytes()), pi_sparse=sparse)


def main():
    H.load_reference()
    games = []
    for net, playouts, seed, K in GAMES:
        g = play(net, playouts, seed, K)
        games.append(g)
        print("selfplay_k16", net, playouts, seed, K, g["n"], g["z"][0], flush=True)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "selfplay_k16.json")
    with open(out, "w") as f:
        json.dump(dict(schedule="detloop busy (canonical uvloop FIFO schedule)", games=games), f)
    print("wrote", out)


if __name__ == "__main__":
    main()

"""BASELINE config 5: ai_count=2 play mode, mcts vs mcts, 1200 playouts -- move latency through the reference-shaped
API (cchess_main.select_move('mcts'), exploration off, one tree, search_threads=1 semantics)."""
import contextlib, io, json, sys, time
sys.path.insert(0, '.')
import numpy as np
from cchess_zero_b200.net import policy_value_network
from cchess_zero_b200.selfplay import cchess_main

playouts = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 7
K = int(sys.argv[4]) if len(sys.argv) > 4 else 1
T = int(sys.argv[5]) if len(sys.argv) > 5 else 1      # search_threads: 1 = one playout at a time, 16 = the reference's default coroutine schedule (exact)
with contextlib.redirect_stdout(io.StringIO()):
    pv = policy_value_network(res_block_nums=blocks)
    m = cchess_main(playout=playouts, in_search_threads=T, network=pv, exploration=False, log_file=False, leaf_parallel=K)
np.random.seed(0)
lat = []
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(moves + 2):
        t0 = time.perf_counter()
        m.select_move("mcts")
        dt = time.perf_counter() - t0
        if i >= 2:
            lat.append(dt)
        if m.check_end()[0]:
            m.game_borad.reload(); m.mcts.reload()
lat = np.array(lat)
print(json.dumps(dict(metric="move_latency_s", config="1 game, mcts vs mcts, %d playouts, res_block_nums=%d, exploration off, select_move('mcts'), search_threads=%d, leaf_parallel=%d%s" % (playouts, blocks, T, K, "" if K == 1 else " (virtual-loss batching, not the reference's K=1 visit counts)"),
                      moves=len(lat), p50=float(np.median(lat)), p95=float(np.percentile(lat, 95)), mean=float(lat.mean()), max=float(lat.max()),
                      playouts_per_s=playouts / float(np.median(lat)))))

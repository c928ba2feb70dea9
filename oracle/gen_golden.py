"""Generates tests/golden/*.json by RUNNING THE UNMODIFIED REFERENCE in this container.

TEST INFRASTRUCTURE ONLY.  Usage (needs /root/reference; ~3-4 minutes):
    python oracle/gen_golden.py
The fixtures are what pins oracle/cchess_oracle.c (tests/test_oracle_golden.py) and, through
it, the CUDA path.  Everything recorded here is an OUTPUT OF THE REFERENCE'S OWN CODE:
  labels.json    create_uci_labels / unflipped_index            (main.py:23-65, 211-217)
  movegen.json   GameBoard.get_legal_moves / sim_do_action / is_kill_move /
                 MCTS_tree.try_flip / generate_inputs            (main.py:219-227, 531-574, 647-702, 743-1109)
                 + set-equality against the GUI rules ChessBoard/chessman/*.can_move
  tree.json      MCTS_tree.main with search_threads=1            (main.py:93-206, 337-493)
  selfplay.json  cchess_main.selfplay / get_action               (main.py:1332-1358, 1493-1554)
  play.json      select_move / get_hint / human_move / check_end  (main.py:1278-1329, 1380-1491), human_color b and w
The evaluator is one of the deterministic stand-in nets of oracle/ref_harness.py (the TF
network cannot run here: SURVEY 0.8) -- NN parity is NOT pinned by these files.
"""
import gzip
import hashlib
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
START = "RNBAKABNR/9/1C5C1/P1P1P1P1P/9/9/p1p1p1p1p/1c5c1/9/rnbakabnr"
APPENDIX_B = [
    ("4K4/9/9/9/9/9/9/9/9/4k4", "w"),
    ("4K4/9/9/9/9/9/9/9/9/4k4", "b"),
    ("R1BAKAB1R/9/1C2C1N2/P1P1P1P1P/2N6/6p2/p1p1p3p/1c2c1n2/9/rnbakab1r", "b"),
    ("4K4/9/9/9/9/2P1p4/2p1P4/9/9/4k4", "w"),
    ("4K4/9/9/9/9/2P1p4/2p1P4/9/9/4k4", "b"),
    ("3AK4/4A4/9/9/4c4/9/9/4C4/4a4/3k1a3", "w"),
    ("3K5/9/9/9/9/9/9/9/9/5k3", "w"),
]


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def gen_labels(ref):
    return dict(labels=ref.labels_array, unflipped_index=ref.unflipped_index,
                sha_labels=sha("\n".join(ref.labels_array).encode()),
                sha_unflipped=sha(np.asarray(ref.unflipped_index, dtype=np.int32).tobytes()))


# ---- second legality oracle: the GUI rules (ChessBoard.py / chessman/*.py) ----
def gui_move_set(state, player):
    import ChessBoard as CB  # reference module
    rows = H.load_reference().GameBoard.board_to_pos_name(state)
    cls = {"K": CB.Shuai, "A": CB.Shi, "R": CB.Che, "B": CB.Xiang, "N": CB.Ma, "P": CB.Bing, "C": CB.Pao}
    board = CB.ChessBoard.__new__(CB.ChessBoard)
    CB.ChessBoard.pieces = dict()
    for y in range(10):
        for x in range(9):
            c = rows[y][x]
            if c.isalpha():
                red = c.isupper()
                CB.ChessBoard.pieces[x, y] = cls[c.upper()](x, y, red, "north" if red else "south")
    out = set()
    with H.quiet():
        for (x, y), p in list(CB.ChessBoard.pieces.items()):
            if p.is_red == (player == "w"):
                for (nx, ny) in p.get_move_locs(board):
                    out.add("abcdefghi"[x] + str(y) + "abcdefghi"[nx] + str(ny))
    return out


def gen_movegen(ref, n_games=40, seed=20260923):
    tree = H.make_mcts(H.FAKE_NETS["mod17"], 1)
    rng = random.Random(seed)
    recs, n_gui = [], 0

    def record(state, player, chosen=None):
        nonlocal n_gui
        moves = ref.GameBoard.get_legal_moves(state, player)
        enc = tree.generate_inputs(state, player)
        flip, _ = tree.try_flip(state, player, tree.is_black_turn(player))
        r = dict(state=state, player=player, moves=" ".join(moves), flip=flip,
                 enc=[int(i) for i in np.nonzero(enc.reshape(-1))[0]])
        if "K" in state and "k" in state:
            assert gui_move_set(state, player) == set(moves), (state, player)
            n_gui += 1
        if chosen is None and moves:
            chosen = rng.choice(moves)
        if chosen is not None:
            nxt = ref.GameBoard.sim_do_action(chosen, state)
            r.update(move=chosen, next=nxt, kill=int(ref.is_kill_move(state, nxt)))
        recs.append(r)
        return r

    for s, p in APPENDIX_B:
        record(s, p)
    for g in range(n_games):
        state, player = START, "w"
        for ply in range(300):
            r = record(state, player)
            if "next" not in r:
                break
            state, player = r["next"], ("b" if player == "w" else "w")
            if "K" not in state or "k" not in state:
                record(state, player)  # movegen on a king-less board is still defined
                break
    return dict(n=len(recs), n_checked_against_gui_rules=n_gui, records=recs)


def gen_tree(ref):
    cases = []

    def run(net, state, player, rr, playouts, note=""):
        t = H.make_mcts(H.FAKE_NETS[net], 1, state)
        with np.errstate(all="ignore"):
            t.main(state, player, rr, playouts)
        sig = np.asarray(H.tree_signature(t.root, ref), dtype=np.int64).reshape(-1, 6)
        root = [[a, int(c.N), H.f32_bits(c.W), H.f32_bits(c.P), H.f32_bits(c.Q)] for a, c in t.root.child.items()]
        cases.append(dict(net=net, state=state, player=player, rr=rr, playouts=playouts, note=note,
                          n_nodes=int(sig.shape[0]), sha_sig=sha(sig.tobytes()), root=root,
                          head=sig[:40].tolist()))
        print("tree", net, player, rr, playouts, sig.shape[0], note, flush=True)

    for net in ("mod17", "hash_signed", "hash_pos"):
        run(net, START, "w", 0, 64, "start")
    run("hash_pos", START, "w", 0, 1200, "config-2 playout count")
    run("hash_signed", START, "w", 0, 600, "start deep")
    mid = "R1BAKAB1R/9/1C2C1N2/P1P1P1P1P/2N6/6p2/p1p1p3p/1c2c1n2/9/rnbakab1r"
    run("hash_pos", mid, "b", 3, 300, "midgame black")
    run("hash_signed", mid, "b", 3, 300, "midgame black")
    run("hash_pos", mid, "w", 58, 300, "draw rule: rr reaches 60 inside the tree")
    run("hash_signed", mid, "b", 59, 200, "draw rule at depth 1")
    endg = "3AK4/4A4/9/9/4c4/9/9/4C4/4a4/3k1a3"
    run("hash_pos", endg, "w", 0, 400, "king capture reachable")
    run("hash_signed", endg, "b", 10, 400, "king capture reachable")
    run("mod17", "4K4/9/9/9/9/2P1p4/2p1P4/9/9/4k4", "w", 0, 300, "flying general")
    run("hash_pos", "4K4/9/9/9/9/9/9/9/9/4k4", "w", 0, 100, "bare kings facing: immediate king capture")
    return dict(cases=cases)


def gen_tree_extra(ref, n_cases=72, seed=77):
    """More search trees from the reference: positions sampled along seeded random games (opening to bare endgames), both
    sides to move, restrict_round near and far from the 60-ply rule.  Pins the oracle only (CPU test)."""
    rng = random.Random(seed)
    nets = ("hash_pos", "hash_signed", "mod17")
    cases = []
    while len(cases) < n_cases:
        state, player = START, "w"
        plies = rng.randint(0, 220)
        for _ in range(plies):
            mv = ref.GameBoard.get_legal_moves(state, player)
            state = ref.GameBoard.sim_do_action(rng.choice(mv), state)
            player = "b" if player == "w" else "w"
            if "K" not in state or "k" not in state:
                break
        if "K" not in state or "k" not in state:
            continue
        net = nets[len(cases) % 3]
        rr = rng.choice([0, 1, 10, 40, 55, 57, 58, 59])
        playouts = rng.choice([60, 120, 200, 300])
        t = H.make_mcts(H.FAKE_NETS[net], 1, state)
        with np.errstate(all="ignore"):
            t.main(state, player, rr, playouts)
        sig = np.asarray(H.tree_signature(t.root, ref), dtype=np.int64).reshape(-1, 6)
        cases.append(dict(net=net, state=state, player=player, rr=rr, playouts=playouts, n_nodes=int(sig.shape[0]), sha_sig=sha(sig.tobytes())))
        print("tree_extra", len(cases), net, player, rr, playouts, sig.shape[0], flush=True)
    return dict(cases=cases)


def gen_selfplay(ref):
    games = []
    for net, playouts, seed in [("hash_pos", 30, 7), ("hash_signed", 20, 3), ("hash_pos", 60, 11),
                                ("mod17", 25, 5), ("hash_signed", 48, 2026), ("hash_pos", 100, 1)]:
        m = H.make_cchess_main(H.FAKE_NETS[net], playouts, 1)
        np.random.seed(seed)
        with H.quiet(), np.errstate(all="ignore"):
            data, n = m.selfplay()
        data = list(data)
        states = [d[0] for d in data]
        pis = np.asarray([d[1] for d in data], dtype=np.float64)
        z = [float(d[2]) for d in data]
        sparse = []
        for p in pis:
            nz = np.nonzero(p)[0]
            sparse.append([[int(i), float(p[i]).hex()] for i in nz])
        games.append(dict(net=net, playouts=playouts, seed=seed, n=n, states=states, z=z,
                          sha_pi=sha(pis.tobytes()), pi_sparse=sparse))
        print("selfplay", net, playouts, seed, n, z[0], flush=True)
    return dict(games=games)


def gen_play(ref):
    """Play-mode surface (main.py:1278-1329, 1394-1491): select_move / get_hint / human_move / check_end, both human colours."""
    def fhex(v):
        return float(v).hex()

    scripts = []
    for hc, net, playouts, seed in (("b", "hash_pos", 40, 5), ("w", "hash_signed", 32, 8)):
        m = H.make_cchess_main(H.FAKE_NETS[net], playouts, 1, exploration=False, human_color=hc)
        np.random.seed(seed)
        steps = []
        with H.quiet(), np.errstate(all="ignore"):
            for rnd in range(5):
                mv, wr = m.select_move("mcts")
                steps.append(dict(op="select_move_mcts", move=[int(x) for x in mv], win_rate=fhex(wr), state=m.game_borad.state,
                                  player=m.game_borad.current_player, rr=m.game_borad.restrict_round))
                if m.check_end()[0]:
                    break
                hint = m.get_hint("mcts", True, lambda: None)
                steps.append(dict(op="get_hint_mcts", hint=[[a, fhex(p)] for a, p in hint]))
                a = hint[min(rnd, len(hint) - 1)][0]                  # the human plays the (rnd+1)-th suggestion
                coord = (ord(a[0]) - 97, int(a[1]), ord(a[2]) - 97, int(a[3]))
                wr = m.human_move(coord, "mcts")
                steps.append(dict(op="human_move_mcts", coord=list(coord), win_rate=fhex(wr), state=m.game_borad.state,
                                  player=m.game_borad.current_player, rr=m.game_borad.restrict_round))
                ended, who = m.check_end()
                steps.append(dict(op="check_end", ended=bool(ended), who=who))
                if ended:
                    break
                hint = m.get_hint("net", False, lambda: None)
                steps.append(dict(op="get_hint_net", hint=[[a, fhex(p)] for a, p in hint]))
        scripts.append(dict(human_color=hc, net=net, playouts=playouts, seed=seed, steps=steps))
        print("play", hc, net, len(steps), flush=True)
    # select_move('net') on its own tree (it does not touch the search tree, main.py:1437-1462)
    m = H.make_cchess_main(H.FAKE_NETS["hash_pos"], 8, 1, exploration=False)
    steps = []
    with H.quiet(), np.errstate(all="ignore"):
        for _ in range(6):
            mv, wr = m.select_move("net")
            steps.append(dict(op="select_move_net", move=[int(x) for x in mv], win_rate=fhex(wr), state=m.game_borad.state))
    scripts.append(dict(human_color="b", net="hash_pos", playouts=8, seed=None, steps=steps))
    return dict(scripts=scripts)


def main():
    ref = H.load_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("labels", gen_labels), ("movegen", gen_movegen), ("tree", gen_tree), ("selfplay", gen_selfplay), ("play", gen_play), ("tree_extra", gen_tree_extra)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        d = fn(ref)
        d["generator"] = "oracle/gen_golden.py on the unmodified reference @7661eea, numpy %s" % np.__version__
        path = os.path.join(OUT, name + (".json.gz" if name == "movegen" else ".json"))
        blob = json.dumps(d, separators=(",", ":")).encode()
        if path.endswith(".gz"):
            with gzip.GzipFile(path, "wb", mtime=0) as f:
                f.write(blob)
        else:
            with open(path, "wb") as f:
                f.write(blob)
        print(name, os.path.getsize(path), "bytes", flush=True)


if __name__ == "__main__":
    main()

"""CPU tests of the UCCI front-end (cchess_zero_b200/ucci.py): FEN <-> reference state strings, move application against the
oracle, and whole protocol sessions driven through cchess_main.get_action/check_end with an oracle-backed stand-in for the
device tree (the GPU twin of the session test lives in test_gpu_zz_ucci.py)."""
import io
from collections import OrderedDict

import numpy as np
import pytest

from cchess_zero_b200 import ucci
from oracle import oracle as O


def test_fen_round_trip_and_aliases():
    st, pl, half = ucci.fen_to_state(ucci.START_FEN)
    assert (st, pl, half) == (O.START, "w", 0) and ucci.START_STATE == O.START
    assert ucci.state_to_fen(O.START, "w", 0, 1) == ucci.START_FEN
    st2, pl2, half2 = ucci.fen_to_state("rheakaehr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RHEAKAEHR b - - 7 12")
    assert st2 == O.START and pl2 == "b" and half2 == 7
    assert ucci.fen_to_state("4k4/9/9/9/9/9/9/9/9/4K4 r")[1] == "w"
    for bad in ("", "9/9/9", "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNX w", "8/9/9/9/9/9/9/9/9/9 w",
                "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR x"):
        with pytest.raises(ucci.UcciError):
            ucci.fen_to_state(bad)


def test_apply_move_matches_the_oracle_over_random_games():
    rng = np.random.RandomState(3)
    b, side, state = O.from_state(O.START), 0, O.START
    for ply in range(600):
        mv = O.legal_moves(b, side)
        m = mv[rng.randint(len(mv))]
        b, cap = O.apply_move(b, m)
        state, capch = ucci.apply_move(state, O.move_str(m))
        assert state == O.to_state(b) and bool(cap) == bool(capch)
        side ^= 1
        if cap in (1, 8):
            b, side, state = O.from_state(O.START), 0, O.START
    with pytest.raises(ucci.UcciError):
        ucci.apply_move(O.START, "a1a2")          # empty source square
    with pytest.raises(ucci.UcciError):
        ucci.parse_move("j0a1")


# ---- protocol sessions ----------------------------------------------------------------------------------------------

class _Child:
    def __init__(self, N, Q):
        self.N, self.Q = N, Q


class _Root:
    def __init__(self, t):
        self._t = t

    @property
    def child(self):
        mv, N, W, P, Q = self._t.tree.root_children()
        return OrderedDict((O.move_str(m), _Child(int(n), float(q))) for m, n, q in zip(mv, N, Q))


class _OracleMCTS:
    """MCTS_tree's surface (mcts.py) over the C oracle tree."""

    def __init__(self, net):
        self.tree, self.net, self.root = O.Tree(), net, _Root(self)
        self.sets = self.updates = self.searches = 0

    def _set_position(self, state, player, rr):
        self.tree.reload(O.from_state(state))
        self.sets += 1

    def main(self, state, player, rr, playouts):
        assert O.to_state(self.tree.root_board()) == state
        assert self.tree.search(0 if player == "w" else 1, rr, playouts, self.net) == 0
        self.searches += 1

    def Q(self, move):
        return self.root.child[move].Q

    def update_tree(self, act):
        self.tree.update(list(self.root.child.keys()).index(act))
        self.updates += 1


def _driver(net="hash_signed"):
    from cchess_zero_b200.selfplay import cchess_main

    class GB:
        state, current_player, restrict_round, round = O.START, "w", 0, 1

    made = []

    def make(options):
        d = cchess_main.__new__(cchess_main)        # the real get_action / check_end text over the stand-in tree
        d.game_borad, d.mcts = GB(), _OracleMCTS(net)
        d.playout_counts, d.exploration, d.temperature = options["playouts"], False, 1
        made.append(d)
        return d
    return make, made


def _session(eng, *lines):
    eng.out = io.StringIO()
    for ln in lines:
        assert eng.handle(ln)
    return eng.out.getvalue().splitlines()


def _bestmove(lines):
    bm = [ln for ln in lines if ln.startswith("bestmove")]
    assert len(bm) == 1
    return bm[0].split()[1]


def _top_moves(tree):
    """Moves sharing the highest visit count: at T = 1e-3 get_action's np.random.choice picks among exactly these."""
    mv, N, *_ = tree.root_children()
    return {O.move_str(m) for m, n in zip(mv, N) if n == N.max()}


def test_handshake_and_options():
    make, made = _driver()
    eng = ucci.UcciEngine(make, playouts=40)
    out = _session(eng, "ucci")
    assert out[0].startswith("id name") and out[-1] == "ucciok" and not made      # no engine yet: ucci must answer at once
    assert _session(eng, "isready") == ["readyok"] and len(made) == 1
    _session(eng, "setoption name playouts value 64")
    assert made[0].playout_counts == 64
    _session(eng, "setoption leaf_parallel 4")
    assert eng.options["leaf_parallel"] == 4 and eng._driver is None              # K is a construction-time property
    assert "unknown option" in _session(eng, "setoption name hash value 1")[0]
    assert "unknown command" in _session(eng, "xyzzy")[0]
    assert "error" in _session(eng, "position fen 9/9 w")[0]
    eng.out = io.StringIO()
    assert eng.handle("quit") is False and eng.out.getvalue() == "bye\n"


def test_go_plays_the_most_visited_move_and_reuses_the_tree():
    import contextlib
    np.random.seed(0)
    make, made = _driver("hash_signed")
    eng = ucci.UcciEngine(make, playouts=60)
    with contextlib.redirect_stdout(io.StringIO()):
        first = _bestmove(_session(eng, "position startpos", "go"))
    d = made[0]
    ref = O.Tree()
    assert ref.search(0, 0, 60, "hash_signed") == 0
    assert first in _top_moves(ref)                      # T = 1e-3: the visit arg-max
    assert d.mcts.sets == 1 and d.mcts.updates == 1 and d.game_borad.current_player == "b" and d.game_borad.round == 2
    # the GUI answers with a reply that the re-rooted tree already holds: no reset, one more update_tree
    mv, N, *_ = ref.root_children()
    ref.update([O.move_str(m) for m in mv].index(first))
    rmv, rN, *_ = ref.root_children()
    assert len(rmv) > 0
    reply = O.move_str(rmv[int(np.argmax(rN))])
    ref.update(int(np.argmax(rN)))
    with contextlib.redirect_stdout(io.StringIO()):
        second = _bestmove(_session(eng, "position startpos moves %s %s" % (first, reply), "go nodes 50"))
    assert d.mcts.sets == 1 and d.mcts.updates == 3       # reply + the engine's own second move
    b = O.from_state(O.START)
    rr = 0
    for m in (first, reply):
        b, cap = O.apply_move(b, O.move_from_str(m))
        rr = 0 if cap else rr + 1
    assert ref.search(0, rr, 50, "hash_signed") == 0
    assert second in _top_moves(ref)
    assert d.playout_counts == 50
    # a move list that departs from the tree resets it
    other = next(O.move_str(m) for m in O.legal_moves(O.from_state(O.START), 0) if O.move_str(m) != first)
    with contextlib.redirect_stdout(io.StringIO()):
        _bestmove(_session(eng, "position startpos moves %s" % other, "go nodes 20"))
    assert d.mcts.sets == 2
    st, _ = ucci.apply_move(O.START, other)
    probe = _session(eng, "probe")[0]
    assert probe.startswith("info string fen ") and ucci.fen_to_state(probe[len("info string fen "):])[:2] == (st, "b")


def test_fen_position_black_to_move_and_game_over():
    import contextlib
    np.random.seed(1)
    make, made = _driver("hash_pos")
    eng = ucci.UcciEngine(make, playouts=30)
    fen = "4k4/9/9/9/4p4/9/9/9/4R4/3K5 b - - 0 1"
    with contextlib.redirect_stdout(io.StringIO()):
        bm = _bestmove(_session(eng, "position fen " + fen, "go"))
    st, pl, _ = ucci.fen_to_state(fen)
    ref = O.Tree(O.from_state(st))
    assert ref.search(1, 0, 30, "hash_pos") == 0
    assert bm in _top_moves(ref)
    # king already captured: nobestmove
    with contextlib.redirect_stdout(io.StringIO()):
        out = _session(eng, "position fen 9/9/9/9/4p4/9/9/9/4R4/3K5 b", "go")
    assert out[-1] == "nobestmove" and "game over (w)" in out[0]
    # 60 quiet plies: tie
    with contextlib.redirect_stdout(io.StringIO()):
        out = _session(eng, "position fen 4k4/9/9/9/4p4/9/9/9/4R4/3K5 b - - 60 40", "go")
    assert out[-1] == "nobestmove" and "(t)" in out[0]

"""GPU twin of tests/test_ucci.py: the same protocol session through the real cchess_main / MCTS_tree / device engine, checked
against the oracle tree.  (Sorted last on purpose: it is the outermost layer.)"""
import contextlib
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ucci_session_on_the_device_engine(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from cchess_zero_b200 import ucci
    from cchess_zero_b200.selfplay import cchess_main
    from oracle import oracle as O
    from oracle.fakenets_np import FAKE_NETS

    class Net:
        def __init__(self, f):
            self.forward = f

    def make(options):
        return cchess_main(playout=options["playouts"], in_search_threads=1, network=Net(FAKE_NETS["hash_signed"]),
                           exploration=False, log_file=False)

    def top(tree):
        mv, N, *_ = tree.root_children()
        return {O.move_str(m) for m, n in zip(mv, N) if n == N.max()}

    def say(eng, *lines):
        eng.out = io.StringIO()
        with contextlib.redirect_stdout(io.StringIO()):
            for ln in lines:
                assert eng.handle(ln)
        return eng.out.getvalue().splitlines()

    np.random.seed(0)
    eng = ucci.UcciEngine(make, playouts=60)
    assert say(eng, "ucci")[-1] == "ucciok" and say(eng, "isready") == ["readyok"]
    first = [l for l in say(eng, "position startpos", "go") if l.startswith("bestmove")][0].split()[1]
    ref = O.Tree()
    assert ref.search(0, 0, 60, "hash_signed") == 0
    assert first in top(ref)
    mv, N, *_ = ref.root_children()
    ref.update([O.move_str(m) for m in mv].index(first))
    rmv, rN, *_ = ref.root_children()
    reply = O.move_str(rmv[int(np.argmax(rN))])
    ref.update(int(np.argmax(rN)))
    second = [l for l in say(eng, "position startpos moves %s %s" % (first, reply), "go nodes 50") if l.startswith("bestmove")][0].split()[1]
    b, rr = O.from_state(O.START), 0
    for m in (first, reply):
        b, cap = O.apply_move(b, O.move_from_str(m))
        rr = 0 if cap else rr + 1
    assert ref.search(0, rr, 50, "hash_signed") == 0          # subtree kept across both moves, like the device tree
    assert second in top(ref)
    # FEN position, black to move
    fen = "4k4/9/9/9/4p4/9/9/9/4R4/3K5 b - - 0 1"
    bm = [l for l in say(eng, "position fen " + fen, "go nodes 30") if l.startswith("bestmove")][0].split()[1]
    st, pl, _ = ucci.fen_to_state(fen)
    ref2 = O.Tree(O.from_state(st))
    assert ref2.search(1, 0, 30, "hash_signed") == 0
    assert bm in top(ref2)
    assert say(eng, "position fen 9/9/9/9/4p4/9/9/9/4R4/3K5 b", "go")[-1] == "nobestmove"

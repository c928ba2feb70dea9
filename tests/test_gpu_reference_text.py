"""The drop-in contract, tested literally: the reference's OWN `class cchess_main` text (main.py:1118-1554, unmodified,
sha256-checked against oracle/reference_manifest.json) is executed over cchess_zero_b200's GameBoard / MCTS_tree /
policy_value_network on the GPU, and must reproduce what the reference itself produced (tests/golden/).

The text comes from oracle/_ref/reference/main.py, staged by oracle/stage_reference.py (git-ignored, travels to the GPU box).
A name, attribute or call shape that the reference's text touches and the package lacks makes these tests fail."""
import contextlib
import hashlib
import io
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def _reference_main_py():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stage_reference as S
    staged = os.path.join(S.DST, "main.py")
    path = staged if os.path.isfile(staged) else os.path.join(S.SRC, "main.py")
    if not os.path.isfile(path):
        pytest.skip("reference text not staged (run `python oracle/stage_reference.py` where /root/reference exists)")
    assert S._sha(path) == S.manifest()["main.py"], "staged main.py is not the unmodified reference"
    return path


@pytest.fixture(scope="module")
def ref_cchess_main():
    from cchess_zero_b200.refapi import bind_reference_main
    cls, ns = bind_reference_main(_reference_main_py())
    return cls, ns


class _FakeNetwork:
    """policy_value_network stand-in with the deterministic evaluators the goldens were generated with."""

    def __init__(self, f):
        self.forward = f

    def save(self, step):
        pass


def _make(cls, ns, playout, net_fn, threads=1, exploration=True, human_color="b"):
    """cchess_main(...) through the reference's own constructor (it builds the package's network and tree), then the
    evaluator is swapped for the golden stand-in exactly as main.py:1144 wires it: MCTS_tree(state, forward, threads)."""
    # (cwd is a tmp dir: the constructor opens ./log_file.txt, main.py:1149, and the network makes ./models)
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(playout, 128, exploration, threads, "cpu", 1, 2, human_color)
    assert type(m.policy_value_netowrk).__name__ == "policy_value_network" and m.mcts.forward == m.policy_value_netowrk.forward
    m.policy_value_netowrk = _FakeNetwork(net_fn)
    m.mcts = ns["MCTS_tree"](m.game_borad.state, net_fn, threads)
    return m


def test_reference_selfplay_text_reproduces_reference_tuples(ref_cchess_main, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from oracle.fakenets_np import FAKE_NETS
    cls, ns = ref_cchess_main
    for g in load_golden("selfplay.json")["games"][:3]:
        m = _make(cls, ns, g["playouts"], FAKE_NETS[g["net"]])
        np.random.seed(g["seed"])
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            data, n = m.selfplay()                                  # main.py:1493-1554, the reference's text
        data = list(data)
        assert n == g["n"]
        assert [d[0] for d in data] == g["states"]
        assert [float(d[2]) for d in data] == g["z"]
        assert sha(np.asarray([d[1] for d in data], dtype=np.float64).tobytes()) == g["sha_pi"]
        m.log_file.close()


@pytest.mark.parametrize("i", range(6))
def test_reference_selfplay_text_at_its_default_search_threads_reproduces_reference_tuples(ref_cchess_main, tmp_path, monkeypatch, i):
    """The reference's default configuration end to end: its own selfplay() text with search_threads = 16 (and 8, 4) over the package's
    MCTS_tree must give the tuples the unmodified reference gives when its coroutines run on the canonical deterministic schedule
    (oracle/gen_golden_k16_selfplay.py; whole games, including king captures right below the root)."""
    monkeypatch.chdir(tmp_path)
    from oracle.fakenets_np import FAKE_NETS
    cls, ns = ref_cchess_main
    g = load_golden("selfplay_k16.json")["games"][i]
    m = _make(cls, ns, g["playouts"], FAKE_NETS[g["net"]], threads=g["search_threads"])
    np.random.seed(g["seed"])
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        data, n = m.selfplay()
    data = list(data)
    assert n == g["n"]
    assert [d[0] for d in data] == g["states"]
    assert [float(d[2]) for d in data] == g["z"]
    assert sha(np.asarray([d[1] for d in data], dtype=np.float64).tobytes()) == g["sha_pi"]
    m.log_file.close()


def test_reference_play_text_reproduces_reference_moves(ref_cchess_main, tmp_path, monkeypatch):
    """select_move / get_hint / human_move / check_end of the reference's text (main.py:1278-1491) over the package."""
    from oracle.fakenets_np import FAKE_NETS
    monkeypatch.chdir(tmp_path)
    cls, ns = ref_cchess_main
    for sc in load_golden("play.json")["scripts"]:
        m = _make(cls, ns, sc["playouts"], FAKE_NETS[sc["net"]], exploration=False, human_color=sc["human_color"])
        if sc["seed"] is not None:
            np.random.seed(sc["seed"])
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            for i, st in enumerate(sc["steps"]):
                op = st["op"]
                if op in ("select_move_mcts", "select_move_net"):
                    mv, wr = m.select_move("mcts" if op.endswith("mcts") else "net")
                    assert [int(x) for x in mv] == st["move"], (sc["human_color"], i)
                    assert float(wr).hex() == st["win_rate"], (sc["human_color"], i)
                    assert m.game_borad.state == st["state"]
                elif op in ("get_hint_mcts", "get_hint_net"):
                    hint = m.get_hint("mcts" if op.endswith("mcts") else "net", op.endswith("mcts"), lambda: None)
                    assert [[a, float(p).hex()] for a, p in hint] == st["hint"], (sc["human_color"], i, op)
                elif op == "human_move_mcts":
                    wr = m.human_move(tuple(st["coord"]), "mcts")
                    assert float(wr).hex() == st["win_rate"], (sc["human_color"], i)
                    assert m.game_borad.state == st["state"] and m.game_borad.current_player == st["player"]
                    assert m.game_borad.restrict_round == st["rr"]
                elif op == "check_end":
                    ended_who = m.check_end()
                    assert bool(ended_who[0]) == st["ended"] and ended_who[1] == st["who"]
        m.log_file.close()


def test_reference_run_loop_trains_the_package_network(ref_cchess_main, tmp_path, monkeypatch):
    """The reference's run() (main.py:1224-1248): selfplay() -> state_to_positions -> data_buffer -> policy_update()
    (random.sample, mcts.forward on a list of positions, 5 x train_step + KL, save, lr_multiplier, log line), entirely
    the reference's text, on the package's real network.  Left through the reference's own exit (KeyboardInterrupt ->
    close log + save).  Also the ADVICE r1 regression: the search after training must see the trained weights."""
    import random
    cls, ns = ref_cchess_main
    monkeypatch.chdir(tmp_path)
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(24, 16, True, 16, "cpu", 1, 2, "b")               # playout 24, batch 16, search_threads 16 (the default)
    pv = m.policy_value_netowrk
    np.random.seed(5)
    random.seed(5)
    x0 = m.mcts.generate_inputs(m.game_borad.state, "w")
    lo_before, _ = pv.forward(np.expand_dims(x0, 0))

    # searched root priors before training (the tree's own evaluator path: native plan + CUDA graph)
    m.mcts.main(m.game_borad.state, "w", 0, 8)
    p_before = np.array([n.P for n in m.mcts.root.child.values()])
    m.mcts.reload()

    calls = {"n": 0, "updates": 0}
    real_selfplay, real_update = m.selfplay, m.policy_update

    def selfplay():
        calls["n"] += 1
        if calls["updates"] >= 1:
            raise KeyboardInterrupt                                   # the reference's way out of run()
        return real_selfplay()

    def policy_update():
        calls["updates"] += 1
        return real_update()

    m.selfplay, m.policy_update = selfplay, policy_update
    out = io.StringIO()
    with contextlib.redirect_stdout(out), np.errstate(all="ignore"):
        m.run()
    assert calls["updates"] == 1 and pv.global_step >= 1 and m.global_step == pv.global_step
    assert len(m.data_buffer) > 16 and m.data_buffer[0][0].shape == (9, 10, 14) and m.data_buffer[0][1].shape == (2086,)
    log = open(os.path.join(str(tmp_path), "log_file.txt")).read()
    assert log.startswith("kl:") and "lr_multiplier" in log and m.log_file.closed
    assert os.path.isfile(os.path.join(str(tmp_path), "models", "checkpoint"))
    assert "train using time" in out.getvalue() and "batch i:1" in out.getvalue()

    lo_after, _ = pv.forward(np.expand_dims(x0, 0))
    assert not np.array_equal(lo_before, lo_after)                    # the network was trained
    m.game_borad.reload()
    m.mcts.reload()
    m.mcts.main(m.game_borad.state, "w", 0, 8)                        # same tree object, its plan / CUDA graph were built before training
    p_after = np.array([n.P for n in m.mcts.root.child.values()])
    assert not np.array_equal(p_before, p_after), "search still evaluates with the pre-training weights"
    # and they are the CURRENT weights: priors = legal logits / their sum (expand, main.py:176-187)
    moves = ns["GameBoard"].get_legal_moves(m.game_borad.state, "w")
    lg = lo_after.flatten()[[ns["label2i"][a] for a in moves]].astype(np.float64)
    want = lg / (1e-8 + lg.sum())
    assert np.allclose(p_after, want, rtol=2e-2, atol=2e-3)

"""Batched self-play driver (many games in lock-step on one GPU) and the reference-shaped
`cchess_main` facade (main.py:1118-1554).

Move choice stays on the host in numpy, exactly as the reference does it (get_action,
main.py:1332-1358): pi = softmax(log(visits)/T) in float64 and
np.random.choice(p = 0.75*pi + 0.25*Dirichlet(0.3)) on a legacy MT19937 RandomState -- one
RandomState per game slot stands in for the reference's global np.random.  The device
produces the integer visit counts; everything before them (select / expand / backup / encode /
move generation / re-rooting) runs in csrc/cz_engine.cu."""
import numpy as np
import torch

from . import rules
from ._lib import MAXCHILD, NLABEL, EngineError
from .engine import Engine


def _flip_board(b):
    """try_flip (main.py:560-574): reverse the rows, swap the colours; files are not mirrored."""
    f = b.reshape(10, 9)[::-1].copy()
    red, blk = (f >= 1) & (f <= 7), f >= 8
    f[red] += 7
    f[blk] -= 7
    return f.reshape(90)


def _flip_move_label_index(mv):
    """label index of the rank-mirrored move (flipped_uci_labels, main.py:23-27 / 1507-1512)."""
    s, d = int(mv) & 127, (int(mv) >> 7) & 127
    s = (9 - s // 9) * 9 + s % 9
    d = (9 - d // 9) * 9 + d % 9
    return rules.label2i[rules.move_to_label(s | (d << 7))]


class GameRecord:
    """(s, pi, z) tuples of one finished game in the reference's format (selfplay, main.py:1493-1554)."""

    def __init__(self):
        self.states, self.pi_idx, self.pi_val, self.players, self.actions, self.visits = [], [], [], [], [], []
        self.z = None
        self.winner = None

    def __len__(self):
        return len(self.states)

    def dense_pi(self):
        out = np.zeros((len(self.states), NLABEL))
        for i, (ix, v) in enumerate(zip(self.pi_idx, self.pi_val)):
            out[i, ix] = v
        return out

    def tuples(self):
        return zip(self.states, self.dense_pi(), self.z)


class SelfPlay:
    """n_games concurrent self-play games on one engine.

    forward_dev(nn_in) -> None must write logits f32 [B,2086] / value f32 [B] into the tensors handed to
    it at construction time via `bind(nn_in, logits, value)`; or pass a callable returning (logits, value)
    device tensors (they are copied into the static buffers)."""

    def __init__(self, n_games, forward, playouts, seeds=None, exploration=True, temperature=1,
                 nn_dtype=torch.float32, arena_words=0, auto_reset=True, device=None, keep_records=True):
        self.engine = Engine(n_games, arena_words, device)
        self.B = n_games
        dev = torch.device("cuda", self.engine.device)
        self.nn_in = torch.zeros((n_games, 9, 10, 14), dtype=nn_dtype, device=dev)
        self.logits = torch.zeros((n_games, NLABEL), dtype=torch.float32, device=dev)
        self.value = torch.zeros((n_games,), dtype=torch.float32, device=dev)
        self.forward = forward
        self.playouts = np.broadcast_to(np.asarray(playouts, dtype=np.int64), (n_games,)).copy()
        seeds = range(n_games) if seeds is None else seeds
        self.rs = [np.random.RandomState(int(s)) for s in seeds]
        self.exploration = exploration
        self.temperature = temperature
        self.auto_reset = auto_reset
        self.keep_records = keep_records
        self.records = [GameRecord() for _ in range(n_games)]
        self.boards = np.tile(rules.state_to_board(rules.START_STATE), (n_games, 1))
        self.sides = np.zeros(n_games, dtype=np.uint8)
        self.live = np.ones(n_games, dtype=bool)
        self.finished = []
        self.plies = 0
        self.waves = 0
        self.graph = None
        rules._init_tables()

    # -- evaluation step ---------------------------------------------------------------------
    def _eval(self, nn_in):
        out = self.forward(nn_in)
        if out is not None:
            lo, v = out
            self.logits.copy_(lo.reshape(self.B, NLABEL))
            self.value.copy_(v.reshape(self.B))

    def capture_graph(self, warmup=3):
        """Capture (wave kernel -> network) into one CUDA graph; the search loop then replays it."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eval(self.nn_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.engine.wave(self.nn_in, self.logits, self.value)
            self._eval(self.nn_in)
        self.graph = g

    def search(self):
        """MCTS_tree.main for every live game: `playouts[g]` playouts each."""
        e = self.engine
        for p in np.unique(self.playouts[self.live]):
            e.begin_search(int(p), (self.live & (self.playouts == p)).astype(np.uint8))
        pmax = int(self.playouts[self.live].max()) if self.live.any() else 0
        waves = 0
        while True:
            if self.graph is not None:
                self.graph.replay()
                e.launches += 1          # the captured k_wave launch
            else:
                e.wave(self.nn_in, self.logits, self.value)
            waves += 1
            if waves > pmax and e.unfinished() == 0:
                break
            if self.graph is None:
                self._eval(self.nn_in)
            if waves > 4 * pmax + 64:
                e.raise_on_error()
                raise EngineError("search did not converge")
        self.waves += waves
        return waves

    # -- one ply for every live game ------------------------------------------------------------
    def step(self):
        e = self.engine
        self.search()
        rc = e.root_children(want_wpq=True)
        choice = np.full(self.B, -1, dtype=np.int32)
        win_rate = np.zeros(self.B, dtype=np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            for g in np.nonzero(self.live)[0]:
                n = int(rc["n"][g])
                if n <= 0:
                    e.raise_on_error()
                    raise EngineError("game %d has no root children" % g)
                visits = rc["visits"][g, :n].astype(np.int64)
                probs = rules.softmax(1.0 / self.temperature * np.log(visits))        # main.py:1341
                rs = self.rs[g]
                if self.exploration:                                                     # main.py:1345-1348
                    p = 0.75 * probs + 0.25 * rs.dirichlet(0.3 * np.ones(len(probs)))
                else:
                    p = probs
                idx = int(rs.choice(n, p=p))
                choice[g] = idx
                win_rate[g] = rc["q"][g, idx]                                            # mcts.Q(act), main.py:1350
                if self.keep_records:
                    rec = self.records[g]
                    mv = rc["moves"][g, :n]
                    black = self.sides[g] == 1
                    sb = _flip_board(self.boards[g]) if black else self.boards[g]
                    rec.states.append(rules.board_to_state(sb))                           # main.py:1504-1505
                    if black:
                        ix = np.fromiter((_flip_move_label_index(m) for m in mv), dtype=np.int64, count=n)
                    else:
                        ix = np.fromiter((rules.label2i[rules.move_to_label(m)] for m in mv), dtype=np.int64, count=n)
                    rec.pi_idx.append(ix)
                    rec.pi_val.append(probs)
                    rec.players.append(int(self.sides[g]))
                    rec.actions.append(rules.move_to_label(mv[idx]))
                    rec.visits.append(visits)
                else:
                    self.records[g].players.append(int(self.sides[g]))
        e.play(choice)
        st = e.status(boards=True)
        self.boards, self.sides = st["boards"], st["side"]
        self.plies += int(self.live.sum())
        done_now = []
        for g in np.nonzero(self.live & (st["terminal"] != 0))[0]:
            rec = self.records[g]
            players = np.asarray(rec.players)
            if st["terminal"][g] == 1:                                                   # main.py:1532-1541
                w = int(st["winner"][g])
                rec.z = np.where(players == w, 1.0, -1.0)
                rec.winner = "w" if w == 0 else "b"
            else:                                                                        # main.py:1542-1545
                rec.z = np.zeros(len(players))
                rec.winner = "t"
            done_now.append((int(g), rec))
            self.finished.append((int(g), rec))
            self.records[g] = GameRecord()
        if done_now:
            mask = np.zeros(self.B, dtype=np.uint8)
            mask[[g for g, _ in done_now]] = 1
            if self.auto_reset:
                e.reset(mask)                                                            # GameBoard.reload + mcts.reload
                st = e.status(boards=True)
                self.boards, self.sides = st["boards"], st["side"]
            else:
                self.live[[g for g, _ in done_now]] = False
        return dict(choice=choice, win_rate=win_rate, finished=done_now, status=st)

    def play_games(self, max_plies=100000):
        """Every slot plays ONE game to the end (auto_reset must be False)."""
        assert not self.auto_reset
        n = 0
        while self.live.any() and n < max_plies:
            self.step()
            n += 1
        self.engine.raise_on_error()
        return sorted(self.finished, key=lambda t: t[0])

"""TEST INFRASTRUCTURE ONLY -- a DETERMINISTIC event loop for the reference's search_threads > 1 coroutines.

The reference's MCTS_tree.main (main.py:473-493) runs `playouts` tree_search coroutines + one prediction_worker on uvloop.  Its
visit counts then depend on the interleaving of: FIFO ready-queue batches (uvloop `_on_idle`: the callbacks present at the start of
an iteration run, new ones wait for the next iteration), `asyncio.sleep(1e-4)` spins (uvloop rounds the delay to 0 ms = call_soon,
so a spin is two ready-queue hops: the sleep future's set_result, then the task wake-up) and prediction_worker's 1 ms libuv timer
(millisecond clock sampled once per loop iteration; the timer callback runs before that iteration's batch, so the worker itself
runs at the END of the batch) -- i.e. on wall-clock time.

DetLoop reproduces exactly that machinery (same batch rule, same hop counts, same timer placement) but replaces the wall clock by a
VIRTUAL clock driven by a cost model, so that the unmodified coroutines of the reference run under a schedule that is a pure function
of the position.  It is used (a) to pin down which schedule the reference follows when it is not timing-sensitive (tests compare it
with the real uvloop runs recorded in tests/golden/k16_stats.json.gz) and (b) as the executable specification of the engine's
`search_threads=16` mode.

Timer placement, as on the real system: uvloop samples the clock (uv_update_time) when it ARMS a timer, libuv keeps time in whole
milliseconds (due = floor(now) + 1 ms) and checks timers once per iteration.  prediction_worker arms its 1 ms sleep as the last
callback of a batch, the next iteration starts microseconds later, so the timer is practically never due there; after one more batch
of real tree work (a Python expansion costs ~2 ms) it always is.

policy "busy": a timer armed during iteration n fires at the start of iteration n + 2 -- the canonical schedule, a pure function of
               the position; it is the one the engine implements.  (The real system deviates only when the batch in between is a
               pure spin batch shorter than the distance to the next millisecond boundary: the timing-dependent cases.)
policy "cost": virtual clock driven by per-callback costs (spin hop 0.03 ms, select 0.1 ms, expand 2 ms, forward 3 ms) with libuv's
               floor-to-millisecond arithmetic: approximates the real phase of the timer against pure-spin iterations."""
import asyncio
import collections
import heapq


class DetLoop(asyncio.SelectorEventLoop):
    def __init__(self, policy="busy", costs=None):
        super().__init__()
        self.policy = policy
        self.vt_ms = 0.0                      # virtual wall clock
        self.uv_now = 0                       # libuv's cached millisecond clock (sampled once per iteration)
        self.timers = []                      # (due_ms, seq, handle)
        self._seq = 0
        self.costs = dict(hop=0.03, select=0.1, expand=2.0, forward=3.0)
        if costs:
            self.costs.update(costs)
        self.iterations = 0

    # ---- clock -------------------------------------------------------------------------------
    def time(self):
        return self.vt_ms / 1000.0

    def charge(self, what):
        self.vt_ms += self.costs[what]

    # ---- scheduling ----------------------------------------------------------------------------
    def call_later(self, delay, callback, *args, context=None):
        when = round(max(delay, 0) * 1000)                       # uvloop loop.pyx: when = <uint64_t>round(delay * 1000)
        if when == 0:
            return self.call_soon(callback, *args, context=context)
        h = asyncio.Handle(callback, args, self, context)
        self._seq += 1
        if self.policy == "busy":
            due = self.iterations + 1 + when                      # armed in iteration n, 1 ms: due at the start of iteration n + 2
        else:
            due = int(self.vt_ms) + when                          # uv_update_time at arming, then loop->time + timeout (whole ms)
        heapq.heappush(self.timers, (due, self._seq, h))
        return h

    def call_at(self, when, callback, *args, context=None):
        return self.call_later(when - self.time(), callback, *args, context=context)

    def _run_once(self):
        self.iterations += 1
        ready = self._ready
        if not ready and not self.timers:
            raise RuntimeError("deterministic loop is idle with nothing scheduled (deadlock)")
        if self.policy == "busy":
            if not ready:                                          # poll would block until the next timer
                self.iterations = max(self.iterations, self.timers[0][0])
            self.uv_now = self.iterations
        else:
            if not ready:
                self.vt_ms = max(self.vt_ms, float(self.timers[0][0]))
            self.uv_now = int(self.vt_ms)                          # uv__update_time (whole milliseconds)
        while self.timers and self.timers[0][0] <= self.uv_now:    # uv__run_timers: callbacks run NOW, what they schedule joins this batch
            _, _, h = heapq.heappop(self.timers)
            if not h._cancelled:
                h._run()
        ntodo = len(ready)                                         # uvloop _on_idle
        for _ in range(ntodo):
            h = ready.popleft()
            if h._cancelled:
                continue
            self.charge("hop")
            h._run()


def run_reference_search(ref, tree, state, player, rr, playouts, loop):
    """MCTS_tree.main (main.py:473-493) with the reference's own coroutines, on `loop` instead of uvloop."""
    import numpy as np
    node = tree.root
    tree.loop = loop
    if not tree.is_expanded(node):
        positions = tree.generate_inputs(node.state, player)
        positions = np.expand_dims(positions, 0)
        action_probs, value = tree.forward(positions)
        if tree.is_black_turn(player):
            action_probs = ref.cchess_main.flip_policy(action_probs)
        moves = ref.GameBoard.get_legal_moves(node.state, player)
        node.expand(moves, action_probs)
        tree.expanded.add(node)
    coros = [tree.tree_search(node, player, rr) for _ in range(playouts)]
    coros.append(tree.prediction_worker())
    loop.run_until_complete(asyncio.gather(*coros))


def make_tree(H, forward, K, state, policy="busy", costs=None):
    """A reference MCTS_tree whose asyncio primitives are bound to a fresh DetLoop, with cost hooks for the virtual clock."""
    ref = H.load_reference()
    loop = DetLoop(policy, costs)
    asyncio.set_event_loop(loop)

    def fwd(x):
        loop.charge("forward")
        return forward(x)
    t = ref.MCTS_tree_py312(state, fwd, K)
    t.loop = loop
    return ref, t, loop


class charged:
    """Context manager: leaf_node.expand / select_new advance the virtual clock of `loop` while active."""

    def __init__(self, ref, loop):
        self.ref, self.loop = ref, loop

    def __enter__(self):
        ln, loop = self.ref.leaf_node, self.loop
        self._e, self._s = ln.expand, ln.select_new

        def expand(node, moves, probs):
            loop.charge("expand")
            return self._e(node, moves, probs)

        def select_new(node, c):
            loop.charge("select")
            return self._s(node, c)
        ln.expand, ln.select_new = expand, select_new

    def __exit__(self, *a):
        self.ref.leaf_node.expand, self.ref.leaf_node.select_new = self._e, self._s

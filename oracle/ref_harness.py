"""TEST INFRASTRUCTURE ONLY -- never imported by the product (cchess_zero_b200/).

Loads the UNMODIFIED reference (chengstone/cchess-zero, /root/reference/main.py)
in this container so that golden vectors can be generated from the reference's
own code (see oracle/gen_golden.py).  /root/reference does not exist on the GPU
box, so nothing under tests/ -m gpu, smoke() or bench.py may import this file.

Shims; the reference source text is not edited:
  1. an empty `tensorflow` module object is placed in sys.modules, because
     main.py:8 and policy_value_network.py:2 import it at module top but only
     the network constructors use it;
  2. asyncio.set_event_loop(new_event_loop()) before MCTS_tree() because
     main.py:252 calls asyncio.get_event_loop() (no implicit loop on py3.12);
  3. MCTS_tree.tree_search (main.py:337-348) is overridden with the identical
     body using `async with self.sem:` -- `with await self.sem:` (main.py:342)
     was removed from Python in 3.9;
  4. cchess_main is built with __new__ + the attribute assignments of
     main.py:1121-1150, skipping the TF network ctor (main.py:1142) and the
     log-file open (main.py:1149); `forward` is any callable with the
     policy_value_network.forward signature (policy_value_network.py:202-214).
"""
import asyncio
import contextlib
import io
import os
import sys
import types
from collections import deque

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import stage_reference as _stage  # noqa: E402

# the live reference tree in the build container; on the GPU box the byte-identical copy staged by oracle/stage_reference.py
REF_DIR = os.environ.get("CCHESS_REFERENCE_DIR") or _stage.staged_dir() or "/root/reference"

_ref = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, "main.py"))


def load_reference():
    """Import /root/reference/main.py as module `main` (shims 1-3)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference not present at %s" % REF_DIR)
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")  # shim 1
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    asyncio.set_event_loop(asyncio.new_event_loop())  # shim 2
    import main as ref  # noqa: E402  (the reference's main.py)

    class MCTS_tree_py312(ref.MCTS_tree):
        async def tree_search(self, node, current_player, restrict_round):  # shim 3
            self.running_simulation_num += 1
            async with self.sem:
                value = await self.start_tree_search(node, current_player, restrict_round)
                self.running_simulation_num -= 1
                return value

    ref.MCTS_tree_py312 = MCTS_tree_py312
    _ref = ref
    return ref


def make_mcts(forward, search_threads=1, state=None):
    ref = load_reference()
    asyncio.set_event_loop(asyncio.new_event_loop())
    st = state or "RNBAKABNR/9/1C5C1/P1P1P1P1P/9/9/p1p1p1p1p/1c5c1/9/rnbakabnr"
    return ref.MCTS_tree_py312(st, forward, search_threads)


def make_cchess_main(forward, playout, search_threads=1, exploration=True, human_color="b"):
    """shim 4: cchess_main without the TF ctor; attribute list = main.py:1121-1150."""
    ref = load_reference()
    m = ref.cchess_main.__new__(ref.cchess_main)
    m.epochs = 5
    m.playout_counts = playout
    m.temperature = 1
    m.batch_size = 128
    m.game_batch = 400
    m.top_steps = 30
    m.top_temperature = 1
    m.eta = 0.03
    m.learning_rate = 0.001
    m.lr_multiplier = 1.0
    m.buffer_size = 10000
    m.data_buffer = deque(maxlen=m.buffer_size)
    m.game_borad = ref.GameBoard()
    m.policy_value_netowrk = None
    m.search_threads = search_threads
    m.mcts = make_mcts(forward, search_threads, m.game_borad.state)
    m.exploration = exploration
    m.resign_threshold = -0.8
    m.global_step = 0
    m.kl_targ = 0.025
    m.log_file = None
    m.human_color = human_color
    return m


@contextlib.contextmanager
def quiet():
    """The reference prints on every game end / GUI capture (main.py:1541, Che.py:45)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


from fakenets_np import FAKE_NETS, NET_IDS, f32_bits, fake_forward_hash, fake_forward_mod17  # noqa: E402,F401


def tree_signature(node, ref=None):
    """Flat DFS listing (label_index, N, W bits, P bits, Q bits, n_children) in child order."""
    ref = ref or load_reference()
    out = []

    def rec(n):
        for a, c in n.child.items():
            out.append((ref.label2i[a], int(c.N), f32_bits(c.W), f32_bits(c.P), f32_bits(c.Q), len(c.child)))
            if c.child:
                rec(c)

    rec(node)
    return out

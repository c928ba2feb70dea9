"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): rules, waves in every dtype, re-root."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from cchess_zero_b200 import rules
from cchess_zero_b200.fakenet import FakeNet
from cchess_zero_b200.selfplay import SelfPlay
from cchess_zero_b200.net import policy_value_network

rules._init_tables()
b = rules.state_to_board(rules.START_STATE)
mv, cnt = rules.legal_moves_batch(np.stack([b] * 5), [0, 1, 0, 1, 0])
enc = rules.encode_batch(np.stack([b] * 5), [0, 1, 0, 1, 0])
nb, cap = rules.apply_moves_batch(np.stack([b] * 2), mv[0, :2])
sp = SelfPlay(8, FakeNet("hash_pos"), 24, seeds=range(8), arena_words=1 << 16, auto_reset=True)
for _ in range(4):
    sp.step()
sp.engine.raise_on_error()
pv = policy_value_network(res_block_nums=1)
sp2 = SelfPlay(8, None, 16, seeds=range(8), arena_words=1 << 16, plan=pv.native_plan(8))
for _ in range(3):
    sp2.step()
sp2.engine.raise_on_error()
print("sanitize run ok", cnt[:2], enc.sum(), sp.plies, sp2.plies)

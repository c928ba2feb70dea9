import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from oracle import oracle as O
from cchess_zero_b200 import rules as R
from test_gpu_parity import _random_positions
boards, sides = _random_positions(O, 400, 123)
mv, cnt = R.legal_moves_batch(boards, sides)
bad = 0
for i in range(len(boards)):
    om = O.legal_moves(boards[i], int(sides[i]))
    if cnt[i] != len(om) or not np.array_equal(mv[i,:cnt[i]], om):
        bad += 1
        if bad <= 5:
            print(i, O.to_state(boards[i]), sides[i])
            print(' gpu', [R.move_to_label(m) for m in mv[i,:cnt[i]]])
            print(' ora', [O.move_str(m) for m in om])
print('bad', bad, 'of', len(boards))
# single
mv1, cnt1 = R.legal_moves_batch(boards[5899:5900], sides[5899:5900])
print('single', cnt1, [R.move_to_label(m) for m in mv1[0,:cnt1[0]]])

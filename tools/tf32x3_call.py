"""A few evaluations of 1024 positions in the tf32x3 mode (for ncu captures of k_epilogue_split and the library convolutions it feeds)."""
import contextlib, io, sys
sys.path.insert(0, '.')
import torch
from cchess_zero_b200.net import policy_value_network
with contextlib.redirect_stdout(io.StringIO()):
    pv = policy_value_network(res_block_nums=7, precision="tf32x3")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.zeros((B, 9, 10, 14), device="cuda")
x[:, :, :, 0] = 1.0
plan = pv.plan()
for _ in range(4):
    lo, v = plan(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    plan(x)
e1.record(); torch.cuda.synchronize()
print("tf32x3 evaluation of %d positions: %.1f us" % (B, e0.elapsed_time(e1) * 100))

"""Debug aid: per-layer clock64 timeline of the cluster trunk (CCHESS_TOWER_TRACE=1 python tools/tower_trace.py [cluster] [positions])."""
import contextlib, io, os, sys
sys.path.insert(0, '.')
os.environ.setdefault("CCHESS_TOWER_TRACE", "1")
import torch
from cchess_zero_b200.net import policy_value_network
cl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with contextlib.redirect_stdout(io.StringIO()):
    pv = policy_value_network(res_block_nums=7)
plan = pv.small_plan(B, cl)
boards = torch.zeros((B, 96), dtype=torch.uint8, device="cuda")
boards[:, :90] = torch.randint(0, 15, (B, 90), dtype=torch.uint8, device="cuda") * (torch.rand((B, 90), device="cuda") < 0.3)
lo = torch.zeros((B, 2086), device="cuda"); vo = torch.zeros((B,), device="cuda")
for _ in range(25):
    plan(boards, lo, vo)
torch.cuda.synchronize()

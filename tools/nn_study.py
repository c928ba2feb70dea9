"""Precision / speed study of the policy-value net (run on the GPU box): max abs error of each
inference precision against an fp64 CPU evaluation of the same weights, and ms per batch."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from cchess_zero_b200.net import PolicyValueNet, InferencePlan
from oracle import oracle as O

torch.manual_seed(0)
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 7
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
net = PolicyValueNet(blocks).eval()
# realistic inputs: encodes of random-play positions
rng = np.random.RandomState(0)
xs = []
b, side = O.from_state(O.START), 0
while len(xs) < 256:
    xs.append(O.encode(b, side))
    mv = O.legal_moves(b, side)
    b, cap = O.apply_move(b, mv[rng.randint(len(mv))]); side ^= 1
    if cap in (1, 8) or len(mv) == 0:
        b, side = O.from_state(O.START), 0
x = torch.from_numpy(np.stack(xs))
with torch.no_grad():
    ref_l, ref_v = net.double()(x.double())
net.float()
netc = net.cuda().to(memory_format=torch.channels_last)
out = {}
for prec in ("fp32", "tf32", "bf16", "fp16"):
    plan = InferencePlan(netc, prec)
    xin = x.cuda().to(plan.dtype)
    l, v = plan(xin)
    el = (l.double().cpu() - ref_l).abs().max().item(); ev = (v.double().cpu().reshape(-1) - ref_v.reshape(-1)).abs().max().item()
    xb = xin.repeat((B + 255) // 256, 1, 1, 1)[:B].contiguous()
    lo = torch.zeros(B, 2086, device='cuda'); vo = torch.zeros(B, device='cuda')
    for _ in range(3): plan(xb, lo, vo)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan(xb, lo, vo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flops = {7: 375.4e6, 19: 1012.4e6}.get(blocks, 0) * B
    out[prec] = dict(fused=plan.fused, max_abs_err_logits=el, max_abs_err_value=ev, logits_absmax=ref_l.abs().max().item(),
                     ms_per_batch=ms, evals_per_s=B / ms * 1e3, tflops=flops / ms / 1e9)
    print(prec, json.dumps(out[prec]), flush=True)

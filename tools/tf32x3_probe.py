"""Probe (GPU): where the residual error of the 3xTF32 convolution comes from, and what each arrangement costs.
Compares, for one 3x3 128->128 convolution over B boards, against an fp64 evaluation:
  tf32            one TF32 product
  cat3            { hi | lo | hi } x { hi | hi | lo }, one convolution, K = 3456           (SplitTf32Plan as first written)
  hh+small        conv(hi, hi) with the small terms conv({ lo | hi }, { hi | lo }) added through the fused epilogue
  hh/G+small      the hi*hi term in G input-channel groups (shorter accumulation chains), chained through the epilogue
  fp32            cuDNN fp32
and a matmul K-sweep to see how the tensor-core accumulation error grows with the chain length."""
import json, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, ".")
from cchess_zero_b200.net import tf32_hi

torch.manual_seed(0)
dev = "cuda"
cl = lambda t: t.contiguous(memory_format=torch.channels_last)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out = {}
for B in (96, 1024):
    x = cl(torch.relu(torch.randn(B, 128, 9, 10, device=dev)) * 1.5)
    w = cl(torch.randn(128, 128, 3, 3, device=dev) * 0.03)
    b = torch.randn(128, device=dev) * 0.1
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
    scale = ref.abs().max().item()
    xh, wh = tf32_hi(x), tf32_hi(w)
    xl, wl = x - xh, w - wh
    r = {}
    torch.backends.cudnn.allow_tf32 = True
    f = lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
    r["tf32"] = ((f().double() - ref).abs().max().item() / scale, timeit(f))
    x3, w3 = cl(torch.cat([xh, xl, xh], 1)), cl(torch.cat([wh, wh, wl], 1))
    f = lambda: torch.cudnn_convolution_relu(x3, w3, b, (1, 1), (1, 1), (1, 1), 1)
    r["cat3"] = ((f().double() - ref).abs().max().item() / scale, timeit(f))
    x2, w2 = cl(torch.cat([xl, xh], 1)), cl(torch.cat([wh, wl], 1))
    xhc, whc = cl(xh), cl(wh)
    def hh_small():
        s = F.conv2d(x2, w2, None, padding=1)
        return torch.cudnn_convolution_add_relu(xhc, whc, s, 1.0, b, (1, 1), (1, 1), (1, 1), 1)
    r["hh+small"] = ((hh_small().double() - ref).abs().max().item() / scale, timeit(hh_small))
    for G in (2, 4):
        c = 128 // G
        xs = [cl(xh[:, i * c:(i + 1) * c]) for i in range(G)]
        ws = [cl(wh[:, i * c:(i + 1) * c]) for i in range(G)]
        zb = torch.zeros(128, device=dev)
        def grouped():
            s = F.conv2d(x2, w2, None, padding=1)
            for i in range(G - 1):
                s = F.conv2d(xs[i], ws[i], None, padding=1).add_(s)
            return torch.cudnn_convolution_add_relu(xs[G - 1], ws[G - 1], s, 1.0, b, (1, 1), (1, 1), (1, 1), 1)
        r["hh/%d+small" % G] = ((grouped().double() - ref).abs().max().item() / scale, timeit(grouped))
    torch.backends.cudnn.allow_tf32 = False
    f = lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
    r["fp32"] = ((f().double() - ref).abs().max().item() / scale, timeit(f, 5))
    torch.backends.cudnn.allow_tf32 = True
    out["conv_B%d" % B] = {k: dict(rel_err=v[0], us=v[1]) for k, v in r.items()}

# matmul K sweep: operands exactly representable in TF32, positive (worst case for a truncating accumulator)
mm = {}
for K in (64, 256, 1024, 4096):
    a = tf32_hi(torch.rand(2048, K, device=dev) + 0.5)
    bm = tf32_hi(torch.rand(K, 256, device=dev) + 0.5)
    ref = a.double() @ bm.double()
    torch.backends.cuda.matmul.allow_tf32 = True
    e_tc = ((a @ bm).double() - ref)
    torch.backends.cuda.matmul.allow_tf32 = False
    e_fp = ((a @ bm).double() - ref)
    mm[K] = dict(tf32_mean_rel=(e_tc / ref).mean().item(), tf32_maxabs_rel=(e_tc / ref).abs().max().item(),
                 fp32_mean_rel=(e_fp / ref).mean().item(), fp32_maxabs_rel=(e_fp / ref).abs().max().item())
out["matmul_exact_tf32_operands"] = mm
print(json.dumps(out, indent=1))

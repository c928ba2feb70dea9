"""Where the per-ply host phase goes (everything in SelfPlay.step() outside search()): wall-clock per part, 1024 x 1200, 7 blocks.
python tools/host_phase.py [plies]"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    from cchess_zero_b200 import _lib
    from cchess_zero_b200.net import policy_value_network
    from cchess_zero_b200.selfplay import SelfPlay
    plies = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
    B, P = 1024, 1200
    with contextlib.redirect_stdout(io.StringIO()):
        pv = policy_value_network(7, precision="fp16", device=0, seed=0)
    sp = SelfPlay(B, None, P, seeds=list(range(B)), device=0, auto_reset=True, keep_records=True, plan=pv.native_plan(B), arena_words=1 << 20)
    sp.capture_graph()
    T = {}

    def wrap(obj, name, key, sync=True):
        f = getattr(obj, name)

        def g(*a, **k):
            if sync:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f(*a, **k)
            if sync:
                torch.cuda.synchronize()
            T.setdefault(key, []).append(time.perf_counter() - t0)
            return r
        setattr(obj, name, g)
    e = sp.engine
    wrap(sp, "search", "search")
    wrap(e, "root_children", "root_children")
    wrap(e, "play", "play(+status)")
    wrap(e, "reset", "reset")
    L = _lib.lib()
    f0 = L.cz_host_choose_moves

    class Proxy:
        def __getattr__(self, k):
            return getattr(L, k)

        def cz_host_choose_moves(self, *a):
            t0 = time.perf_counter()
            r = f0(*a)
            T.setdefault("choose_moves", []).append(time.perf_counter() - t0)
            return r
    _lib._lib = Proxy()
    for _ in range(3):
        sp.step()
    for k in T:
        T[k].clear()
    tot = []
    for _ in range(plies):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sp.step()
        torch.cuda.synchronize()
        tot.append(time.perf_counter() - t0)
    out = {k: 1e3 * float(np.mean(v)) for k, v in T.items() if v}
    out["step_total_ms"] = 1e3 * float(np.mean(tot))
    out["host_other_ms"] = out["step_total_ms"] - sum(v for k, v in out.items() if k not in ("step_total_ms", "choose_moves"))
    print(json.dumps(out))
    if "--cprofile" in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(plies):
            sp.step()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()

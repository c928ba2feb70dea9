"""The one collective of the path: all_gather of finished-game (s, pi, z) tuples.

Games are independent, so ranks never communicate during search; when games end, each rank packs its new
tuples into fixed-size records and one all_gather_into_tensor (NCCL over NVLink on GPUs, gloo in the CPU
tests) gives every rank the whole batch for its replay buffer (main.py:1234-1240 feeds data_buffer)."""
import numpy as np
import torch
import torch.distributed as dist

STATE_BYTES = 100          # canonical state string (<= 90 squares + 9 slashes), zero padded
MAXC = 128
REC_BYTES = STATE_BYTES + 4 + MAXC * 2 + MAXC * 8 + 8      # state | n | int16 label[128] | f64 prob[128] | f64 z


def pack_records(records, cap):
    """records: iterable of GameRecord (selfplay.py).  Returns (uint8 [cap, REC_BYTES], n_packed, leftover)."""
    buf = np.zeros((cap, REC_BYTES), dtype=np.uint8)
    k = 0
    flat = [(s, ix, pv, z) for r in records for s, ix, pv, z in zip(r.states, r.pi_idx, r.pi_val, r.z)]
    for s, ix, pv, z in flat[:cap]:
        row = buf[k]
        sb = s.encode()
        row[: len(sb)] = np.frombuffer(sb, dtype=np.uint8)
        n = len(ix)
        row[STATE_BYTES:STATE_BYTES + 4] = np.frombuffer(np.int32(n).tobytes(), dtype=np.uint8)
        o = STATE_BYTES + 4
        row[o:o + 2 * n] = np.frombuffer(np.asarray(ix, dtype=np.int16).tobytes(), dtype=np.uint8)
        o += 2 * MAXC
        row[o:o + 8 * n] = np.frombuffer(np.asarray(pv, dtype=np.float64).tobytes(), dtype=np.uint8)
        o += 8 * MAXC
        row[o:o + 8] = np.frombuffer(np.float64(z).tobytes(), dtype=np.uint8)
        k += 1
    return buf, k, flat[cap:]


def unpack_records(buf, n):
    """-> list of (state str, pi dense float64 [2086], z float)"""
    out = []
    for row in np.asarray(buf)[:n]:
        s = bytes(row[:STATE_BYTES]).rstrip(b"\0").decode()
        c = int(np.frombuffer(row[STATE_BYTES:STATE_BYTES + 4].tobytes(), dtype=np.int32)[0])
        o = STATE_BYTES + 4
        ix = np.frombuffer(row[o:o + 2 * c].tobytes(), dtype=np.int16)
        o += 2 * MAXC
        pv = np.frombuffer(row[o:o + 8 * c].tobytes(), dtype=np.float64)
        o += 8 * MAXC
        z = float(np.frombuffer(row[o:o + 8].tobytes(), dtype=np.float64)[0])
        pi = np.zeros(2086)
        pi[ix.astype(np.int64)] = pv
        out.append((s, pi, z))
    return out


def all_gather_tuples(records, device, cap=2048, group=None):
    """Every rank contributes up to `cap` tuples; returns the list of all ranks' tuples (rank-major)."""
    world = dist.get_world_size(group)
    buf, k, _ = pack_records(records, cap)
    mine = torch.from_numpy(buf).to(device)
    cnt = torch.tensor([k], dtype=torch.int32, device=device)
    allr = torch.empty((world * cap, REC_BYTES), dtype=torch.uint8, device=device)   # rank-major concatenation
    allc = torch.empty((world,), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(allr, mine, group=group)
    dist.all_gather_into_tensor(allc, cnt, group=group)
    allr, allc = allr.cpu().numpy().reshape(world, cap, REC_BYTES), allc.cpu().numpy()
    out = []
    for r in range(world):
        out.extend(unpack_records(allr[r], int(allc[r])))
    return out


def shard_seeds(n_games_per_rank, rank, base_seed=0):
    """Per-game RNG seeds that do not depend on the world size: game g of rank r gets base + r*n + g."""
    return [base_seed + rank * n_games_per_rank + g for g in range(n_games_per_rank)]

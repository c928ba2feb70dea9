"""TEST INFRASTRUCTURE ONLY -- numpy versions of the deterministic stand-in evaluators.

Pure numpy, no dependency on /root/reference, so GPU-side tests may import it (e.g. to hand a host
`forward` callable to the MCTS_tree / cchess_main drop-ins).  The same functions are restated in
oracle/cchess_oracle.c (co_fake_forward) and cchess_zero_b200/fakenet.py (torch, on device)."""
import numpy as np

# --------------------------------------------------------------------------------------
# Deterministic stand-in evaluators (the reference's TF network cannot run).
# Both are functions of the encode tensor only, produce float32 values that are exactly
# computable with integer arithmetic, and are re-implemented independently in
# oracle/cchess_oracle.c (CPU) and cchess_zero_b200/fakenet.py (torch, on device).
# --------------------------------------------------------------------------------------
M32 = np.uint64(0xFFFFFFFF)


def _mix32(h):
    h = np.asarray(h, dtype=np.uint64) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846CA68B)) & M32
    h ^= h >> np.uint64(16)
    return h


def fake_forward_hash(positions, signed=True):
    """logits[j], value = 24-bit hashes of the set of non-zero cells, scaled by 2^-23 / 2^-24."""
    x = np.asarray(positions, dtype=np.float32).reshape(len(positions), -1)
    B = x.shape[0]
    idx = (np.arange(x.shape[1], dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B1) & M32
    key = ((x != 0).astype(np.uint64) * idx[None, :]).sum(axis=1) & M32
    key = _mix32(key)
    j = np.arange(2086, dtype=np.uint64)
    h = _mix32((key[:, None] + j[None, :] * np.uint64(0x85EBCA6B) + np.uint64(1)) & M32) >> np.uint64(8)
    hv = _mix32(key ^ np.uint64(0xC2B2AE35)) >> np.uint64(8)
    if signed:
        logits = ((h.astype(np.int64) - (1 << 23)).astype(np.float32) / np.float32(1 << 23))
    else:
        logits = (h.astype(np.int64).astype(np.float32) / np.float32(1 << 24))
    value = ((hv.astype(np.int64) - (1 << 23)).astype(np.float32) / np.float32(1 << 23)).reshape(B, 1)
    return logits.astype(np.float32), value.astype(np.float32)


def fake_forward_mod17(positions):
    """SURVEY Appendix B's exactly-representable net (multiples of 1/16)."""
    x = np.asarray(positions, dtype=np.float32).reshape(len(positions), -1)
    B = x.shape[0]
    logits = np.zeros((B, 2086), dtype=np.float32)
    value = np.zeros((B, 1), dtype=np.float32)
    j = np.arange(2086, dtype=np.int64)
    for b in range(B):
        c = np.nonzero(x[b])[0].astype(np.int64)
        s = ((131 * c[:, None] + 31 * j[None, :]) % 17).sum(axis=0) % 17
        logits[b] = (s - 8).astype(np.float32) / np.float32(16)
        value[b, 0] = np.float32((int(c.sum()) % 17) - 8) / np.float32(16)
    return logits, value


FAKE_NETS = {
    "hash_signed": lambda p: fake_forward_hash(p, True),
    "hash_pos": lambda p: fake_forward_hash(p, False),
    "mod17": fake_forward_mod17,
}
NET_IDS = {"hash_signed": 0, "hash_pos": 1, "mod17": 2}


def f32_bits(v):
    """Bit pattern of float32(v); every NaN is reported as 0x7FC00000 (x86 produces 0xFFC00000 for
    0/0, the GPU 0x7FFFFFFF -- the payload carries no meaning and the reference never inspects it)."""
    v = np.float32(v)
    return 0x7FC00000 if np.isnan(v) else int(v.view(np.uint32))



"""Deterministic stand-in evaluators on the device (torch integer ops) for parity tests and smoke().

They restate oracle/ref_harness.py's fake_forward_hash / fake_forward_mod17 (functions of the encode
tensor only, float32 outputs exactly computable with integer arithmetic) so that the CUDA tree can be
compared bit-for-bit with the reference / the CPU oracle, which cannot share a real network's
floating-point rounding.  Not used by the product's self-play path (that uses net.py)."""
import torch

M32 = 0xFFFFFFFF


def _mix32(h):
    h = h & M32
    h = h ^ (h >> 16)
    h = (h * 0x7FEB352D) & M32
    h = h ^ (h >> 15)
    h = (h * 0x846CA68B) & M32
    h = h ^ (h >> 16)
    return h


class FakeNet:
    def __init__(self, kind, device="cuda"):
        assert kind in ("hash_signed", "hash_pos", "mod17")
        self.kind = kind
        dev = torch.device(device)
        self.idx = ((torch.arange(1260, dtype=torch.int64, device=dev) + 1) * 0x9E3779B1) & M32
        self.j = torch.arange(2086, dtype=torch.int64, device=dev)
        if kind == "mod17":
            c = torch.arange(1260, dtype=torch.int64, device=dev)
            self.M = ((131 * c[:, None] + 31 * self.j[None, :]) % 17).to(torch.float32)
            self.c = c.to(torch.float32)

    @torch.no_grad()
    def __call__(self, x):
        """x: [B, 9,10,14] (any float dtype, 0/1 valued) -> (logits f32 [B,2086], value f32 [B])"""
        B = x.shape[0]
        nz = x.reshape(B, -1) != 0
        if self.kind == "mod17":
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            s = (nz.to(torch.float32) @ self.M).to(torch.int64) % 17
            sc = (nz.to(torch.float32) @ self.c).to(torch.int64) % 17
            torch.backends.cuda.matmul.allow_tf32 = prev
            return ((s - 8).to(torch.float32) / 16.0).contiguous(), ((sc - 8).to(torch.float32) / 16.0).contiguous()
        key = (nz.to(torch.int64) * self.idx[None, :]).sum(dim=1) & M32
        key = _mix32(key)
        h = _mix32((key[:, None] + self.j[None, :] * 0x85EBCA6B + 1) & M32) >> 8
        hv = _mix32(key ^ 0xC2B2AE35) >> 8
        if self.kind == "hash_signed":
            logits = (h - (1 << 23)).to(torch.float32) / float(1 << 23)
        else:
            logits = h.to(torch.float32) / float(1 << 24)
        value = (hv - (1 << 23)).to(torch.float32) / float(1 << 23)
        return logits.contiguous(), value.contiguous()

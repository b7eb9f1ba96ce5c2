"""VAE encode / decode alone at the bench batch (256 sequences x 60 frames): device time and the attention kernel's share.
    python tools/profile_vae.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import _lib, ops, synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
op = ops.VaeOp({k: v.to(dev) for k, v in S.vae_state(0).items()}, 75, max_batch=B, max_frames=60)
g = torch.Generator().manual_seed(0)
feats = torch.randn(B, 60, 75, generator=g).to(dev)
eps = torch.randn(B, 256, generator=g).to(dev)
z = torch.randn(B, 256, generator=g).to(dev)
lens = torch.full((B,), 60, dtype=torch.int32)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, fn in (("encode", lambda: op.encode(feats, lens, eps)), ("decode", lambda: op.decode(z, lens, 60))):
    ms = timed(fn)
    for i in range(8):
        _lib.prof_read(i)
    _lib.prof_enable(True)
    fn()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    a_ms, a_n = _lib.prof_read(4)
    print(f"vae {name} B={B}: {ms:.3f} ms; attention kernel {a_n} launches, {a_ms / max(a_n, 1) * 1e3:.1f} us each")
out = op.decode(z, lens, 60)
print("decode checksum", float(out.double().abs().sum()))

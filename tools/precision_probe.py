"""CPU probe of GEMM operand formats (test tooling, not product): runs the oracle's ego_eval with the
operands of every F.linear in one stage rounded to a candidate format (fp32 accumulate) and reports the
drift of the predicted joints against the all-fp32 run.  Used to choose the tensor-core operand format per
stage (DESIGN.md "precision placement").

  python tools/precision_probe.py [B] [n_points]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restate as O  # noqa: E402
from seeme_b200 import synthetic as S  # noqa: E402

_real_linear = F.linear
MODE = {"cur": "fp32"}


def _q(x, fmt):
    if fmt == "fp32":
        return x
    if fmt == "bf16":
        return x.bfloat16().float()
    if fmt == "fp16":
        return x.half().float()
    if fmt == "tf32":
        i = x.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    raise ValueError(fmt)


def _lin(x, w, b=None):
    m = MODE["cur"]
    if m == "fp32":
        return _real_linear(x, w, b)
    if m in ("bf16", "fp16", "tf32"):
        return _real_linear(_q(x, m), _q(w, m), b)
    if m in ("bf16x3", "fp16x3"):
        f = m[:4]
        xh, wh = _q(x, f), _q(w, f)
        xl, wl = _q(x - xh, f), _q(w - wh, f)
        return _real_linear(xh, wh, b) + _real_linear(xl, wh) + _real_linear(xh, wl)
    if m in ("fp16x2",):      # exact-ish activation (hi+lo), fp16 weight
        f = "fp16"
        xh, wh = _q(x, f), _q(w, f)
        xl = _q(x - xh, f)
        return _real_linear(xh, wh, b) + _real_linear(xl, wh)
    raise ValueError(m)


class stage_mode:
    def __init__(self, fn_name, mode):
        self.fn_name, self.mode = fn_name, mode

    def __enter__(self):
        self.orig = getattr(O, self.fn_name)
        orig, mode = self.orig, self.mode

        def wrapped(*a, **k):
            prev = MODE["cur"]
            MODE["cur"] = mode
            try:
                return orig(*a, **k)
            finally:
                MODE["cur"] = prev
        setattr(O, self.fn_name, wrapped)

    def __exit__(self, *a):
        setattr(O, self.fn_name, self.orig)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    gs = float(sys.argv[3]) if len(sys.argv) > 3 else 7.5
    torch.set_num_threads(8)
    W = {"denoiser": S.denoiser_state(0), "vae": S.vae_state(0), "pointnet": S.pointnet_state(0),
         "output_scene": S.output_scene_state(0)}
    smpl, stats = S.smpl_buffers(), S.norm_stats()
    batch = S.make_batch(B, n_points=N)
    g = torch.Generator().manual_seed(7)
    noise = {"eps_int": torch.randn(1, B, 256, generator=g), "eps_unc": torch.randn(1, B, 256, generator=g),
             "x_T": torch.randn(B, 1, 256, generator=g)}
    F.linear = _lin
    O.F.linear = _lin

    def run():
        with torch.no_grad():
            return O.ego_eval(W, smpl, stats, batch, noise, guidance_scale=gs)

    ref = run()
    print(f"B={B} N={N} guidance={gs} |z|max={float(ref['z'].abs().max()):.1f}")
    stages = {"pointnet": "pointnet_forward", "denoiser": "denoiser_forward", "vae_dec": "vae_decode",
              "vae_enc": "vae_encode"}
    for st, fn in stages.items():
        for mode in ("bf16", "fp16", "tf32", "fp16x2", "bf16x3", "fp16x3"):
            with stage_mode(fn, mode):
                r = run()
            d = (r["joints_rst"] - ref["joints_rst"])
            mpjpe = d.norm(dim=-1).mean() * 1000
            zrel = float((r["z"] - ref["z"]).abs().max() / ref["z"].abs().max())
            print(f"{st:9s} {mode:7s} joints max-abs {float(d.abs().max()) * 1000:9.4f} mm   MPJPE drift {float(mpjpe):8.4f} mm"
                  f"   z rel {zrel:.2e}", flush=True)


if __name__ == "__main__":
    main()

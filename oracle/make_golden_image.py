"""TEST INFRASTRUCTURE ONLY -- golden output of the reference's image backbone for ``tests/``.

Builds the UNMODIFIED ``resnet50`` of ``/root/reference/EgoHMR/models/resnet.py:99-224`` (``pretrained=False``: the
ImageNet checkpoint is a download), loads the seeded ``seeme_b200.synthetic.resnet50_state`` with ``strict=True``, runs it in
eval mode on seeded crops and stores the ``[B,2048]`` features in ``tests/golden/resnet50_image.npz``; also checks that
``oracle/restate.image_backbone_forward`` reproduces them.  Container only (the reference tree is absent on the GPU box).

    python oracle/make_golden_image.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SEEME_REFERENCE_ROOT", "/root/reference")
BATCH, SEED = 2, 0


def main():
    from seeme_b200 import synthetic as S
    from oracle import restate
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_resnet", os.path.join(REF, "EgoHMR", "models", "resnet.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    net = ref.resnet50(pretrained=False)
    sd = S.resnet50_state(SEED)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = S.images(BATCH, SEED)
    with torch.no_grad():
        want = net(x)
        got = restate.image_backbone_forward(sd, x)
    err = (want - got).abs().max().item()
    print(f"reference vs restatement: max|diff| = {err:.3e}, |feat| max = {want.abs().max().item():.3f}, mean = {want.mean().item():.4f}")
    assert err <= 1e-5 * max(1.0, want.abs().max().item())
    out = os.path.join(ROOT, "tests", "golden", "resnet50_image.npz")
    np.savez_compressed(out, feat=want.numpy(), batch=BATCH, seed=SEED)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

"""Event trace of the tensor-core block-0 kernel (CTA 0, first tiles): builds an instrumented copy of the library
(-DPF_TRACE) next to the product one, runs one scene-encoder pass and prints per-tile event times (cycles).

    python tools/pf_trace.py build      # here (no GPU): writes seeme_b200/lib/libseeme_b200_trace.so
    python tools/pf_trace.py            # on the GPU box
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACE_LIB = os.path.join(ROOT, "seeme_b200", "lib", "libseeme_b200_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    sys.path.insert(0, ROOT)
    from seeme_b200 import build as B
    B.build(verbose=False)
    obj = os.path.join(B.OBJ, "pointnet_fused_trace.o")
    subprocess.run([B.NVCC, *B.ARCH, *[f for f in B.FLAGS if f not in ("-Xptxas", "-v")], "-DPF_TRACE", "-c",
                    os.path.join(B.CSRC, "pointnet_fused.cu"), "-o", obj], check=True)
    objs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B.sources() if s != "pointnet_fused.cu"] + [obj]
    subprocess.run([B.NVCC, *B.ARCH, "-shared", "-o", TRACE_LIB, *objs], check=True)
    print(TRACE_LIB)
    sys.exit(0)

os.environ["SEEME_B200_LIB"] = TRACE_LIB
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from seeme_b200 import _lib, ops, synthetic as S  # noqa: E402

B = 32
dev = "cuda:0"
W = {k: v.to(dev) for k, v in S.pointnet_state(0).items()}
Wo = {k: v.to(dev) for k, v in S.output_scene_state(0).items()}
op = ops.PointNetOp(W, Wo, max_batch=B, max_points=20000, precision=16)
p = S.egobody_scene(B, 20000, torch.Generator().manual_seed(3)).to(dev)
for _ in range(3):
    op(p)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "block":
    # the pointnet_block_kernel launches come after block 0 in a pass: the trace buffer holds the LAST launch (block 3)
    TILES, SLOTS = 24, 64
    buf = (C.c_longlong * (TILES * SLOTS))()
    assert _lib.lib().seeme_pf_trace_read(buf, TILES * SLOTS) == TILES * SLOTS
    t = [[buf[i * SLOTS + k] for k in range(SLOTS)] for i in range(TILES)]
    names = {0: "wait x_full", 1: "x_full seen", 6: "out_drained(prev) seen"}
    for kc in range(4):
        names[2 + kc] = f"S({kc}) issue"
        names[7 + kc] = f"G1(nh{kc >> 1},kp{kc & 1}) issue"
        names[11 + kc] = f"G2({kc}) issue"
        names[16 + 2 * kc] = f"  H: s_done{kc}"
        names[17 + 2 * kc] = f"  H: relu{kc} done"
    names.update({24: "  H: h_full0", 25: "  H: epiH0 done", 26: "  H: h_full1", 27: "  H: epiH1 done", 28: "    O: out_full",
                  29: "    O: out_drained", 30: "    O: staged", 31: "    O: stored + next X load issued"})
    for j in (8, 9):
        base = t[j][2]
        print(f"--- tile {j}: S(0) issue of tile {j + 1} at +{t[j + 1][2] - base}")
        for k, v in sorted(((k, t[j][k]) for k in names if t[j][k]), key=lambda kv: kv[1]):
            print(f"{v - base:8d}  {names[k]}")
    print("tile period (cycles):", [t[j + 1][2] - t[j][2] for j in range(2, TILES - 1)])
    sys.exit(0)
TILES, SLOTS = 24, 64
buf = (C.c_longlong * (TILES * SLOTS))()
n = _lib.lib().seeme_pf_trace_read(buf, TILES * SLOTS)
assert n == TILES * SLOTS, n
t = [[buf[i * SLOTS + k] for k in range(SLOTS)] for i in range(TILES)]
names = {0: "F0 issue"}
for kc in range(8):
    names[1 + 2 * kc] = f"wait full{kc}"
    names[2 + 2 * kc] = f"G1({kc}) issue"
    names[26 + 2 * kc] = f"  H: f_full{kc}"
    names[27 + 2 * kc] = f"  H: gen{kc} done"
names[17] = "PF issue (out_drained)"
for kc in range(4):
    names[18 + 2 * kc] = f"wait g2_full{kc}"
    names[19 + 2 * kc] = f"G2({kc}) issue"
for st in range(12):
    names[48 + st] = f"        P: load step {st}"
names.update({60: "  H: gen4 after empty wait", 61: "  H: gen4 after tmem ld", 62: "  H: gen4 after cvt+sts", 63: "  H: gen4 after fence"})
names.update({42: "  H: h_full", 43: "  H: epiH done", 44: "    O: out_full", 45: "    O: fr_free", 46: "    O: out_drained", 47: "    O: tile done"})
for j in (8, 9):
    base = t[j][2]
    print(f"--- tile {j}: G1(0) issue of tile {j + 1} at +{t[j + 1][2] - base}")
    for k, v in sorted(((k, t[j][k]) for k in names if t[j][k]), key=lambda kv: kv[1]):
        print(f"{v - base:8d}  {names[k]}")
per = [t[j + 1][2] - t[j][2] for j in range(2, TILES - 1)]
print("tile period (cycles):", per)

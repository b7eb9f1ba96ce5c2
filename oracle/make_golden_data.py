"""TEST INFRASTRUCTURE ONLY -- golden items of the reference's EgoBody dataset class for ``tests/test_data.py``.

Writes the seeded synthetic recordings of ``seeme_b200.egobody_data.write_synthetic`` in the reference's on-disk layout,
loads them with the UNMODIFIED ``EgoBodyData3`` (``/root/reference/mld/data/humanml/data/dataset.py:1055-1794``; the
third-party imports it never calls on this branch -- smplx, spacy, trimesh, rich, yacs -- are stubbed) and stores every
item in ``tests/golden/egobody_items.npz``.  Container only (the reference tree does not exist on the GPU box).

    python oracle/make_golden_data.py
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SEEME_REFERENCE_ROOT", "/root/reference")
LENGTHS, SEED, N_POINTS = (60, 60, 37, 60, 12), 0, 2000


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m


def reference_items(root_parent):
    sys.path.insert(0, REF)
    for name in ["smplx", "spacy", "trimesh", "rich", "rich.progress", "yacs", "yacs.config", "clip"]:
        sys.modules.setdefault(name, _Stub(name))
    sys.modules["rich.progress"].track = lambda it, *a, **k: it
    sys.modules["yacs.config"].CfgNode = dict
    import mld.data.humanml.data.dataset as D
    cwd = os.getcwd()
    os.chdir(root_parent)           # the class reads ./datasets/EgoBody/...
    try:
        ds = D.EgoBodyData3(None, None, "test.txt", "./datasets/EgoBody/our_process_smpl_split_NEW",
                            condition=["text", "scene", "interactee"], predict_transl=True, motion_length=60,
                            data_type="angle", progress_bar=False)
        return {ds.name_list[i]: ds[i] for i in range(len(ds))}
    finally:
        os.chdir(cwd)


def reference_gimo_items(root_parent):
    """GimoData with condition ["text"]: the scene branch needs trimesh to read the scan and cannot run here"""
    import mld.data.humanml.data.dataset as D
    cwd = os.getcwd()
    os.chdir(root_parent)
    try:
        ds = D.GimoData(None, None, "test.txt", "./datasets/GIMO/processed", condition=["text"], predict_transl=True,
                        motion_length=60, data_type="angle", progress_bar=False)
        return {ds.name_list[i]: ds[i] for i in range(len(ds))}
    finally:
        os.chdir(cwd)


def main():
    from seeme_b200.egobody_data import write_synthetic, write_synthetic_gimo
    with tempfile.TemporaryDirectory() as tmp:
        write_synthetic(os.path.join(tmp, "datasets", "EgoBody"), "test", LENGTHS, SEED, N_POINTS)
        write_synthetic_gimo(os.path.join(tmp, "datasets", "GIMO"), "test", 3, SEED)
        items = reference_items(tmp)
        gimo = reference_gimo_items(tmp)
    out = {}
    for name, it in items.items():
        motion, transl, beta, utils_, scene, length, imgs = it
        key = name[:-4]
        for k, v in (("motion", motion), ("transl", transl), ("beta", beta), ("utils", utils_), ("scene", scene), ("length", length)):
            out[f"{key}/{k}"] = v.numpy()
        out[f"{key}/imgs"] = np.array(imgs)
    for name, it in gimo.items():
        motion, transl, beta, utils_, length = it
        for k, v in (("motion", motion), ("transl", transl), ("beta", beta), ("utils", utils_), ("length", length)):
            out[f"gimo/{name[:-4]}/{k}"] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "egobody_items.npz")
    np.savez_compressed(path, **out)
    print(path, len(items), "items;", {k: (v.dtype, v.shape) for k, v in out.items() if k.startswith("seq_0002")})


if __name__ == "__main__":
    main()

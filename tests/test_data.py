"""Batch producer for the reference's on-disk EgoBody format against items of the UNMODIFIED reference
dataset class (``tests/golden/egobody_items.npz``, made by ``oracle/make_golden_data.py``)."""
import os

import numpy as np
import pytest
import torch

from seeme_b200 import egobody_data as E

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "egobody_items.npz")
LENGTHS, SEED, N_POINTS = (60, 60, 37, 60, 12), 0, 2000


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("egobody") / "datasets" / "EgoBody")
    E.write_synthetic(root, "test", LENGTHS, SEED, N_POINTS)
    return E.EgoBodySequences(root, "test", condition=["text", "scene", "interactee"], motion_length=60, predict_transl=True)


def test_items_match_reference_dataset_bit_for_bit(dataset):
    g = np.load(GOLDEN)
    assert len(dataset) == len(LENGTHS)
    for i, name in enumerate(dataset.names):
        motion, transl, beta, utils_, scene, length, imgs = dataset[i]
        key = name[:-4]
        for k, v in (("motion", motion), ("transl", transl), ("beta", beta), ("utils", utils_), ("scene", scene), ("length", length)):
            ref = g[f"{key}/{k}"]
            assert v.numpy().dtype == ref.dtype and tuple(v.shape) == ref.shape, (key, k, v.dtype, ref.dtype)
            assert np.array_equal(v.numpy(), ref), (key, k, float(np.abs(v.numpy() - ref).max()))
        assert list(imgs) == list(g[f"{key}/imgs"])
        assert int(length) == LENGTHS[i]
        # zero padding happens BEFORE normalisation (dataset.py:1497-1530): padded pose rows are -mean/std, not zero
        if LENGTHS[i] < 60:
            pad = motion[LENGTHS[i]:, 0, 3:]
            assert torch.allclose(pad, torch.tensor(-dataset.mean[0, 3:72] / dataset.std[0, 3:72]).expand_as(pad).to(pad.dtype))


def test_collate_and_prefetching_batches(dataset):
    got = list(E.batches(dataset, batch_size=2, pin=False, prefetch=2))
    assert [b[0].shape[0] for b in got] == [2, 2, 1]
    motion, transl, beta, utils_, scene, length, imgs = got[0]
    assert motion.shape == (2, 60, 2, 72) and transl.shape == (2, 2, 60, 3) and beta.shape == (2, 2, 60, 10)
    assert utils_.shape == (2, 60, 6) and scene.shape == (2, N_POINTS, 3) and length.shape == (2, 1)
    assert len(imgs) == 60 and len(imgs[0]) == 2 and isinstance(imgs[0][0], str)          # T tuples of B strings (mld.py:1101)
    # same values as item-wise access, and the rank shard selector
    assert torch.equal(got[1][0][1], dataset[3][0])
    shard = list(E.batches(dataset, batch_size=2, pin=False, indices=[4, 0]))
    assert torch.equal(shard[0][5].reshape(-1), torch.tensor([12, 60], dtype=torch.int32))


@pytest.mark.gpu
def test_pipeline_consumes_producer_batches(dataset):
    """pinned batches of the producer through MLD.run_test_batches (ragged lengths, host-resident inputs)"""
    import seeme_b200
    dev = "cuda:0"
    model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=7.5, max_batch=2, n_points=N_POINTS, pipeline_depth=2)
    outs = list(model.run_test_batches(E.batches(dataset, batch_size=2, pin=True)))
    # a batch is decoded to max(lengths) frames (mld_vae.py:253, mld.py:1405-1408): the last batch holds one 12-frame sequence
    assert [tuple(o.shape) for o in outs] == [(2, 60, 24, 3), (2, 60, 24, 3), (1, 12, 24, 3)]
    assert all(bool(torch.isfinite(o).all()) for o in outs)
    # the same sequences through the synchronous call give the same joints
    b = next(iter(E.batches(dataset, batch_size=2, pin=False)))
    torch.manual_seed(0)
    rs = model.ego_eval(tuple(x.to(dev) if torch.is_tensor(x) else x for x in b))
    assert rs["lengths"] == [60, 60] and rs["joints_rst"].shape == (2, 60, 24, 3)


def test_gimo_items_match_reference_dataset_and_scene_branch(tmp_path):
    """GimoData (dataset.py:1797-2509): motion / transl / beta / utils bit-identical to the unmodified class (goldens made with
    condition ["text"]: its scene branch needs trimesh); the scene branch -- own PLY reader, 1/1.03 rescale, transform_norm --
    is checked against a direct NumPy evaluation with the same RNG state (unpinned against the reference)."""
    root = str(tmp_path / "datasets" / "GIMO")
    E.write_synthetic_gimo(root, "test", 3, SEED)
    g = np.load(GOLDEN)
    ds = E.GimoSequences(root, "test", condition=["text"])
    assert len(ds) == 3
    for i, name in enumerate(ds.names):
        motion, transl, beta, utils_, length = ds[i]
        for k, v in (("motion", motion), ("transl", transl), ("beta", beta), ("utils", utils_), ("length", length)):
            ref = g[f"gimo/{name[:-4]}/{k}"]
            assert v.numpy().dtype == ref.dtype and np.array_equal(v.numpy(), ref), (name, k)
    assert motion.shape == (60, 2, 66)
    # scene branch
    dss = E.GimoSequences(root, "val", condition=["text", "scene"], n_points=777)           # "val" falls back to the test split
    np.random.seed(5)
    motion2, _, _, _, scene, length, imgs = dss[1]
    assert torch.equal(motion2, ds[1][0]) and scene.shape == (777, 3) and scene.dtype == torch.float32 and len(imgs) == 60
    obj = os.path.join(os.path.dirname(root), "gimo_raw", "group", "GIMO", imgs[0].split("/")[-4], "scene_obj")
    pts = E.read_ply_vertices(os.path.join(obj, "scene_downsampled.ply"))
    assert pts.shape == (5000, 3)
    np.random.seed(5)
    sel = pts[np.random.choice(range(len(pts)), 777)] / 1.03
    tn = np.loadtxt(os.path.join(obj, "transform_norm.txt")).reshape(4, 4)
    want = sel @ tn[:3, :3].T + tn[:3, 3] / 1.03
    assert np.allclose(scene.numpy(), want, atol=1e-5)
    # the model consumes GIMO batches with the 66-wide features
    b = E.collate([dss[0], dss[1]])
    assert b[0].shape == (2, 60, 2, 66) and b[4].shape == (2, 777, 3)
    with pytest.raises(ValueError):
        E.GimoSequences(root, "test", condition=["text"], motion_length=30)


@pytest.mark.gpu
def test_gimo_model_consumes_gimo_producer_batches(tmp_path):
    """config_mld_gimo.yaml (scene-only conditioning, 66-wide rows) on pinned batches of the GIMO producer"""
    import seeme_b200
    root = str(tmp_path / "datasets" / "GIMO")
    E.write_synthetic_gimo(root, "test", 3, SEED)
    ds = E.GimoSequences(root, "test", condition=["text", "scene"], n_points=1500)
    model = seeme_b200.build_model("config_mld_gimo.yaml", device="cuda:0", guidance_scale=7.5, max_batch=2, n_points=1500, pipeline_depth=2)
    np.random.seed(0)
    outs = list(model.run_test_batches(E.batches(ds, batch_size=2, pin=True)))
    assert [tuple(o.shape) for o in outs] == [(2, 60, 24, 3), (1, 60, 24, 3)]
    assert all(bool(torch.isfinite(o).all()) for o in outs)

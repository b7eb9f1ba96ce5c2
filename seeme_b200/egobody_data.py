"""Batch producer for the reference's on-disk EgoBody format: the step in front of the hot path.

Mirrors the live branch of ``EgoBodyData3`` (``mld/data/humanml/data/dataset.py:1055-1794``) for the north-star
configuration -- ``DATA_TYPE angle``, scene (+ interactee) conditioning, ``PREDICT_TRANSL`` -- item for item, bit for bit
(``tests/test_data.py`` checks it against items produced by the unmodified reference class), and adds what the B200 path
needs on top: batches collated straight into PINNED host memory by a background thread, so that
``MLD.ego_eval_async`` / ``run_test_batches`` copy them on a pipeline slot's stream without ever blocking.

On-disk layout (relative to ``root``, the reference hard-codes ``./datasets/EgoBody``; ``dataset.py:1086-1230``):

    our_process_smpl_split_NEW/{mean,std}.npy                      float64 [1, >= 75]: 3 global-orient + 69 pose + 3 transl dims
    our_process_smpl_split_NEW/<split>/<name>.npy                   pickled dict per sequence:
        video [T], recording_utils {original_imgname [T] str, fx/cx/cy/scale [T], center [T,2]},
        wearer / interactee {global_orient [T,1,3], body_pose [T,1,69], betas [T,1,10], transl [T,1,3]}
    Egohmr_scene_preprocess_s1_release/map_dict_<split>.pkl         image name -> scene key
    Egohmr_scene_preprocess_s1_release/pcd_verts_dict_<split>.pkl   scene key -> [20000, 3] points, kinect main frame
    transf_matrices_all_seqs.pkl                                    recording -> {trans_kinect2holo [4,4], trans_world2pv {timestamp: [4,4]}}

An item is the reference's tuple ``(motion [60,2,72], transl [2,60,3], beta [2,60,10], utils [60,6], scene [20000,3],
length [1] int32, list_imgname)``; a batch is its ``default_collate`` (``list_imgname`` becomes T tuples of B strings,
``mld.py:1101``).  Sequences shorter than ``motion_length`` are zero-padded BEFORE normalisation, as the reference does.
"""
from __future__ import annotations

import os
import pickle
import queue
import threading
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch

# kinect -> PV camera: flip y and z after world2pv . kinect2holo (dataset.py:1190-1192, 1283-1284)
_ADD_TRANS = np.array([[1.0, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]])


class EgoBodySequences(torch.utils.data.Dataset):
    POSE_DIMS = 69                  # body-pose dims per frame; the statistics are [3 global orient | POSE_DIMS | 3 transl | ...]

    def __init__(self, root: str, split: str = "test", condition: Sequence[str] = ("text", "scene", "interactee"),
                 motion_length: int = 60, predict_transl: bool = True):
        self.root, self.split = root, split
        self.condition = list(condition)
        self.motion_length = int(motion_length)
        self.predict_transl = bool(predict_transl)
        self._load(root, split)

    def _transl_stats(self):
        n = 3 + self.POSE_DIMS
        return self.mean[0, n:n + 3], self.std[0, n:n + 3]

    def _image_names(self, d) -> List[str]:
        return [str(n) for n in d["recording_utils"]["original_imgname"]]

    def _load(self, root: str, split: str) -> None:
        base = os.path.join(root, "our_process_smpl_split_NEW")
        self.mean = np.load(os.path.join(base, "mean.npy"))
        self.std = np.load(os.path.join(base, "std.npy"))
        seq_dir = os.path.join(base, split)
        # the reference keeps os.listdir order (its "sort by length" key is the dict's key count, a constant); sorted names
        # make the order reproducible across file systems
        self.names = sorted(n for n in os.listdir(seq_dir) if n.endswith(".npy"))
        self.items = [np.load(os.path.join(seq_dir, n), allow_pickle=True).item() for n in self.names]
        if "scene" in self.condition:
            pre = os.path.join(root, "Egohmr_scene_preprocess_s1_release")
            with open(os.path.join(pre, f"map_dict_{split}.pkl"), "rb") as f:
                self.scene_map = pickle.load(f)
            with open(os.path.join(pre, f"pcd_verts_dict_{split}.pkl"), "rb") as f:
                self.scene_verts = pickle.load(f)
            with open(os.path.join(root, "transf_matrices_all_seqs.pkl"), "rb") as f:
                self.transf = pickle.load(f)

    def __len__(self) -> int:
        return len(self.items)

    # -- pieces -----------------------------------------------------------------------------------------------------
    def _scene(self, first_image: str) -> torch.Tensor:
        parts = first_image.split("/")
        rec, stamp = parts[1], parts[4].split("_")[0]
        t = self.transf[rec]
        kinect2holo = t["trans_kinect2holo"].astype(np.float32)
        holo2pv = t["trans_world2pv"][str(stamp)].astype(np.float32)
        k2pv = np.matmul(_ADD_TRANS, np.matmul(holo2pv, kinect2holo))
        pts = self.scene_verts[self.scene_map[first_image]]
        pts = pts.dot(k2pv[:3, :3].transpose()) + k2pv[:3, 3].reshape(1, -1)
        return torch.tensor(pts, dtype=torch.float32)

    def _pad(self, x: torch.Tensor, n_pad: int) -> torch.Tensor:
        if n_pad == 0:
            return x
        return torch.cat([x, torch.zeros((n_pad,) + tuple(x.shape[1:]))], dim=0)

    def __getitem__(self, i: int):
        d = self.items[i]
        rec = d["recording_utils"]
        names = self._image_names(d)
        T = len(d["video"])
        L = self.motion_length
        n_pad = L - T
        mean, std = self.mean, self.std
        P = self.POSE_DIMS
        t_mean, t_std = self._transl_stats()

        def person(p):
            pose = np.array(p["body_pose"])
            if n_pad:
                pose = np.concatenate([pose, np.zeros((n_pad, 1, P))], axis=0)
            pose = (pose.reshape(L, -1) - mean[0, 3:3 + P]) / std[0, 3:3 + P]
            pose = torch.tensor(pose, dtype=torch.float32).unsqueeze(1)                       # [L,1,69]
            go = self._pad(torch.tensor(np.array(p["global_orient"]), dtype=torch.float32), n_pad)
            # float32 tensor (-, /) float64 statistics: the reference mixes a tensor with an ndarray, which promotes to float64
            go = (go - torch.from_numpy(mean[0, :3])) / torch.from_numpy(std[0, :3])
            tr = self._pad(torch.tensor(np.array(p["transl"]), dtype=torch.float32), n_pad)
            if self.predict_transl:
                tr = (tr - torch.from_numpy(np.ascontiguousarray(t_mean))) / torch.from_numpy(np.ascontiguousarray(t_std))
            be = self._pad(torch.tensor(np.array(p["betas"]), dtype=torch.float32), n_pad)
            return pose, go, tr, be

        w_pose, w_go, w_tr, w_be = person(d["wearer"])
        i_pose, i_go, i_tr, i_be = person(d["interactee"])
        motion = torch.cat([torch.cat([w_go, i_go], dim=1), torch.cat([w_pose, i_pose], dim=1)], dim=-1)   # [L,2,72]
        transl = torch.cat([w_tr, i_tr], dim=1).permute(1, 0, 2)                                            # [2,L,3]
        beta = torch.cat([w_be, i_be], dim=1).permute(1, 0, 2)                                              # [2,L,10]
        cols = [torch.tensor(np.array(rec[k])).reshape(-1, w) for k, w in (("fx", 1), ("cx", 1), ("cy", 1), ("center", 2), ("scale", 1))]
        utils_ = torch.cat(cols, dim=1)
        if n_pad:
            utils_ = torch.cat([utils_, torch.zeros((n_pad, 6))], dim=0)
        length = torch.tensor([T], dtype=torch.int32)
        if "scene" in self.condition:
            return motion, transl, beta, utils_, self._scene(names[0] if names else ""), length, names
        return motion, transl, beta, utils_, length


def read_ply_vertices(path: str) -> np.ndarray:
    """x, y, z of the vertex element of an ASCII or binary-little-endian PLY file (what ``trimesh.load_mesh(...).vertices`` returns
    for the GIMO scene scans; faces and extra vertex properties are skipped)."""
    with open(path, "rb") as f:
        fmt, n_vert, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element"):
                in_vertex = line.split()[1] == "vertex"
                if in_vertex:
                    n_vert = int(line.split()[2])
            elif line.startswith("property") and in_vertex:
                props.append((line.split()[-1], line.split()[1]))
            elif line == "end_header":
                break
        names = [p[0] for p in props]
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=n_vert, ndmin=2)
            return np.asarray(rows[:, [names.index(c) for c in "xyz"]], dtype=np.float64)
        if fmt != "binary_little_endian":
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        np_t = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1",
                "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4", "short": "<i2", "ushort": "<u2"}
        rec = np.frombuffer(f.read(n_vert * sum(np.dtype(np_t[t]).itemsize for _, t in props)),
                            dtype=np.dtype([(n, np_t[t]) for n, t in props]), count=n_vert)
        return np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float64)


class GimoSequences(EgoBodySequences):
    """``GimoData`` (``dataset.py:1797-2509``): same per-sequence dicts, 21 body joints (63 pose dims; the statistics are
    [3 | 63 | ... | 3 transl at the END]), image names taken from ``video``, and the scene cloud sampled (20 000 points with
    replacement through NumPy's global RNG, like the reference) from ``<scene_root>/<scene>/scene_obj/scene_downsampled.ply``,
    scaled by 1/1.03 and moved by ``transform_norm.txt``.  The motion / transl / beta / utils part is pinned against the
    unmodified class (``tests/test_data.py``); the scene branch is NOT (the reference needs ``trimesh`` to read the scan,
    which is not installable here), and the reference's zero-padding of short GIMO sequences is broken (it pads 69 pose
    dims onto 63): full-length sequences only, as in the released data."""
    POSE_DIMS = 63

    def __init__(self, root: str, split: str = "test", condition: Sequence[str] = ("text", "scene"), motion_length: int = 60,
                 predict_transl: bool = True, motion_dir: Optional[str] = None, scene_root: Optional[str] = None, n_points: int = 20000):
        self.motion_dir = motion_dir if motion_dir is not None else os.path.join(root, "processed")
        self.scene_root = scene_root if scene_root is not None else os.path.join(os.path.dirname(root.rstrip("/")), "gimo_raw", "group", "GIMO")
        self.n_points = int(n_points)
        super().__init__(root, split, condition, motion_length, predict_transl)

    def _transl_stats(self):
        return self.mean[0, -3:], self.std[0, -3:]

    def _image_names(self, d) -> List[str]:
        return [str(n) for n in d["video"]]

    def _load(self, root: str, split: str) -> None:
        self.mean = np.load(os.path.join(root, "processed", "mean.npy"))
        self.std = np.load(os.path.join(root, "processed", "std.npy"))
        seq_dir = os.path.join(self.motion_dir, "test" if split == "val" else split)       # GIMO has no val split (dataset.py:1841-1843)
        self.names = sorted(n for n in os.listdir(seq_dir) if n.endswith(".npy"))
        self.items = [np.load(os.path.join(seq_dir, n), allow_pickle=True).item() for n in self.names]
        for n, d in zip(self.names, self.items):
            if len(d["video"]) != self.motion_length:
                raise ValueError(f"{n}: {len(d['video'])} frames; GIMO sequences must have motion_length = {self.motion_length} frames")

    def _scene(self, first_image: str) -> torch.Tensor:
        scene = first_image.split("/")[-4]
        base = os.path.join(self.scene_root, scene, "scene_obj")
        scale = 1.03
        tn = np.loadtxt(os.path.join(base, "transform_norm.txt")).reshape((4, 4))
        tn[:3, 3] /= scale
        pts = read_ply_vertices(os.path.join(base, "scene_downsampled.ply"))
        pts = pts[np.random.choice(range(len(pts)), self.n_points)]
        pts = pts * (1 / scale)
        pts = (tn[:3, :3] @ pts.T + tn[:3, 3:]).T
        return torch.from_numpy(pts).float()


def collate(items: List[tuple], pin: bool = False):
    """``default_collate`` of the reference's item tuples; tensors optionally land in pinned host memory."""
    out = []
    for col in zip(*items):
        if torch.is_tensor(col[0]):
            t = torch.stack(list(col), dim=0)
            out.append(t.pin_memory() if pin else t)
        else:                                           # list_imgname: B lists of T strings -> T tuples of B strings
            out.append([tuple(x) for x in zip(*col)])
    return tuple(out)


def batches(dataset: EgoBodySequences, batch_size: int, pin: Optional[bool] = None, prefetch: int = 2,
            indices: Optional[Sequence[int]] = None) -> Iterator[tuple]:
    """Batches in dataset order (``shuffle=False`` like the reference's test loader), collated into pinned host memory by
    a background thread ``prefetch`` batches ahead of the consumer.  ``indices`` selects this rank's shard."""
    pin = torch.cuda.is_available() if pin is None else pin
    idx = list(range(len(dataset))) if indices is None else list(indices)
    chunks = [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]
    q: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))

    def work():
        try:
            for c in chunks:
                q.put(collate([dataset[i] for i in c], pin=pin))
            q.put(None)
        except BaseException as e:      # noqa: BLE001  (re-raised in the consumer)
            q.put(e)

    threading.Thread(target=work, daemon=True).start()
    while True:
        b = q.get()
        if b is None:
            return
        if isinstance(b, BaseException):
            raise b
        yield b


# ---- synthetic data in the on-disk format (tests, goldens, demos: the real EgoBody recordings are not redistributable) ----
def write_synthetic(root: str, split: str = "test", lengths: Sequence[int] = (60, 60, 37, 60, 12), seed: int = 0,
                    n_points: int = 20000) -> None:
    g = np.random.default_rng(seed)
    base = os.path.join(root, "our_process_smpl_split_NEW")
    os.makedirs(os.path.join(base, split), exist_ok=True)
    os.makedirs(os.path.join(root, "Egohmr_scene_preprocess_s1_release"), exist_ok=True)
    np.save(os.path.join(base, "mean.npy"), g.normal(0, 0.1, (1, 78)))
    np.save(os.path.join(base, "std.npy"), g.uniform(0.2, 0.6, (1, 78)))
    scene_map, scene_verts, transf = {}, {}, {}
    for s, T in enumerate(lengths):
        rec = f"recording_{s:03d}"
        stamps = [132754997786014666 + 333 * t for t in range(T)]
        imgs = [f"egocentric_color/{rec}/2021-09-07-{s:06d}/PV/{st}_frame_{t:05d}.jpg" for t, st in enumerate(stamps)]

        def person():
            return {"global_orient": g.normal(0, 0.5, (T, 1, 3)), "body_pose": g.normal(0, 0.3, (T, 1, 69)),
                    "betas": np.repeat(g.normal(0, 0.5, (1, 1, 10)), T, axis=0), "transl": g.normal(0, 1.0, (T, 1, 3))}

        item = {"video": [f"{rec}/{t}" for t in range(T)],
                "recording_utils": {"original_imgname": imgs, "fx": list(g.uniform(600, 700, T)), "cx": list(g.uniform(300, 340, T)),
                                    "cy": list(g.uniform(160, 200, T)), "center": g.uniform(100, 500, (T, 2)),
                                    "scale": list(g.uniform(0.5, 2.0, T))},
                "wearer": person(), "interactee": person()}
        np.save(os.path.join(base, split, f"seq_{s:04d}.npy"), item, allow_pickle=True)
        key = f"scene_{s % 2}"
        for im in imgs:
            scene_map[im] = key
        if key not in scene_verts:
            scene_verts[key] = np.stack([g.uniform(-3, 3, n_points), g.uniform(-3, 3, n_points), g.uniform(0.3, 6, n_points)], axis=1)

        def rigid():
            q, _ = np.linalg.qr(g.normal(size=(3, 3)))
            m = np.eye(4)
            m[:3, :3], m[:3, 3] = q * np.sign(np.linalg.det(q)), g.normal(0, 1.0, 3)
            return m

        transf[rec] = {"trans_kinect2holo": rigid(), "trans_world2pv": {str(st): rigid() for st in stamps}}
    pre = os.path.join(root, "Egohmr_scene_preprocess_s1_release")
    with open(os.path.join(pre, f"map_dict_{split}.pkl"), "wb") as f:
        pickle.dump(scene_map, f)
    with open(os.path.join(pre, f"pcd_verts_dict_{split}.pkl"), "wb") as f:
        pickle.dump(scene_verts, f)
    with open(os.path.join(root, "transf_matrices_all_seqs.pkl"), "wb") as f:
        pickle.dump(transf, f)


def write_synthetic_gimo(root: str, split: str = "test", n_seq: int = 3, seed: int = 0, n_scene_points: int = 5000,
                         scene_root: Optional[str] = None) -> None:
    """GIMO-shaped recordings: ``<root>/processed/{mean,std}.npy`` ([1, 69]: 3 + 63 + 3), ``<root>/processed/<split>/*.npy`` and,
    under ``scene_root``, one binary PLY scan + ``transform_norm.txt`` per scene."""
    g = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "processed", split), exist_ok=True)
    np.save(os.path.join(root, "processed", "mean.npy"), g.normal(0, 0.1, (1, 69)))
    np.save(os.path.join(root, "processed", "std.npy"), g.uniform(0.2, 0.6, (1, 69)))
    scene_root = scene_root if scene_root is not None else os.path.join(os.path.dirname(root.rstrip("/")), "gimo_raw", "group", "GIMO")
    T = 60
    for s in range(n_seq):
        scene = f"scene_{s % 2}"
        video = [f"{scene}/2022-01-0{s}/eye_pc/{t}.ply" for t in range(T)]          # [-4] of the path is the scene

        def person():
            return {"global_orient": g.normal(0, 0.5, (T, 1, 3)), "body_pose": g.normal(0, 0.3, (T, 1, 63)),
                    "betas": np.repeat(g.normal(0, 0.5, (1, 1, 10)), T, axis=0), "transl": g.normal(0, 1.0, (T, 1, 3))}

        item = {"video": video,
                "recording_utils": {"fx": list(g.uniform(600, 700, T)), "cx": list(g.uniform(300, 340, T)), "cy": list(g.uniform(160, 200, T)),
                                    "center": g.uniform(100, 500, (T, 2)), "scale": list(g.uniform(0.5, 2.0, T))},
                "wearer": person(), "interactee": person()}
        np.save(os.path.join(root, "processed", split, f"seq_{s:04d}.npy"), item, allow_pickle=True)
        obj = os.path.join(scene_root, scene, "scene_obj")
        if not os.path.exists(os.path.join(obj, "scene_downsampled.ply")):
            os.makedirs(obj, exist_ok=True)
            pts = np.stack([g.uniform(-4, 4, n_scene_points), g.uniform(0, 3, n_scene_points), g.uniform(-4, 4, n_scene_points)], axis=1).astype("<f4")
            with open(os.path.join(obj, "scene_downsampled.ply"), "wb") as f:
                f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                         "end_header\n" % n_scene_points).encode("ascii"))
                f.write(pts.tobytes())
            q, _ = np.linalg.qr(g.normal(size=(3, 3)))
            m = np.eye(4)
            m[:3, :3], m[:3, 3] = q, g.normal(0, 1.0, 3)
            np.savetxt(os.path.join(obj, "transform_norm.txt"), m)

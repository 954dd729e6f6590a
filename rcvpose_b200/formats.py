"""On-disk formats either side of the voting path -- host-side readers only.

    read_depth(path)          reference AccumulatorSpace.py:482-490   LINEMOD `.dpt` (uint32 h, w header + uint16 mm), else an image
    read_ply_points(path)     replaces o3d.io.read_point_cloud(...).points (AccumulatorSpace.py:505, :533)
    read_split(path)          the `Split/val.txt` list (AccumulatorSpace.py:502-503)
    load_keypoints / load_pose  `Outside9.npy` (:535) and `pose/pose<N>.npy` (:566)
    read_xyz_points(path)     YCB-Video `models/<cls>/points.xyz` (AccumulatorSpace.py:988; 3DRadius_ycb.py)
    load_ycb_meta(path)       YCB-Video `<seq>/<frame>-meta.mat`: intrinsics, factor_depth, poses, class indices (:1015-1016, :1050-1057)
    read_ycb_radius_h5(...)   the per-class `<cls>.hdf5` the reference's YCB ground-truth generator writes (3DRadius_ycb.py:200-251):
                              `/3Dradius_pt<k>_dm/<seq>_<frame>` float64 (H, W) radius maps in decimetres, `/JPEGImages/<seq>_<frame>`
                              uint8 RGB.  Needs `h5py` (as the reference does, rmap_dataset.py:7); it is not in this image, so the
                              reader is gated: without h5py it raises ImportError naming the missing module.

The writers exist so that tests and tools can lay out a synthetic dataset in the reference's directory structure.
"""
import os
import struct

import numpy as np


def read_depth(path):
    """LINEMOD_ORIG `.dpt`: two uint32 (h, w) then h*w uint16 depth in mm; any other extension is read as an image.
    Returns the (h, w) array in the file's own integer dtype (AccumulatorSpace.py:482-490)."""
    if path[-3:] == "dpt":
        with open(path, "rb") as f:
            hdr = np.fromfile(f, dtype=np.uint32, count=2)
            if hdr.size != 2:
                raise ValueError("%s: truncated .dpt header" % path)
            h, w = int(hdr[0]), int(hdr[1])
            data = np.fromfile(f, dtype=np.uint16, count=w * h)
            if data.size != w * h:
                raise ValueError("%s: expected %d depth values, file holds %d" % (path, w * h, data.size))
            return data.reshape((h, w))
    from PIL import Image
    return np.asarray(Image.open(path)).copy()


def write_depth_dpt(path, depth):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    with open(path, "wb") as f:
        np.array(depth.shape, dtype=np.uint32).tofile(f)
        depth.tofile(f)


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_points(path):
    """Vertex positions (N,3) float64 of a PLY file (ascii, binary_little_endian or binary_big_endian): what
    `np.asarray(o3d.io.read_point_cloud(path).points)` gives the reference (AccumulatorSpace.py:505, :533).  Other vertex
    properties (normals, colours) and other elements (faces) are skipped."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: PLY header without end_header" % path)
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if not elements:
                    raise ValueError("%s: property before element" % path)
                elements[-1][2].append(tok[1:])
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError("%s: unsupported PLY format %r" % (path, fmt))
        for name, count, props in elements:
            if name != "vertex":
                if any(p[0] == "list" for p in props):
                    raise ValueError("%s: element %r precedes the vertices and has list properties" % (path, name))
                if fmt == "ascii":
                    for _ in range(count):
                        f.readline()
                else:
                    f.seek(count * sum(np.dtype(_PLY_TYPES[p[0]]).itemsize for p in props), os.SEEK_CUR)
                continue
            if any(p[0] == "list" for p in props):
                raise ValueError("%s: list property on vertices" % path)
            names = [p[1] for p in props]
            if not all(a in names for a in "xyz"):
                raise ValueError("%s: vertices without x, y, z" % path)
            if fmt == "ascii":
                rows = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(props)))
                if rows.shape[0] != count:
                    raise ValueError("%s: expected %d vertices, found %d" % (path, count, rows.shape[0]))
                return np.ascontiguousarray(rows[:, [names.index(a) for a in "xyz"]], dtype=np.float64)
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(p[1], end + _PLY_TYPES[p[0]]) for p in props])
            rec = np.fromfile(f, dtype=dt, count=count)
            if rec.size != count:
                raise ValueError("%s: expected %d vertices, file holds %d" % (path, count, rec.size))
            return np.stack([rec[a].astype(np.float64) for a in "xyz"], axis=1)
    raise ValueError("%s: no vertex element" % path)


def write_ply_points(path, xyz, binary=True):
    xyz = np.asarray(xyz)
    with open(path, "wb") as f:
        f.write(("ply\nformat %s 1.0\ncomment rcvpose_b200\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                 "end_header\n" % ("binary_little_endian" if binary else "ascii", len(xyz))).encode("ascii"))
        if binary:
            f.write(b"".join(struct.pack("<3f", *row) for row in xyz))
        else:
            for row in xyz:
                f.write(("%.9g %.9g %.9g\n" % tuple(row)).encode("ascii"))


def read_rgb(path):
    """(H,W,3) uint8 RGB of an image file, as FCResBackbone opens it (AccumulatorSpace.py:140-142: Image.open(...).convert('RGB'))."""
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)


def read_split(path):
    """One frame stem per line (AccumulatorSpace.py:502-503)."""
    with open(path, "r") as f:
        return [s.replace("\n", "") for s in f.readlines()]


def load_keypoints(path):
    """`Outside9.npy`: (9,3) keypoints in metres, object frame; rows 1..3 are the ones voted for (AccumulatorSpace.py:535, :660)."""
    k = np.load(path)
    if k.ndim != 2 or k.shape[1] != 3 or k.shape[0] < 4:
        raise ValueError("%s: expected (>=4, 3) keypoints, got %s" % (path, k.shape))
    return np.asarray(k, dtype=np.float64)


def load_pose(path):
    """`pose<N>.npy`: (3,4) [R | t] with t in metres (AccumulatorSpace.py:566, :665-666)."""
    rt = np.load(path)
    if rt.shape != (3, 4):
        raise ValueError("%s: expected a (3,4) pose, got %s" % (path, rt.shape))
    return np.asarray(rt, dtype=np.float64)


def read_xyz_points(path):
    """`points.xyz`: one "x y z" line per CAD point, metres -> (N,3) float64 (what o3d.io.read_point_cloud gives the YCB evaluator,
    AccumulatorSpace.py:988-990)."""
    pts = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if pts.shape[1] < 3:
        raise ValueError("%s: expected at least 3 columns, got %d" % (path, pts.shape[1]))
    return np.ascontiguousarray(pts[:, :3])


def load_ycb_meta(path):
    """YCB-Video `-meta.mat` (scipy.io.loadmat, AccumulatorSpace.py:1015): dict(intrinsic_matrix (3,3) float64, factor_depth float,
    cls_indexes (M,) int, poses (M,3,4) float64 -- object-major, translation in metres).  `pose_of(meta, class_id)` picks one object
    (:1016)."""
    import scipy.io
    m = scipy.io.loadmat(path)
    for k in ("intrinsic_matrix", "factor_depth", "cls_indexes", "poses"):
        if k not in m:
            raise ValueError("%s: no %r in the meta file" % (path, k))
    poses = np.asarray(m["poses"], dtype=np.float64)
    if poses.ndim == 2:
        poses = poses[:, :, None]
    return dict(intrinsic_matrix=np.asarray(m["intrinsic_matrix"], dtype=np.float64).reshape(3, 3), factor_depth=float(np.ravel(m["factor_depth"])[0]),
                cls_indexes=np.ravel(m["cls_indexes"]).astype(np.int64), poses=np.ascontiguousarray(poses.transpose(2, 0, 1)))


def pose_of(meta, class_id):
    """(3,4) ground-truth pose of object `class_id` in a frame's meta, or None if the object is not in the frame."""
    hit = np.nonzero(meta["cls_indexes"] == class_id)[0]
    return meta["poses"][hit[0]] if len(hit) else None


def read_ycb_radius_h5(path, frame_key, keypoints=(1, 2, 3), with_image=False):
    """Radius maps of one frame from the per-class HDF5 file of the reference's YCB generator (3DRadius_ycb.py:200-251).

    path       `<root_save>/<class>.hdf5`
    frame_key  `<sequence>_<6-digit frame>` (the dataset name the generator uses, :213, :250)
    keypoints  indices k of `/3Dradius_pt<k>_dm` (the evaluators vote for keypoints 1..3 of Outside9.npy)
    Returns (Kp, H, W) float32 radius maps in DECIMETRES (the generator stores Radius3DMap * 10; the voting entry points take
    decimetres with radius_scale = 100 mm, the convention of the networks' output), and the (H, W, 3) uint8 image if asked.
    h5py is imported here, not at module import: the voting path does not depend on it."""
    try:
        import h5py
    except ImportError as e:   # not in this image; the reference needs it too (rmap_dataset.py:7)
        raise ImportError("read_ycb_radius_h5 needs the h5py module (HDF5 radius maps of 3DRadius_ycb.py): %s" % e)
    with h5py.File(path, "r") as f:
        maps = []
        for k in keypoints:
            node = "/3Dradius_pt%d_dm/%s" % (k, frame_key)
            if node not in f:
                raise KeyError("%s: no dataset %s" % (path, node))
            maps.append(np.asarray(f[node], dtype=np.float32))
        out = np.stack(maps, axis=0)
        if with_image:
            return out, np.asarray(f["/JPEGImages/%s" % frame_key], dtype=np.uint8)
    return out

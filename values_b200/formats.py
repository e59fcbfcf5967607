"""On-disk formats of the C2 -> C3 hand-off, without medpy / ITK.

The reference moves every map between its test script and its evaluation through image files:
`medpy.io.save(array, path, hdr)` writes gzip-compressed NIfTI (`.nii.gz`) in
uncertainty_modeling/data_carrier_3D.py:233-371, `medpy.io.load(path)` reads them back in
evaluation/experiment_dataloader.py:38-49, 85-91, 117-160 and aggregate_uncertainties.py:77-79;
the 2D path writes fp32 `.tif` / `.png` with cv2 (uncertainty_modeling/test_2D.py:145-158).

    load(path) -> (array, Header)             medpy.io.load: array indexed [x, y, z] (2D: [W, H])
    save(array, path, hdr=False)              medpy.io.save
    load_to_device(path) -> (tensor, Header)  raw payload uploaded as it lies in the file, axis
    save_from_device(tensor, path, hdr)       order reversed on the GPU (values_reverse_axes)

What is device work here is the axis reversal (the payload is x-fastest; Python indexes [x][y][z]);
gzip and the file system are host work and stay on the host.

Parity: the NIfTI reader / writer follows the NIfTI-1 specification (348-byte header, `n+1`
single-file layout) and ITK's conventions as medpy uses them (LPS <-> RAS sign flip of the first
two axes, qform and sform both written, datatype = the array's dtype, bool -> uint8).  medpy, ITK
and nibabel are absent from the build image and the reference ships no fixture file, so the header
GEOMETRY written here is **parity-unpinned**; the voxel payload (what the hot path consumes) is
pinned by spec-built known-answer files and round trips in tests/test_formats.py.  The cv2 formats
are pinned against cv2 itself.
"""
from __future__ import annotations

import gzip
import os
import struct
import warnings
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

# NIfTI-1 datatype codes <-> numpy
_NIFTI_DTYPES = {
    2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
    256: np.int8, 512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64,
}
_NIFTI_CODES = {np.dtype(v): k for k, v in _NIFTI_DTYPES.items()}
_HDR = struct.Struct("<i10s18sihcB8h3fhhhh8ffffhcBffffii80s24shh6f4f4f4f16s4s")
assert _HDR.size == 348


class Header:
    """The three things medpy's Header carries (medpy/io/header.py): voxel spacing, offset and
    direction cosines, in ITK's LPS convention."""

    def __init__(self, spacing: Optional[Sequence[float]] = None, offset: Optional[Sequence[float]] = None,
                 direction=None):
        self.spacing = None if spacing is None else tuple(float(s) for s in spacing)
        self.offset = None if offset is None else tuple(float(o) for o in offset)
        self.direction = None if direction is None else np.asarray(direction, dtype=np.float64)

    def get_voxel_spacing(self):
        return self.spacing

    def get_offset(self):
        return self.offset

    def get_direction(self):
        return self.direction

    def __repr__(self):
        return f"Header(spacing={self.spacing}, offset={self.offset})"


def _is_nifti(path) -> bool:
    p = str(path).lower()
    return p.endswith(".nii") or p.endswith(".nii.gz")


def _read_bytes(path) -> bytes:
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    return raw


def read_nifti_raw(path) -> Tuple[np.ndarray, Header]:
    """NIfTI-1 single file -> (payload as a C-order array in FILE order, i.e. [..., Z, Y, X], Header).
    Scaling (scl_slope / scl_inter) is applied when the header asks for it, as ITK does."""
    raw = _read_bytes(path)
    if len(raw) < 352:
        raise ValueError(f"{path}: not a NIfTI-1 file (too short)")
    (sizeof_hdr,) = struct.unpack_from("<i", raw, 0)
    end = "<"
    if sizeof_hdr != 348:
        (sizeof_hdr,) = struct.unpack_from(">i", raw, 0)
        end = ">"
        if sizeof_hdr != 348:
            raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    magic = raw[344:348]
    if magic not in (b"n+1\x00", b"ni1\x00"):
        raise ValueError(f"{path}: bad NIfTI magic {magic!r}")
    if magic == b"ni1\x00":
        raise NotImplementedError(f"{path}: two-file NIfTI (.hdr/.img) is not used by the reference")
    dim = struct.unpack_from(end + "8h", raw, 40)
    datatype, bitpix = struct.unpack_from(end + "hh", raw, 70)
    pixdim = struct.unpack_from(end + "8f", raw, 76)
    vox_offset, slope, inter = struct.unpack_from(end + "3f", raw, 108)
    qform_code, sform_code = struct.unpack_from(end + "hh", raw, 252)
    quat = struct.unpack_from(end + "6f", raw, 256)
    srow = np.asarray(struct.unpack_from(end + "12f", raw, 280), dtype=np.float64).reshape(3, 4)
    nd = int(dim[0])
    if not 1 <= nd <= 7:
        raise ValueError(f"{path}: dim[0] = {nd}")
    if datatype not in _NIFTI_DTYPES:
        raise NotImplementedError(f"{path}: NIfTI datatype {datatype}")
    shape_xyz = [int(d) for d in dim[1:1 + nd]]
    while len(shape_xyz) > 1 and shape_xyz[-1] == 1:     # ITK drops trailing singleton axes
        shape_xyz.pop()
    dt = np.dtype(_NIFTI_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape_xyz))
    off = int(vox_offset) if vox_offset >= 352 else 352
    if len(raw) < off + n * dt.itemsize:
        raise ValueError(f"{path}: truncated payload")
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape_xyz[::-1])
    if end == ">":
        data = data.astype(dt.newbyteorder("<"))
    if slope not in (0.0, 1.0) or (slope != 0.0 and inter != 0.0):
        data = data.astype(np.float64) * float(slope) + float(inter)
    k = min(len(shape_xyz), 3)
    spacing = [float(abs(p)) if p != 0 else 1.0 for p in pixdim[1:1 + len(shape_xyz)]]
    # geometry: RAS (NIfTI) -> LPS (ITK): flip the sign of the first two world axes
    flip = np.array([-1.0, -1.0, 1.0])
    if sform_code > 0:
        m = srow[:, :3] * flip[:, None]
        offset = (srow[:, 3] * flip)[:k]
        norm = np.linalg.norm(m, axis=0)
        norm[norm == 0] = 1.0
        direction = (m / norm)[:k, :k]
    elif qform_code > 0:
        b, c, d = quat[:3]
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        rot = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                        [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        if pixdim[0] < 0:
            rot[:, 2] *= -1.0
        direction = (rot * flip[:, None])[:k, :k]
        offset = (np.asarray(quat[3:6]) * flip)[:k]
    else:
        direction, offset = np.eye(k), np.zeros(k)
    return data, Header(spacing, offset, direction)


def _nifti_header_bytes(shape_xyz: Sequence[int], dtype: np.dtype, hdr) -> bytes:
    nd = len(shape_xyz)
    if not 1 <= nd <= 7:
        raise ValueError("NIfTI supports 1 to 7 dimensions")
    code = _NIFTI_CODES.get(np.dtype(dtype))
    if code is None:
        raise TypeError(f"dtype {dtype} has no NIfTI-1 datatype code")
    spacing = [1.0] * nd
    offset = [0.0] * 3
    direction = np.eye(3)
    if hdr:
        sp = hdr.get_voxel_spacing() if hasattr(hdr, "get_voxel_spacing") else None
        if sp is not None:
            for i, v in enumerate(list(sp)[:nd]):
                spacing[i] = float(v)
        of = hdr.get_offset() if hasattr(hdr, "get_offset") else None
        if of is not None:
            for i, v in enumerate(list(of)[:3]):
                offset[i] = float(v)
        dr = hdr.get_direction() if hasattr(hdr, "get_direction") else None
        if dr is not None:
            dr = np.asarray(dr, dtype=np.float64)
            k = min(3, dr.shape[0])
            direction[:k, :k] = dr[:k, :k]
    dim = [nd] + [int(s) for s in shape_xyz] + [1] * (7 - nd)
    sp3 = (spacing + [1.0, 1.0, 1.0])[:3]
    flip = np.array([-1.0, -1.0, 1.0])
    ras = direction * flip[:, None]                        # LPS -> RAS
    srow = np.concatenate([ras * np.asarray(sp3)[None, :], (np.asarray(offset) * flip)[:, None]], axis=1)
    # quaternion of the (proper) rotation; qfac carries a reflection
    rot = ras.copy()
    qfac = 1.0
    if np.linalg.det(rot) < 0:
        rot[:, 2] *= -1.0
        qfac = -1.0
    a = 0.5 * np.sqrt(max(0.0, 1.0 + rot[0, 0] + rot[1, 1] + rot[2, 2]))
    if a > 1e-6:
        b = 0.25 * (rot[2, 1] - rot[1, 2]) / a
        c = 0.25 * (rot[0, 2] - rot[2, 0]) / a
        d = 0.25 * (rot[1, 0] - rot[0, 1]) / a
    else:   # 180 degree rotations (the identity LPS direction is one: diag(-1, -1, 1) -> d = 1)
        xd, yd, zd = 1.0 + rot[0, 0] - rot[1, 1] - rot[2, 2], 1.0 + rot[1, 1] - rot[0, 0] - rot[2, 2], \
            1.0 + rot[2, 2] - rot[0, 0] - rot[1, 1]
        if xd >= yd and xd >= zd:
            b = 0.5 * np.sqrt(xd); c = 0.25 * (rot[0, 1] + rot[1, 0]) / b; d = 0.25 * (rot[0, 2] + rot[2, 0]) / b
        elif yd >= zd:
            c = 0.5 * np.sqrt(yd); b = 0.25 * (rot[0, 1] + rot[1, 0]) / c; d = 0.25 * (rot[1, 2] + rot[2, 1]) / c
        else:
            d = 0.5 * np.sqrt(zd); b = 0.25 * (rot[0, 2] + rot[2, 0]) / d; c = 0.25 * (rot[1, 2] + rot[2, 1]) / d
    pixdim = [qfac] + spacing + [1.0] * (7 - nd)
    itemsize = np.dtype(dtype).itemsize
    return _HDR.pack(
        348, b"", b"", 0, 0, b"r", 0, *dim, 0.0, 0.0, 0.0, 0, code, 8 * itemsize, 0,
        *pixdim, 352.0, 1.0, 0.0, 0, b"\x00", 2 if nd <= 3 else 10, 0.0, 0.0, 0.0, 0.0, 0, 0,
        b"", b"", 1, 1, float(b), float(c), float(d), *(float(v) for v in srow[:, 3]),
        *(float(v) for v in srow[0]), *(float(v) for v in srow[1]), *(float(v) for v in srow[2]),
        b"", b"n+1\x00")


def write_nifti_raw(payload_file_order: np.ndarray, path, hdr=False, compresslevel: int = 1) -> None:
    """payload_file_order: C-contiguous array in FILE order [..., Z, Y, X] (x fastest)."""
    a = np.ascontiguousarray(payload_file_order)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    head = _nifti_header_bytes(a.shape[::-1], a.dtype, hdr) + b"\x00\x00\x00\x00"
    if str(path).lower().endswith(".gz"):
        # no file name, mtime = 0: identical arrays give identical files
        with open(path, "wb") as f, gzip.GzipFile(filename="", fileobj=f, mode="wb", compresslevel=compresslevel, mtime=0) as g:
            g.write(head)
            g.write(a.data)
    else:
        with open(path, "wb") as f:
            f.write(head)
            f.write(a.data)


# ------------------------------------------------------------------ medpy.io signatures (host)
def load(path) -> Tuple[np.ndarray, Header]:
    """medpy.io.load(path) -> (array, header): NIfTI arrays are indexed [x, y, z] (a transposed,
    Fortran-ordered view of the payload, as medpy returns it); cv2 formats come back [W, H]
    (the reference swaps them where it matters, evaluation/metrics/ace.py:21-23)."""
    path = os.fspath(path)
    if not os.path.exists(path):
        raise FileNotFoundError(path)   # medpy raises ImageLoadingError(IOError)
    if _is_nifti(path):
        data, hdr = read_nifti_raw(path)
        return data.T, hdr
    import cv2

    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise IOError(f"Failes to load image {path}")
    if img.ndim == 3:
        img = img[:, :, ::-1]                      # ITK reads RGB, cv2 BGR
        return np.transpose(img, (1, 0, 2)), Header((1.0, 1.0), (0.0, 0.0), np.eye(2))
    return img.T, Header((1.0, 1.0), (0.0, 0.0), np.eye(2))


def save(arr, path, hdr=False, force: bool = True, use_compression: bool = False) -> None:
    """medpy.io.save(arr, filename, hdr=False, force=True): arr indexed [x, y, z] ([W, H])."""
    path = os.fspath(path)
    if not force and os.path.exists(path):
        raise IOError(f"{path} already exists")
    arr = np.asarray(arr)
    if _is_nifti(path):
        write_nifti_raw(np.ascontiguousarray(arr.T), path, hdr)
        return
    import cv2

    img = arr.T if arr.ndim == 2 else np.transpose(arr, (1, 0, 2))[:, :, ::-1]
    if not cv2.imwrite(path, np.ascontiguousarray(img)):
        raise IOError(f"could not write {path}")


# ------------------------------------------------------------------ device forms
def reverse_axes(t: torch.Tensor) -> torch.Tensor:
    """CUDA tensor [n0, n1, n2] (or [n0, n2], [n0]) -> contiguous [n2, n1, n0]: out[c, b, a] = in[a, b, c]."""
    if t.device.type != "cuda":
        raise RuntimeError("reverse_axes expects a CUDA tensor (no CPU fallback)")
    if t.dim() == 1:
        return t.contiguous().clone()
    if t.dim() > 3:
        raise NotImplementedError("reverse_axes supports up to 3 axes")
    t = t.contiguous()
    n0, n2 = t.shape[0], t.shape[-1]
    n1 = t.shape[1] if t.dim() == 3 else 1
    out = torch.empty(tuple(reversed(t.shape)), dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        rc = _lib.lib.values_reverse_axes(t.data_ptr(), out.data_ptr(), t.element_size(), n0, n1, n2,
                                          _lib.stream_ptr(t.device))
    _lib.check(rc)
    return out


_TORCH_OK = {np.dtype(np.uint8), np.dtype(np.int8), np.dtype(np.int16), np.dtype(np.int32), np.dtype(np.int64),
             np.dtype(np.float32), np.dtype(np.float64)}


def load_to_device(path, device: Optional[torch.device] = None) -> Tuple[torch.Tensor, Header]:
    """`load` with the result on the GPU, contiguous and indexed [x, y, z] ([W, H] for cv2 formats):
    the payload is uploaded in file order and the axes are reversed by values_reverse_axes."""
    dev = device or _lib.require_cuda()
    path = os.fspath(path)
    if _is_nifti(path):
        data, hdr = read_nifti_raw(path)
    else:
        import cv2

        data = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if data is None:
            raise IOError(f"Failes to load image {path}")
        hdr = Header((1.0, 1.0), (0.0, 0.0), np.eye(2))
        if data.ndim == 3:
            raise NotImplementedError("colour images are not maps of this path")
    if data.dtype not in _TORCH_OK:   # uint16 / uint32 / uint64 payloads: widen on the host
        data = data.astype(np.int64 if data.dtype.kind in "ui" else np.float64)
    with warnings.catch_warnings():   # the payload is a read-only view of the file's bytes; it is only read
        warnings.simplefilter("ignore", UserWarning)
        t = torch.from_numpy(np.ascontiguousarray(data)).to(dev, non_blocking=True)
    return reverse_axes(t), hdr


def save_from_device(t: torch.Tensor, path, hdr=False) -> None:
    """`save` for a CUDA tensor indexed [x, y, z] ([W, H]): axes reversed on the GPU, one D2H copy."""
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    payload = reverse_axes(t).cpu().numpy()
    path = os.fspath(path)
    if _is_nifti(path):
        write_nifti_raw(payload, path, hdr)
        return
    import cv2

    if not cv2.imwrite(path, payload):
        raise IOError(f"could not write {path}")

"""Minimal NIfTI-1 reader / writer with the slice of the `nibabel` surface that the reference driver uses
(heart_main.py:13,211,223,252,300-303,349-352; utils.py:307): `load(path)` -> image with `.get_data()`, `.get_fdata()`,
`.affine`, `.shape`, `.header`; `Nifti1Image(data, affine)`; `save(img, path)`.  Single-file `.nii` / `.nii.gz`.

Not on the hot path: it exists so that `heart_main.py` can run unmodified where nibabel is not installed (SURVEY.md 8f,
rank 3).  `install_as_nibabel()` registers this module under the name `nibabel` if the real package is absent.
Format reference: the public NIfTI-1 header definition (nifti1.h): 348-byte header, data in Fortran (x-fastest) order.
"""
import gzip
import struct
import sys

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


class Nifti1Header(dict):
    def get_zooms(self):
        return tuple(float(z) for z in self["pixdim"][1:1 + self["ndim"]])

    def get_data_dtype(self):
        return np.dtype(_DTYPES[self["datatype"]])


class Nifti1Image(object):
    def __init__(self, dataobj, affine, header=None):
        self._data = np.asarray(dataobj)
        self.affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64).reshape(4, 4)
        self.header = header if header is not None else Nifti1Header(
            ndim=self._data.ndim, datatype=_CODES.get(self._data.dtype.name, 16),
            pixdim=[1.0] + [float(np.linalg.norm(self.affine[:3, i])) for i in range(3)] + [1.0] * 4, scl_slope=0.0, scl_inter=0.0)

    @property
    def shape(self):
        return self._data.shape

    def get_data(self):
        """array in the stored dtype; scaled to float when the header carries a slope / intercept (nibabel's behaviour)"""
        slope, inter = self.header.get("scl_slope", 0.0), self.header.get("scl_inter", 0.0)
        inter = inter if np.isfinite(inter) else 0.0
        if np.isfinite(slope) and slope != 0.0 and (slope != 1.0 or inter != 0.0):
            return self._data.astype(np.float64) * slope + inter
        return self._data

    def get_fdata(self, dtype=np.float64):
        return np.asarray(self.get_data(), dtype=dtype)

    @property
    def dataobj(self):
        return self._data


def _quaternion_affine(b, c, d, qfac, pixdim, offset):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    zooms = np.array([pixdim[1], pixdim[2], pixdim[3] * (-1.0 if qfac < 0 else 1.0)])
    A = np.eye(4)
    A[:3, :3] = R * zooms
    A[:3, 3] = offset
    return A


def load(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError("%s: shorter than a NIfTI-1 header" % path)
    end = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(end + "i", raw[:4])[0] != 348:
        raise ValueError("%s: not a NIfTI-1 file (sizeof_hdr != 348)" % path)
    if raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError("%s: bad NIfTI-1 magic %r" % (path, raw[344:348]))
    if raw[344:347] == b"ni1":
        raise ValueError("%s: header/image pairs (.hdr/.img) are not supported" % path)
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "2h", raw[70:74])
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(end + "3f", raw[108:120])
    qform, sform = struct.unpack(end + "2h", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(end + "6f", raw[256:280])
    srow = np.array(struct.unpack(end + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype code %d" % (path, datatype))
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise ValueError("%s: bad dim[0] = %d" % (path, ndim))
    shape = tuple(int(v) for v in dim[1:1 + ndim])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    off = int(vox_offset) if vox_offset >= 352 else 352
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    data = data.astype(dt.newbyteorder("="), copy=False)
    if sform > 0:
        affine = np.vstack([srow, [0, 0, 0, 1]])
    elif qform > 0:
        affine = _quaternion_affine(qb, qc, qd, pixdim[0], pixdim, (qx, qy, qz))
    else:
        affine = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    hdr = Nifti1Header(ndim=ndim, datatype=datatype, bitpix=bitpix, pixdim=list(pixdim), scl_slope=float(slope),
                       scl_inter=float(inter), qform_code=qform, sform_code=sform)
    return Nifti1Image(data, affine, hdr)


def save(img, path):
    data = np.asarray(img.dataobj if hasattr(img, "dataobj") else img.get_data())
    if data.dtype.name not in _CODES:
        data = data.astype(np.float32)
    if not 1 <= data.ndim <= 7:
        raise ValueError("NIfTI-1 stores 1..7 dimensions, got %d" % data.ndim)
    affine = np.asarray(img.affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<2h", hdr, 70, _CODES[data.dtype.name], data.dtype.itemsize * 8)
    zooms = [float(np.linalg.norm(affine[:3, i])) or 1.0 for i in range(3)]
    struct.pack_into("<8f", hdr, 76, 1.0, *(zooms + [1.0] * 4))
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)          # vox_offset, scl_slope, scl_inter
    hdr[123] = 2                                                # xyzt_units: millimetres
    struct.pack_into("<2h", hdr, 252, 0, 1)                     # qform_code 0, sform_code 1 (scanner anatomical)
    struct.pack_into("<12f", hdr, 280, *affine[:3].reshape(-1))
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + np.asfortranarray(data).astype(data.dtype.newbyteorder("<"), copy=False).tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)


def install_as_nibabel():
    """make `import nibabel` resolve to this module when the real package is not installed; returns True if it did"""
    try:
        import nibabel  # noqa: F401
        return False
    except ImportError:
        sys.modules["nibabel"] = sys.modules[__name__]
        return True

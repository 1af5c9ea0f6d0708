"""Image I/O with the reference's conversion rules — the data formats either side of the filter calls.

The reference harnesses read and write their frames through ``tools/halide_image_io.h``:
``load_and_convert_image`` (:2757-2780) loads a file in its native element type and converts it to the type of the
buffer on the left-hand side with a fixed rule per type pair (:80-710 — e.g. u8 -> u16 is ``x * 0x0101``, u16 -> u8 is the
rounded divide by 257, u8 -> f32 is ``x / 255.0f``, f32 -> u8 is ``lround(x * 255.0f)`` taken mod 256), and
``convert_and_save_image`` (:2795-2816) picks the closest type / dimensionality the file format can hold
(``best_save_format``, :2447-2472) and converts before saving.  A natural-image run of our filters is only comparable
with the reference harness if the input arrays are byte-identical, so the same rules live here, in numpy, pinned against
the reference header itself (``oracle/_ref/ref_image_io``, built from the header where it lies; ``tests/test_image_io.py``).

Array convention: a Halide image with dimensions (x, y, c, ...) is a C-contiguous numpy array of shape (..., c, y, x) —
the layout ``HalideBuffer.from_numpy`` wraps.  Formats: ``.pgm .ppm .png .npy .tmp .mat`` (8- and 16-bit PNM / PNG,
big-endian on disk like the reference; ``.npy`` and ``.tmp`` and ``.mat`` carry the extents in Halide order, x first,
with a planar x-fastest payload, exactly as the reference writes them; ``.tiff`` is write-only, as in the reference).
Not implemented: ``.jpg``.
PNG pixels are decoded here (zlib + the five row filters); files are written unfiltered, so bytes on disk differ from
libpng's while every sample is identical.

Host-side only: nothing here touches the GPU or the filter library.
"""
import os
import struct
import zlib

import numpy as np

_U = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}
_NAMES = ("bool", "uint8", "uint16", "uint32", "uint64", "int8", "int16", "int32", "int64", "float32", "float64")


# ---- element conversion (tools/halide_image_io.h:80-710) --------------------------------------------------------------
def _as_unsigned(a):
    """Signed integers enter every rule through the implicit C++ conversion to the unsigned type of the same width;
    int32 / int64 sources of the u64 / float / double targets go through uint64 (sign-extended) instead (:322-330, :634-640)."""
    if a.dtype.kind == "i":
        return a.view(_U[a.dtype.itemsize * 8])
    return a


def _lround(x):
    """std::lround / std::llround: nearest, halves away from zero; computed on float64 (exact for float32 inputs)."""
    x = x.astype(np.float64)
    r = np.trunc(x)
    frac = x - r
    r = r + np.where(np.abs(frac) >= 0.5, np.sign(x), 0.0)
    return r.astype(np.int64)


def _to_u8(a):
    k = a.dtype
    if k == np.bool_:
        return a.astype(np.uint8)
    if k.kind in "ui":
        a = _as_unsigned(a)
        bits = a.dtype.itemsize * 8
        if bits == 8:
            return a.copy()
        if bits == 16:
            t = a.astype(np.uint32) + np.uint32(0x80)
            return ((t * np.uint32(255) + np.uint32(255)) >> np.uint32(16)).astype(np.uint8)
        if bits == 64:
            a = (a >> np.uint64(32)).astype(np.uint32)
        return ((a.astype(np.uint64) + np.uint64(0x00808080)) // np.uint64(0x01010101)).astype(np.uint8)
    if k == np.float32:
        return (_lround(a * np.float32(255.0)) & 0xFF).astype(np.uint8)
    return (_lround(a * 255.0) & 0xFF).astype(np.uint8)


def _to_u16(a):
    k = a.dtype
    if k == np.bool_:
        return a.astype(np.uint16)
    if k.kind in "ui":
        a = _as_unsigned(a)
        bits = a.dtype.itemsize * 8
        if bits == 8:
            return a.astype(np.uint16) * np.uint16(0x0101)
        if bits == 16:
            return a.copy()
        if bits == 32:
            return (a >> np.uint32(16)).astype(np.uint16)
        return (a >> np.uint64(48)).astype(np.uint16)
    if k == np.float32:
        return (_lround(a * np.float32(65535.0)) & 0xFFFF).astype(np.uint16)
    return (_lround(a * 65535.0) & 0xFFFF).astype(np.uint16)


def _float_to_u32(a):
    # (uint32_t)std::llround(in * 4294967295.0): the product is a double for float sources too
    return (_lround(a.astype(np.float64) * 4294967295.0) & 0xFFFFFFFF).astype(np.uint32)


def _to_u32(a):
    k = a.dtype
    if k == np.bool_:
        return a.astype(np.uint32)
    if k.kind in "ui":
        a = _as_unsigned(a)
        bits = a.dtype.itemsize * 8
        if bits == 8:
            return a.astype(np.uint32) * np.uint32(0x01010101)
        if bits == 16:
            return a.astype(np.uint32) * np.uint32(0x00010001)
        if bits == 32:
            return a.copy()
        return (a >> np.uint64(32)).astype(np.uint32)
    return _float_to_u32(a)


def _to_u64(a):
    k = a.dtype
    if k == np.bool_:
        return a.astype(np.uint64)
    if k.kind == "i" and a.dtype.itemsize >= 4:
        return a.astype(np.int64).view(np.uint64)  # int32 / int64 -> convert<uint64_t, uint64_t>(in): sign extension (:322-330)
    if k.kind in "ui":
        a = _as_unsigned(a)
        bits = a.dtype.itemsize * 8
        if bits == 8:
            return a.astype(np.uint64) * np.uint64(0x0101010101010101)
        if bits == 16:
            return a.astype(np.uint64) * np.uint64(0x0001000100010001)
        if bits == 32:
            return a.astype(np.uint64) * np.uint64(0x0000000100000001)
        return a.copy()
    return _float_to_u32(a).astype(np.uint64) * np.uint64(0x0000000100000001)


def _to_f32(a):
    k = a.dtype
    if k == np.bool_:
        return a.astype(np.float32)
    if k.kind == "i" and a.dtype.itemsize >= 4:
        a = a.astype(np.int64).view(np.uint64)  # int32 / int64 -> convert<float, uint64_t>(in) (:634-640)
    if k.kind in "ui":
        a = _as_unsigned(a)
        bits = a.dtype.itemsize * 8
        if bits == 8:
            return a.astype(np.float32) / np.float32(255.0)
        if bits == 16:
            return a.astype(np.float32) / np.float32(65535.0)
        if bits == 64:
            a = (a >> np.uint64(32)).astype(np.uint32)
        return (a.astype(np.float64) / 4294967295.0).astype(np.float32)
    return a.astype(np.float32)


def _to_f64(a):
    k = a.dtype
    if k.kind in "ui" and a.dtype.itemsize <= 2:
        return _to_f32(a).astype(np.float64)  # `return in / 255.0f;` — a float division, then widened (:662-668)
    if k == np.bool_:
        return a.astype(np.float64)
    if k.kind == "i" and a.dtype.itemsize >= 4:
        a = a.astype(np.int64).view(np.uint64)
    if k.kind in "ui":
        a = _as_unsigned(a)
        if a.dtype.itemsize == 8:
            a = (a >> np.uint64(32)).astype(np.uint32)
        return a.astype(np.float64) / 4294967295.0
    return a.astype(np.float64)


def convert(arr, dtype):
    """Element-wise ``Internal::convert<To, From>`` of the reference (same value for every input bit pattern; float
    sources outside the representable range of the target and NaNs are undefined there and not reproduced)."""
    a = np.asarray(arr)
    dt = np.dtype(dtype)
    if dt.name not in _NAMES or a.dtype.name not in _NAMES:
        raise TypeError(f"convert: unsupported type pair {a.dtype} -> {dt}")
    if dt == np.bool_:
        return a != 0
    if dt.kind == "f":
        return _to_f32(a) if dt == np.float32 else _to_f64(a)
    bits = dt.itemsize * 8
    u = {8: _to_u8, 16: _to_u16, 32: _to_u32, 64: _to_u64}[bits](a)
    return u.view(dt) if dt.kind == "i" else u  # signed targets: the unsigned result, reinterpreted (:345-550)


# ---- formats ----------------------------------------------------------------------------------------------------------
def _ext(path):
    return os.path.splitext(path)[1].lower().lstrip(".")


def _planar(extents, dtype, payload):
    """numpy view of a planar, x-fastest payload whose extents are in Halide order (x first)."""
    n = int(np.prod(extents))
    a = np.frombuffer(payload, dtype=dtype, count=n)
    return a.reshape(tuple(reversed(extents))).copy()


def _halide_extents(arr):
    return list(reversed(arr.shape))


def _interleaved_to_planar(rows, w, h, channels):
    """rows: [h][w * channels] samples in file order (interleaved) -> (h, w) or (channels, h, w)."""
    if channels == 1:
        return np.ascontiguousarray(rows.reshape(h, w))
    return np.ascontiguousarray(rows.reshape(h, w, channels).transpose(2, 0, 1))


def _planar_to_interleaved(arr):
    """(h, w) or (c, h, w) -> [h][w * c] samples in file order."""
    if arr.ndim == 2:
        return np.ascontiguousarray(arr)
    return np.ascontiguousarray(arr.transpose(1, 2, 0)).reshape(arr.shape[1], -1)


# .pgm / .ppm (:1030-1182): binary P5 / P6, maxval 255 or 65535, 16-bit samples big-endian
def _load_pnm(path, channels):
    data = open(path, "rb").read()
    magic = b"P6" if channels == 3 else b"P5"
    pos = 0

    def token():
        nonlocal pos
        while pos < len(data) and data[pos:pos + 1].isspace():
            pos += 1
        start = pos
        while pos < len(data) and not data[pos:pos + 1].isspace():
            pos += 1
        return data[start:pos]
    if token() != magic:
        raise ValueError(f"{path}: not a {magic.decode()} file")
    w, h, maxval = int(token()), int(token()), int(token())
    pos += 1  # the single whitespace byte after maxval
    if w <= 0 or h <= 0:
        raise ValueError(f"{path}: invalid width or height")
    if maxval not in (255, 65535):
        raise ValueError(f"{path}: invalid bit depth")
    dt = np.dtype(">u2") if maxval == 65535 else np.dtype(np.uint8)
    rows = np.frombuffer(data, dtype=dt, count=w * h * channels, offset=pos).astype(dt.newbyteorder("="))
    return _interleaved_to_planar(rows, w, h, channels)


def _save_pnm(arr, path, channels):
    got = 1 if arr.ndim == 2 else (arr.shape[0] if arr.ndim == 3 else -1)
    if got != channels:
        raise ValueError("Wrong number of channels")
    bits = arr.dtype.itemsize * 8
    h, w = arr.shape[-2], arr.shape[-1]
    rows = _planar_to_interleaved(arr)
    with open(path, "wb") as f:
        f.write(b"%s\n%d %d\n%d\n" % (b"P6" if channels == 3 else b"P5", w, h, (1 << bits) - 1))
        f.write(rows.astype(">u2" if bits == 16 else np.uint8).tobytes())


# .npy (:1183-1500): v1 header, extents in Halide order with 'fortran_order': False over an x-fastest payload
_NPY_CODES = {"f": {2: np.float16, 4: np.float32, 8: np.float64}, "i": {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64},
              "u": {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}}


def _load_npy(path):
    data = open(path, "rb").read()
    if data[:6] != b"\x93NUMPY" or data[6] not in (1, 2, 3) or data[7] != 0:
        raise ValueError(f"{path}: bad .npy magic / version")
    if data[6] == 1:
        hlen, off = struct.unpack_from("<H", data, 8)[0], 10
    else:
        hlen, off = struct.unpack_from("<I", data, 8)[0], 12
    if (off + hlen) % 64 != 0:
        raise ValueError(f"{path}: .npy header is not aligned properly")
    header = data[off:off + hlen].decode("latin1")
    import ast
    d = ast.literal_eval(header.strip())
    descr, shape = d["descr"], tuple(d["shape"])
    if d.get("fortran_order", False) or descr[0] not in "<|":
        raise ValueError(f"{path}: unsupported .npy layout")
    code, nbytes = descr[1], int(descr[2:])
    if code not in _NPY_CODES or nbytes not in _NPY_CODES[code]:
        raise ValueError(f"{path}: unsupported type in load_npy")
    if any(e <= 0 for e in shape):
        raise ValueError(f"{path}: bad extent in .npy file")
    return _planar(list(shape), _NPY_CODES[code][nbytes], data[off + hlen:])


def _save_npy(arr, path):
    code = {"f": "f", "i": "i", "u": "u"}[arr.dtype.kind]
    order = "|" if arr.dtype.itemsize == 1 else "<"
    ext = _halide_extents(arr)
    shape = "(" + ",".join(str(e) for e in ext) + ("," if len(ext) == 1 else "") + ")"
    header = "{'descr': '%s%s%d', 'fortran_order': False, 'shape': %s}\n" % (order, code, arr.dtype.itemsize, shape)
    unpadded = 6 + 2 + 2 + len(header)
    header += " " * (((unpadded + 63) & ~63) - unpadded)
    with open(path, "wb") as f:
        f.write(b"\x93NUMPY\x01\x00" + struct.pack("<H", len(header)) + header.encode("latin1"))
        f.write(np.ascontiguousarray(arr).tobytes())


# .tmp (:1612-1745): five int32 (four extents, type code), planar payload; always four dimensions
_TMP_TYPES = (np.float32, np.float64, np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64)


def _load_tmp(path):
    data = open(path, "rb").read()
    hdr = struct.unpack_from("<5i", data, 0)
    if not (all(e > 0 for e in hdr[:4]) and 0 <= hdr[4] < len(_TMP_TYPES)):
        raise ValueError(f"{path}: bad header on .tmp file")
    return _planar(list(hdr[:4]), _TMP_TYPES[hdr[4]], data[20:])


def _save_tmp(arr, path):
    ext = _halide_extents(arr) + [1] * (4 - arr.ndim)
    codes = [np.dtype(t) for t in _TMP_TYPES]
    if arr.dtype not in codes or arr.ndim > 4:
        raise ValueError("Unsupported type for .tmp file")
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", *ext, codes.index(arr.dtype)))
        f.write(np.ascontiguousarray(arr).tobytes())


# .mat (:1747-2100): MATLAB level 5, one uncompressed numeric matrix
_MI = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 9: np.float64,
       12: np.int64, 13: np.uint64}
_MX_CLASS = {np.dtype(np.float64): 6, np.dtype(np.float32): 7, np.dtype(np.int8): 8, np.dtype(np.uint8): 9,
             np.dtype(np.int16): 10, np.dtype(np.uint16): 11, np.dtype(np.int32): 12, np.dtype(np.uint32): 13,
             np.dtype(np.int64): 14, np.dtype(np.uint64): 15}


def _load_mat(path):
    data = open(path, "rb").read()
    pos = 128
    tag, _ = struct.unpack_from("<2I", data, pos)
    pos += 8
    if tag != 14:
        raise ValueError(f"{path}: bad matrix header")
    flags = struct.unpack_from("<4I", data, pos)
    pos += 16
    if flags[0] != 6 or flags[1] != 8:
        raise ValueError(f"{path}: bad flags")
    shape_tag, shape_bytes = struct.unpack_from("<2I", data, pos)
    pos += 8
    if shape_tag != 5:
        raise ValueError(f"{path}: bad shape header")
    dims = shape_bytes // 4
    extents = list(struct.unpack_from("<%di" % dims, data, pos))
    pos += 4 * dims + (4 if dims & 1 else 0)
    if any(e <= 0 for e in extents):
        raise ValueError(f"{path}: bad extent in .mat file")
    name_tag, name_len = struct.unpack_from("<2I", data, pos)
    pos += 8
    if not (name_tag >> 16):           # long form: the name follows, padded to 8 bytes
        if name_tag != 1:
            raise ValueError(f"{path}: bad name header")
        pos += (name_len + 7) // 8 * 8
    ptype, _ = struct.unpack_from("<2I", data, pos)
    pos += 8
    if ptype not in _MI:
        raise ValueError(f"{path}: unknown payload type")
    return _planar(extents, _MI[ptype], data[pos:])


def _save_mat(arr, path):
    """Byte for byte what the reference writes (:1916-2100): variable name from the file name, `dimensions * 4` shape bytes,
    extents padded with 1 up to two dimensions and with 0 to an even count, payload padded to eight bytes."""
    if arr.dtype not in _MX_CLASS or arr.ndim < 2:
        raise ValueError("Unsupported image for .mat file")
    mi = {v: k for k, v in _MI.items()}[arr.dtype.type]
    payload = np.ascontiguousarray(arr).tobytes()
    if len(payload) >> 32:
        raise ValueError("Buffer too large to save as .mat")
    name = path[:path.rfind(".")] if "." in path else path
    name = name[name.rfind("/") + 1:]
    if not name or not (name[0].isascii() and name[0].isalpha()):
        name = "v" + name
    name = "".join(c if (c.isascii() and c.isalnum()) else "_" for c in name).encode("latin1")
    name_size = len(name)
    name += b"\0" * (-len(name) % 8)
    dims = max(arr.ndim, 2)
    padded_dims = dims + (dims & 1)
    extents = _halide_extents(arr) + [1] * (dims - arr.ndim) + [0] * (padded_dims - dims)
    padding = 7 - ((len(payload) - 1) & 7)
    header = bytearray(b"MATLAB 5.0 MAT-file, produced by Halide".ljust(128))
    header[124:126] = struct.pack("<H", 0x0100)
    header[126:128] = b"IM"
    with open(path, "wb") as f:
        f.write(bytes(header))
        f.write(struct.pack("<2I", 14, 40 + padded_dims * 4 + len(name) + len(payload) + padding))
        f.write(struct.pack("<4I", 6, 8, _MX_CLASS[arr.dtype], 1))
        f.write(struct.pack("<2i", 5, arr.ndim * 4) + struct.pack("<%di" % padded_dims, *extents))
        f.write(struct.pack("<2I", 1, name_size) + name)
        f.write(struct.pack("<2I", mi, len(payload)) + payload + b"\0" * padding)


# .tiff (:2109-2375): the reference only WRITES TIFF (uncompressed, one strip per channel, planar), and cannot read it
def _save_tiff(arr, path):
    if arr.ndim > 4 or arr.dtype.kind not in "iuf":
        raise ValueError("Can only save TIFF files with <= 4 dimensions of an integer or float type")
    ext = _halide_extents(arr) + [1] * (4 - arr.ndim)
    width, height, depth, channels = ext
    if channels in (0, 1) and depth < 5:
        channels, depth = depth, 1
    bpe = arr.dtype.itemsize
    elements = int(np.prod(arr.shape))
    header_size = 210
    tags = []

    def tag16(code, count, value):
        tags.append(struct.pack("<HhiHH", code, 3, count, value & 0xFFFF, 0))

    def tag32(code, count, value, type_code=4):
        tags.append(struct.pack("<Hhii", code, type_code, count, value))
    tag32(256, 1, width)
    tag32(257, 1, height)
    tag16(258, 1, bpe * 8)
    tag16(259, 1, 1)
    tag16(262, 1, 2 if channels >= 3 else 1)
    tag32(273, channels, header_size)
    tag16(277, 1, channels)
    tag32(278, 1, height)
    tag32(279, channels, elements * bpe if channels == 1 else header_size + channels * 4)
    tag32(282, 1, 194, type_code=5)
    tag32(283, 1, 202, type_code=5)
    tag16(284, 1, 1 if channels == 1 else 2)
    tag16(296, 1, 1)
    tag16(339, 1, {"i": 2, "u": 1, "f": 3}[arr.dtype.kind])
    tag32(32997, 1, depth)
    header = struct.pack("<hhih", 0x4949, 42, 8, 15) + b"".join(tags) + struct.pack("<i4i", 0, 1, 1, 1, 1)
    assert len(header) == header_size
    with open(path, "wb") as f:
        f.write(header)
        if channels > 1:
            plane = width * height * depth * bpe
            first = header_size + channels * 4 * 2
            f.write(struct.pack("<%di" % channels, *[first + i * plane for i in range(channels)]))
            f.write(struct.pack("<%di" % channels, *([plane] * channels)))
        f.write(np.ascontiguousarray(arr).tobytes())


def _load_tiff(path):
    raise ValueError("Reading TIFF is not yet supported")   # (the reference's words, tools/halide_image_io.h:2111)


# .png (:867-1030 semantics: 8 / 16-bit samples as stored, no palette expansion, channels from the colour type)
_PNG_CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


def _png_unfilter(raw, h, stride, bpp):
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    pos = 0
    for y in range(h):
        ft = raw[pos]
        line = np.frombuffer(raw, np.uint8, stride, pos + 1).astype(np.int32)
        pos += stride + 1
        if ft == 0:
            cur = line
        elif ft == 1:   # Sub: running sum per byte lane
            cur = np.cumsum(line.reshape(-1, bpp), axis=0).reshape(-1) & 0xFF
        elif ft == 2:   # Up
            cur = (line + prev) & 0xFF
        elif ft in (3, 4):
            cur = np.zeros(stride, np.int32)
            ln, pv = line.tolist(), prev.tolist()
            c = [0] * stride
            for i in range(stride):
                a = c[i - bpp] if i >= bpp else 0
                b = pv[i]
                if ft == 3:
                    c[i] = (ln[i] + ((a + b) >> 1)) & 0xFF
                else:
                    cc = pv[i - bpp] if i >= bpp else 0
                    p = a + b - cc
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - cc)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
                    c[i] = (ln[i] + pred) & 0xFF
            cur = np.array(c, np.int32)
        else:
            raise ValueError("bad PNG filter type")
        out[y] = cur
        prev = cur
    return out


def _load_png(path):
    data = open(path, "rb").read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not recognized as a PNG file")
    pos, idat, ihdr = 8, [], None
    while pos < len(data):
        n, kind = struct.unpack_from(">I4s", data, pos)
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    w, h, depth, ctype, _, _, interlace = ihdr
    if depth not in (8, 16):
        raise ValueError(f"{path}: unsupported PNG bit depth")
    if interlace:
        raise ValueError(f"{path}: interlaced PNG files are not supported")
    channels = _PNG_CHANNELS[ctype]
    bpp = channels * depth // 8
    fast = _load_png_fast(path, w, h, depth, channels)
    if fast is not None:
        return fast
    rows = _png_unfilter(zlib.decompress(b"".join(idat)), h, w * bpp, bpp)
    samples = rows.view(">u2").astype(np.uint16) if depth == 16 else rows
    return _interleaved_to_planar(samples, w, h, channels)


def _load_png_fast(path, w, h, depth, channels):
    """OpenCV decodes the same samples much faster than the row loop above (Average / Paeth rows are sequential); used when
    importable and the file is a plain gray / RGB / RGBA image, cross-checked against the decoder above in the tests."""
    if os.environ.get("HALIDE_B200_PNG_PURE") or channels == 2:
        return None
    try:
        import cv2
    except Exception:
        return None
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None or im.shape[:2] != (h, w) or im.dtype != (np.uint16 if depth == 16 else np.uint8):
        return None
    if im.ndim == 2:
        return im if channels == 1 else None
    if im.shape[2] != channels:
        return None
    im = im[:, :, [2, 1, 0] + ([3] if channels == 4 else [])]  # BGR(A) -> RGB(A)
    return np.ascontiguousarray(im.transpose(2, 0, 1))


def _save_png(arr, path):
    channels = 1 if arr.ndim == 2 else arr.shape[0]
    if arr.ndim not in (2, 3) or not 1 <= channels <= 4 or arr.dtype not in (np.uint8, np.uint16):
        raise ValueError("Can't write this image as PNG")
    h, w = arr.shape[-2], arr.shape[-1]
    depth = arr.dtype.itemsize * 8
    rows = _planar_to_interleaved(arr).astype(">u2" if depth == 16 else np.uint8)
    raw = b"".join(b"\0" + rows[y].tobytes() for y in range(h))

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[channels]
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


_INT_UINT_FLOAT = [np.dtype(t) for t in (np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64,
                                         np.float16, np.float32, np.float64)]
_TMP_SET = [np.dtype(t) for t in _TMP_TYPES]


def save_query(path):
    """The (dtype, dimensions) pairs a format can hold, in the order ``std::set<FormatInfo>`` iterates them (type code
    int < uint < float, then bits, then dimensions): ``best_save_format`` keeps the FIRST best score."""
    e = _ext(path)
    if e == "pgm":
        s = [(np.uint8, 2), (np.uint16, 2)]
    elif e == "ppm":
        s = [(np.uint8, 3), (np.uint16, 3)]
    elif e == "png":
        s = [(np.uint8, 2), (np.uint8, 3), (np.uint16, 2), (np.uint16, 3)]
    elif e == "npy":
        s = [(t, d) for t in _INT_UINT_FLOAT for d in (1, 2, 3, 4)]
    elif e == "tmp":
        s = [(t, 4) for t in _TMP_SET]
    elif e == "mat":
        s = [(t, d) for t in _TMP_SET for d in range(2, 16)]
    elif e in ("tiff",):
        s = [(t, d) for t in _INT_UINT_FLOAT if t != np.float16 for d in (1, 2, 3, 4)]
    else:
        raise ValueError(f'unsupported file extension "{e}"')
    key = lambda td: ({"i": 0, "u": 1, "f": 2}[np.dtype(td[0]).kind], np.dtype(td[0]).itemsize, td[1])
    return sorted(((np.dtype(t), d) for t, d in s), key=key)


def best_save_format(arr, formats):
    """tools/halide_image_io.h:2447-2472 — too few dimensions cost 1024 each, too few bits 8 each, extra bits 1 each, a
    different type code 1; the first format with the lowest score wins."""
    code = {"i": 0, "u": 1, "f": 2, "b": 1}[arr.dtype.kind]
    bits = 1 if arr.dtype == np.bool_ else arr.dtype.itemsize * 8
    best, best_score = None, 0x7FFFFFFF
    for t, d in formats:
        fbits = t.itemsize * 8
        score = max(0, arr.ndim - d) * 1024 + max(0, bits - fbits) * 8 + max(0, fbits - bits)
        score += 1 if {"i": 0, "u": 1, "f": 2}[t.kind] != code else 0
        if score < best_score:
            best, best_score = (t, d), score
    return best


_LOADERS = {"pgm": lambda p: _load_pnm(p, 1), "ppm": lambda p: _load_pnm(p, 3), "png": _load_png, "npy": _load_npy,
            "tmp": _load_tmp, "mat": _load_mat, "tiff": _load_tiff}
_SAVERS = {"pgm": lambda a, p: _save_pnm(a, p, 1), "ppm": lambda a, p: _save_pnm(a, p, 3), "png": _save_png,
           "npy": _save_npy, "tmp": _save_tmp, "mat": _save_mat, "tiff": _save_tiff}


def load(path):
    """``Halide::Tools::load``: the file's own element type and dimensionality."""
    e = _ext(path)
    if e not in _LOADERS:
        raise ValueError(f'unsupported file extension "{e}"')
    return _LOADERS[e](path)


def load_and_convert_image(path, dtype):
    """``Buffer<T> im = load_and_convert_image(path)`` (:2757-2780)."""
    a = load(path)
    return a if a.dtype == np.dtype(dtype) else convert(a, dtype)


def save(arr, path):
    """``Halide::Tools::save`` (:2690-2706): fails unless the format holds this type and dimensionality exactly."""
    arr = np.asarray(arr)
    if (arr.dtype, arr.ndim) not in save_query(path):
        raise ValueError("Image cannot be saved in this format")
    _SAVERS[_ext(path)](arr, path)


def convert_and_save_image(arr, path):
    """``convert_and_save_image`` (:2795-2816): convert to the closest type the format holds, add trailing dimensions of
    extent 1 if the format needs more (``.tmp`` is always four-dimensional), then save."""
    arr = np.asarray(arr)
    t, d = best_save_format(arr, save_query(path))
    if t != arr.dtype:
        arr = convert(arr, t)
    while arr.ndim < d:
        arr = arr[np.newaxis]   # Halide add_dimension(): a new OUTERMOST dimension of extent 1
    _SAVERS[_ext(path)](arr, path)

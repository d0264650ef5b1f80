"""Minimal scanline OpenEXR writer / reader (host side of SampleWriter, src/rlUtil.h:44-96).

The reference's debug writer hands three FLOAT planes named B, G, R to the vendored tinyexr
(`SaveMultiChannelEXRToFile`, src/ext/tinyexr.h:122) and asks for HALF on disk
(src/rlUtil.h:63-64).  This module writes the same image as an uncompressed scanline
OpenEXR 2.0 file (compression NO_COMPRESSION instead of tinyexr's ZIP: any EXR reader
decodes both), and reads such files back for the tests.  Pure numpy; no CUDA involved.
"""
import struct

import numpy as np

_MAGIC = 20000630
_PIXEL_TYPES = {1: np.float16, 2: np.float32}


def _attr(name, typ, payload):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_scanline_exr(path, planes, channel_names, half=True):
    """planes: [C, H, W] float array; channel_names: C names in plane order."""
    planes = np.asarray(planes)
    c, h, w = planes.shape
    if len(channel_names) != c:
        raise ValueError("one name per plane")
    order = sorted(range(c), key=lambda i: channel_names[i])       # the chlist is alphabetical
    ptype, dtype = (1, np.float16) if half else (2, np.float32)
    chlist = b"".join(channel_names[i].encode() + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1) for i in order) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = b"".join([
        _attr("channels", "chlist", chlist),
        _attr("compression", "compression", b"\0"),
        _attr("dataWindow", "box2i", box),
        _attr("displayWindow", "box2i", box),
        _attr("lineOrder", "lineOrder", b"\0"),
        _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)),
        _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)),
    ]) + b"\0"
    head = struct.pack("<ii", _MAGIC, 2) + header
    line_bytes = c * w * np.dtype(dtype).itemsize
    first = len(head) + 8 * h
    offsets = struct.pack("<%dQ" % h, *[first + y * (8 + line_bytes) for y in range(h)])
    data = np.ascontiguousarray(planes[order].astype(dtype).transpose(1, 0, 2))     # [H, C, W]
    with open(path, "wb") as fh:
        fh.write(head)
        fh.write(offsets)
        for y in range(h):
            fh.write(struct.pack("<ii", y, line_bytes))
            fh.write(data[y].tobytes())
    return path


def read_scanline_exr(path):
    """Reads an uncompressed scanline EXR written by write_scanline_exr: ({name: [H, W]}, attrs)."""
    buf = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", buf, 0)
    if magic != _MAGIC or (version & 0xff) != 2:
        raise ValueError("not an OpenEXR 2.0 file")
    pos, attrs = 8, {}
    while buf[pos] != 0:
        end = buf.index(b"\0", pos); name = buf[pos:end].decode(); pos = end + 1
        end = buf.index(b"\0", pos); typ = buf[pos:end].decode(); pos = end + 1
        size, = struct.unpack_from("<i", buf, pos); pos += 4
        attrs[name] = (typ, buf[pos:pos + size]); pos += size
    pos += 1
    if attrs["compression"][1] != b"\0":
        raise ValueError("only NO_COMPRESSION files are supported")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans, raw, p = [], attrs["channels"][1], 0
    while raw[p] != 0:
        end = raw.index(b"\0", p); nm = raw[p:end].decode(); p = end + 1
        ptype, = struct.unpack_from("<i", raw, p); p += 16
        chans.append((nm, _PIXEL_TYPES[ptype]))
    offsets = struct.unpack_from("<%dQ" % h, buf, pos)
    out = {nm: np.zeros((h, w), dtype=dt) for nm, dt in chans}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        q = off + 8
        for nm, dt in chans:
            nbytes = w * np.dtype(dt).itemsize
            out[nm][y - y0] = np.frombuffer(buf, dtype=dt, count=w, offset=q)
            q += nbytes
    return out, attrs

"""Host mirror of the export side of js/planet-mesh.js: exportMap / exportFilename (:1752-1961).

The reference builds one map triangle per mesh side, renders them through WebGL tiles, applies the sRGB curve and lets the
browser encode the canvas (`canvas.toBlob(…, 'image/png')`).  Here the triangles, the rasterisation and the sRGB step are
CUDA kernels behind `pb_export_map` (csrc/pb_export.h); this module maps the reference's export type names onto the C ABI's
colour modes and writes the PNG container the browser would.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from .engine import DeviceMesh

# export type (js/planet-mesh.js:1779-1793) → PB_COLOR_* mode
_EXPORT_MODE = {"landmask": 4, "landheightmap": 3, "heightmap": 2, "biome": 1, "koppen": 6}


def exportFilename(type: str, seed) -> str:
    """js/planet-mesh.js:1952-1961"""
    return {
        "landmask": f"orogen-landmask-{seed}.png",
        "landheightmap": f"orogen-land-heightmap-{seed}.png",
        "heightmap": f"orogen-heightmap-{seed}.png",
        "biome": f"orogen-satellite-{seed}.png",
        "koppen": f"orogen-climate-{seed}.png",
    }.get(type, f"orogen-colormap-{seed}.png")


def exportMapPixels(mesh: DeviceMesh, type: str, width: int, r_elevation, r_koppen=None, want_sides: bool = False):
    """The ImageData of exportMap(type, width): uint8[height, width, 4] (top row = north), height = width / 2.

    `r_koppen` is `debugLayers.koppen`; like the reference, the 'biome' and 'koppen' types fall back to the elevation colour
    map when no climate has been computed (:1764-1765, 1788-1793).  numpy arrays in → numpy out; torch.cuda tensors in →
    torch.cuda tensors out (the pixels stay in HBM).  With want_sides also the int32[height, width] side index that owns
    every pixel (-1 = background)."""
    width = int(width)
    height = width // 2
    mode = _EXPORT_MODE.get(type, 0)
    if mode in (1, 6) and r_koppen is None:
        mode = 0
    n = mesh.numRegions
    rgba = mesh._new(r_elevation, "u8", 4 * width * height)
    sides = mesh._new(r_elevation, "i32", width * height) if want_sides else None
    mesh._begin(r_elevation, r_koppen, rgba, sides)
    mesh.lib.check(mesh.lib.dll.pb_export_map(
        mesh._mesh, mode, width, mesh._ptr(r_elevation, "f32", n, "r_elevation"),
        None if mode not in (1, 6) else mesh._ptr(r_koppen, "u8", n, "r_koppen"),
        mesh._ptr(rgba, "u8", 4 * width * height, "rgba"),
        None if sides is None else mesh._ptr(sides, "i32", width * height, "pixelSide")))
    rgba = rgba.reshape(height, width, 4)
    return (rgba, sides.reshape(height, width)) if want_sides else rgba


def encode_png(rgba: np.ndarray, level: int = 6) -> bytes:
    """8-bit RGBA PNG (colour type 6, filter 0 on every row) — what `canvas.toBlob(cb, 'image/png')` delivers."""
    a = np.ascontiguousarray(rgba, np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("rgba must be uint8[height, width, 4]")
    h, w = a.shape[:2]
    raw = np.empty((h, 1 + 4 * w), np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = a.reshape(h, 4 * w)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + chunk(b"IEND", b""))


def decode_png(data: bytes) -> np.ndarray:
    """Inverse of encode_png for the subset it writes (8-bit RGBA, filter 0) — used by the round-trip tests."""
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG")
    pos, w, h, idat = 8, 0, 0, b""
    while pos < len(data):
        (n,), tag = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        if crc != zlib.crc32(tag + body) & 0xFFFFFFFF:
            raise ValueError(f"bad CRC in {tag!r}")
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if (depth, ctype, interlace) != (8, 6, 0):
                raise ValueError("only 8-bit non-interlaced RGBA is supported")
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    if raw[:, 0].any():
        raise ValueError("only filter type 0 is supported")
    return raw[:, 1:].reshape(h, w, 4).copy()


def exportMap(mesh: DeviceMesh, type: str, width: int, r_elevation, r_koppen=None, seed="") -> tuple[str, bytes]:
    """exportMap(type, width) → (download filename, PNG bytes)."""
    px = exportMapPixels(mesh, type, width, r_elevation, r_koppen)
    if not isinstance(px, np.ndarray):
        px = px.cpu().numpy()
    return exportFilename(type, seed), encode_png(px)

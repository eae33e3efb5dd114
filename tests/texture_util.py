"""Helpers of the texture tests: tiny image writers (PNG through zlib, TGA, PFM), a textured test scene and a
reader of the flat file's texture section (csrc/host/flat_io.cpp, IILEFLT3)."""
import struct
import zlib

import numpy as np


def write_png(path, rgb8, filters=(0, 1, 2, 3, 4)):
    """8-bit RGB (h, w, 3) or grey (h, w) PNG; scanline filters cycle through `filters` so that the reader's
    defiltering is exercised."""
    a = np.asarray(rgb8, np.uint8)
    grey = a.ndim == 2
    h, w = a.shape[:2]
    bpp = 1 if grey else 3
    rows = a.reshape(h, w * bpp).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * bpp, np.int32)
    for y in range(h):
        f = filters[y % len(filters)]
        cur = rows[y]
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        upleft = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if f == 0:
            pred = np.zeros_like(cur)
        elif f == 1:
            pred = left
        elif f == 2:
            pred = prev
        elif f == 3:
            pred = (left + prev) // 2
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
        raw.append(f)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if grey else 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(bytes(raw), 6)))
        f.write(chunk(b"IEND", b""))


def write_tga(path, rgb8):
    a = np.asarray(rgb8, np.uint8)
    h, w = a.shape[:2]
    with open(path, "wb") as f:
        f.write(struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, w, h, 24, 0x20))   # top-to-bottom
        f.write(a[..., ::-1].tobytes())


def write_pfm(path, rgb):
    a = np.asarray(rgb, np.float32)
    h, w = a.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (w, h))
        f.write(a[::-1].astype("<f4").tobytes())


def test_image(w, h, seed=0):
    """A smooth pattern plus noise: gradients exercise the filtering, noise the exactness."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(x * 0.35) * np.cos(y * 0.2), 40 + 200 * ((x // 4 + y // 4) % 2),
                    255 * x / max(w - 1, 1)], axis=-1)
    img = np.clip(img + rng.integers(-20, 20, img.shape), 0, 255)
    return img.astype(np.uint8)


TEX_DTYPE = np.dtype([("type", "<i4"), ("channels", "<i4"), ("value", "<f4", 3), ("tex1", "<i4"), ("tex2", "<i4"),
                      ("image", "<i4"), ("su", "<f4"), ("sv", "<f4"), ("du", "<f4"), ("dv", "<f4"), ("wrap", "<i4"),
                      ("trilinear", "<i4"), ("maxAnisotropy", "<f4"), ("pad", "<f4")])
IMG_DTYPE = np.dtype([("width", "<i4"), ("height", "<i4"), ("nLevels", "<i4"), ("channels", "<i4"),
                      ("texelOffset", "<u8"), ("pad", "<u8")])


def read_flat_textures(path, setup_bytes=520):
    """(textures, images, texels, materialTextures) of an IILEFLT3 file."""
    b = open(path, "rb").read()
    assert b[:8] == b"IILEFLT3", b[:8]
    c = np.frombuffer(b, "<u4", 8, 8)
    n_verts, has_n, has_uv, n_prims, n_spheres, n_mat, n_lights, n_nodes = [int(v) for v in c]
    at = 8 + 32 + setup_bytes
    at += n_verts * 12 * (1 + has_n) + n_verts * 8 * has_uv
    at += n_prims * 32 + n_spheres * 160 + n_mat * 96 + n_lights * 272 + n_nodes * 32
    at += n_prims * 4 if n_nodes else 0
    t = np.frombuffer(b, "<u8", 4, at)
    at += 32
    tex = np.frombuffer(b, TEX_DTYPE, int(t[0]), at)
    at += int(t[0]) * 64
    img = np.frombuffer(b, IMG_DTYPE, int(t[1]), at)
    at += int(t[1]) * 32
    texels = np.frombuffer(b, "<f4", int(t[2]), at)
    at += int(t[2]) * 4
    mt = np.frombuffer(b, "<i4", int(t[3]) * 12, at).reshape(-1, 12)
    return tex, img, texels, mt


def pyramid_levels(img, texels):
    out, at = [], int(img["texelOffset"])
    w, h, c = int(img["width"]), int(img["height"]), int(img["channels"])
    for _ in range(int(img["nLevels"])):
        out.append(texels[at:at + w * h * c].reshape(h, w, c))
        at += w * h * c
        w, h = max(1, w // 2), max(1, h // 2)
    return out


def read_ref_pyramid(path, channels):
    b = open(path, "rb").read()
    n = struct.unpack_from("<i", b, 0)[0]
    at, out = 4, []
    for _ in range(n):
        w, h = struct.unpack_from("<ii", b, at)
        at += 8
        out.append(np.frombuffer(b, "<f4", w * h * channels, at).reshape(h, w, channels))
        at += w * h * channels * 4
    return out

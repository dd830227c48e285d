"""TEST INFRASTRUCTURE ONLY.  A PNG decoder / encoder written from the PNG specification (ISO/IEC 15948: chunk layout, zlib stream, the five scanline
filters), numpy + zlib only -- no imaging library.

Why it exists: the reference decodes depth maps with `pyspng.load` (src/training/dataset.py:310-323); pyspng is a third-party binding of libspng that is
not installed in this image (environment.yml lists it without a pin).  PNG is lossless, so every conforming decoder returns the same samples; pyspng
returns them as an [h, w, c] (c > 1) or [h, w] array of uint8 / native-endian uint16.  `load()` below has that contract and is injected as the `pyspng`
module when oracle/make_dataset_golden.py runs the UNMODIFIED reference reader on the depth fixture, so that the product's PIL-based decode
(3dgp_b200/training/dataset.py::_decode_depth) is pinned against reference logic + an independent decoder rather than against itself."""
import struct
import zlib

import numpy as np

_SIG = b'\x89PNG\r\n\x1a\n'
_CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def load(data):
    """bytes of a non-interlaced, non-palette PNG -> ndarray [h, w] (one channel) or [h, w, c], uint8 or uint16."""
    assert data[:8] == _SIG, 'not a PNG'
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        n, kind = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(kind + body) == struct.unpack('>I', data[pos + 8 + n:pos + 12 + n])[0], 'chunk CRC'
        if kind == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', body)
        elif kind == b'IDAT':
            idat.append(body)
        elif kind == b'IEND':
            break
        pos += 12 + n
    w, h, depth, ctype, _comp, _filt, interlace = hdr
    assert depth in (8, 16) and ctype in _CHANNELS and interlace == 0, 'unsupported PNG flavour'
    ch = _CHANNELS[ctype]
    bpp = ch * depth // 8                      # bytes per complete pixel: the distance the filters look back
    stride = w * bpp
    raw = zlib.decompress(b''.join(idat))
    assert len(raw) == h * (stride + 1)
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        ft = raw[y * (stride + 1)]
        line = np.frombuffer(raw, dtype=np.uint8, count=stride, offset=y * (stride + 1) + 1).astype(np.int32)
        cur = np.zeros(stride, dtype=np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                pred = a if ft == 1 else ((a + b) >> 1 if ft == 3 else _paeth(int(a), int(b), int(c)))
                cur[i] = (line[i] + pred) & 255
        out[y] = cur
        prev = cur
    if depth == 16:
        px = out.reshape(h, w, ch, 2).astype(np.uint16)
        arr = (px[..., 0] << 8) | px[..., 1]    # network byte order in the file
    else:
        arr = out.reshape(h, w, ch)
    return arr[:, :, 0] if ch == 1 else arr


def save(arr, filter_type=0):
    """[h, w] or [h, w, c] uint8 / uint16 -> PNG bytes (one IDAT; every scanline with the same filter type: 0 none, 1 sub, 2 up, 3 average, 4 Paeth)."""
    arr = np.asarray(arr)
    a3 = arr[:, :, None] if arr.ndim == 2 else arr
    h, w, ch = a3.shape
    ctype = {v: k for k, v in _CHANNELS.items()}[ch]
    depth = 16 if arr.dtype == np.uint16 else 8
    assert arr.dtype in (np.uint8, np.uint16)
    rows = a3.astype('>u2').view(np.uint8).reshape(h, -1) if depth == 16 else a3.reshape(h, -1)
    bpp = ch * depth // 8
    rows = rows.astype(np.int32)
    lines = []
    for y in range(h):
        cur = rows[y]
        if filter_type == 1:
            f = cur.copy(); f[bpp:] = (cur[bpp:] - cur[:-bpp]) & 255
        elif filter_type == 2:
            f = (cur - (rows[y - 1] if y else 0)) & 255
        elif filter_type in (3, 4):
            up = rows[y - 1] if y else np.zeros_like(cur)
            f = cur.copy()
            for i in range(len(cur)):
                a = cur[i - bpp] if i >= bpp else 0
                c = up[i - bpp] if i >= bpp else 0
                f[i] = (cur[i] - ((a + up[i]) >> 1 if filter_type == 3 else _paeth(int(a), int(up[i]), int(c)))) & 255
        else:
            f = cur
        lines.append(bytes([filter_type]) + f.astype(np.uint8).tobytes())

    def chunk(kind, body):
        return struct.pack('>I', len(body)) + kind + body + struct.pack('>I', zlib.crc32(kind + body))
    return _SIG + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, depth, ctype, 0, 0, 0)) + chunk(b'IDAT', zlib.compress(b''.join(lines), 6)) + chunk(b'IEND', b'')

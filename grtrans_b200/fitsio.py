"""Minimal FITS reader for grtrans output files (the image has neither astropy nor cfitsio).

Supports what camera.f90:156-344 writes: a primary float32 vector (pixel locations) and one IMAGE extension per camera
with BITPIX=-32, NAXIS=1.  `HDU.header.get(i)` returns the value of the i-th card like astropy's index lookup, which is
how grtrans_batch.read_grtrans_output finds nx, ny, nu (cards 6, 7, 8)."""
import numpy as np


class Header(object):
    def __init__(self, cards):
        self.cards = cards  # list of (key, value)

    def get(self, k, default=None):
        if isinstance(k, int):
            return self.cards[k][1] if k < len(self.cards) else default
        for key, v in self.cards:
            if key == k.upper():
                return v
        return default

    def keys(self):
        return [c[0] for c in self.cards]


class HDU(object):
    def __init__(self, header, data):
        self.header = header
        self.data = data


def _parse_value(s):
    s = s.split('/')[0].strip() if not s.strip().startswith("'") else s.strip()
    if s.startswith("'"):
        return s[1:s.index("'", 1)].rstrip()
    if s in ('T', 'F'):
        return s == 'T'
    try:
        return int(s)
    except ValueError:
        return float(s.replace('D', 'E'))


def open(path):  # noqa: A001  (mirrors astropy.io.fits.open)
    import builtins
    with builtins.open(path, 'rb') as f:
        raw = f.read()
    hdus = []
    pos = 0
    while pos < len(raw):
        cards = []
        end = False
        while not end:
            block = raw[pos:pos + 2880].decode('ascii')
            pos += 2880
            for i in range(36):
                c = block[80 * i:80 * i + 80]
                key = c[:8].strip()
                if key == 'END':
                    end = True
                    break
                if key == 'HIERARCH':
                    k, v = c[9:].split('=', 1)
                    cards.append((k.strip(), _parse_value(v)))
                elif c[8:10] == '= ':
                    cards.append((key, _parse_value(c[10:])))
        h = Header(cards)
        n = h.get('NAXIS1', 0) if h.get('NAXIS', 0) else 0
        nbytes = 4 * n
        data = np.frombuffer(raw, dtype='>f4', count=n, offset=pos).astype(np.float32)
        pos += ((nbytes + 2879) // 2880) * 2880
        hdus.append(HDU(h, data))
    return hdus

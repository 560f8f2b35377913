"""Minimal FITS binary-table reader (no astropy).

The reference reads its NIKA beam and transfer-function files with
``astropy.io.fits.open(f)[''].data[0]`` (reference ``joxsz_funcs.py:22-23``):
the first (unnamed) BINTABLE extension, row 0, whose fields are fixed-length
vectors.  astropy is not a dependency of this package, so the few FITS
features those files use are decoded here directly:

* 2880-byte blocks of 80-character header cards,
* one empty primary HDU followed by BINTABLE extensions,
* TFORMn of the form ``rT`` with T in ``L B I J K E D A`` (big-endian).

Run once at set-up time on the host; never on the likelihood path.
"""
from __future__ import annotations

import numpy as np

_BLOCK = 2880
_CARD = 80

# FITS TFORM letter -> numpy big-endian dtype
_TFORM_DTYPES = {
    "L": "S1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8",
    "E": ">f4", "D": ">f8", "A": "S1",
}


def _parse_header(buf: bytes, offset: int):
    """Return (cards dict, offset of the first byte after the header blocks)."""
    cards = {}
    pos = offset
    while True:
        if pos + _CARD > len(buf):
            raise ValueError("FITS header runs past end of file (no END card)")
        card = buf[pos:pos + _CARD].decode("ascii", "replace")
        pos += _CARD
        key = card[:8].strip()
        if key == "END":
            break
        if card[8:10] != "= ":
            continue  # COMMENT / HISTORY / blank
        body = card[10:]
        if body.lstrip().startswith("'"):
            start = body.index("'") + 1
            end = start
            while True:  # a doubled quote is an escaped quote
                end = body.index("'", end)
                if body[end:end + 2] == "''":
                    end += 2
                    continue
                break
            value = body[start:end].replace("''", "'").rstrip()
        else:
            text = body.split("/", 1)[0].strip()
            if text in ("T", "F"):
                value = text == "T"
            else:
                try:
                    value = int(text)
                except ValueError:
                    try:
                        value = float(text.replace("D", "E"))
                    except ValueError:
                        value = text
        cards[key] = value
    # header occupies whole blocks
    pos = offset + -(-(pos - offset) // _BLOCK) * _BLOCK
    return cards, pos


def _data_nbytes(cards) -> int:
    naxis = int(cards.get("NAXIS", 0))
    if naxis == 0:
        return 0
    n = abs(int(cards["BITPIX"])) // 8
    for i in range(1, naxis + 1):
        n *= int(cards[f"NAXIS{i}"])
    n *= int(cards.get("GCOUNT", 1))
    n += int(cards.get("PCOUNT", 0))
    return n


def _split_tform(tform: str):
    tform = tform.strip()
    i = 0
    while i < len(tform) and tform[i].isdigit():
        i += 1
    repeat = int(tform[:i]) if i else 1
    letter = tform[i]
    if letter not in _TFORM_DTYPES:
        raise NotImplementedError(f"TFORM {tform!r} not supported by this reader")
    return repeat, letter


class BinTableRow:
    """One table row; fields by position (``row[i]``, ``row[:n]``) or by name."""

    def __init__(self, names, values):
        self.names = list(names)
        self._values = list(values)

    def __len__(self):
        return len(self._values)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._values[self.names.index(key)]
        return self._values[key]

    def __iter__(self):
        return iter(self._values)


def read_bintable(filename: str, ext: int = 1):
    """Decode BINTABLE extension number ``ext`` (1 = first extension).

    Returns a list of :class:`BinTableRow`, native-endian numpy arrays per field
    (scalars for repeat count 1).
    """
    with open(filename, "rb") as fh:
        buf = fh.read()
    if buf[:6] != b"SIMPLE":
        raise ValueError(f"{filename}: not a FITS file")
    pos = 0
    hdu = 0
    while pos < len(buf):
        cards, data_start = _parse_header(buf, pos)
        nbytes = _data_nbytes(cards)
        if hdu == ext:
            if cards.get("XTENSION") != "BINTABLE":
                raise ValueError(f"{filename}: HDU {ext} is not a BINTABLE")
            return _decode_table(cards, buf[data_start:data_start + nbytes])
        pos = data_start + -(-nbytes // _BLOCK) * _BLOCK
        hdu += 1
    raise ValueError(f"{filename}: no HDU {ext}")


def _decode_table(cards, data: bytes):
    row_len = int(cards["NAXIS1"])
    nrows = int(cards["NAXIS2"])
    nfields = int(cards["TFIELDS"])
    names, layout = [], []
    off = 0
    for i in range(1, nfields + 1):
        repeat, letter = _split_tform(str(cards[f"TFORM{i}"]))
        dt = np.dtype(_TFORM_DTYPES[letter])
        names.append(str(cards.get(f"TTYPE{i}", f"COL{i}")).strip())
        layout.append((off, repeat, letter, dt))
        off += repeat * dt.itemsize
    if off != row_len:
        raise ValueError(f"row length mismatch: TFORMs give {off}, NAXIS1 says {row_len}")
    rows = []
    for r in range(nrows):
        base = r * row_len
        vals = []
        for (o, repeat, letter, dt) in layout:
            raw = np.frombuffer(data, dtype=dt, count=repeat, offset=base + o)
            if letter == "A":
                vals.append(b"".join(raw.tolist()).decode("ascii", "replace").rstrip())
                continue
            if letter == "L":
                arr = raw == b"T"
            else:
                arr = raw.astype(dt.newbyteorder("="))
            vals.append(arr[0] if repeat == 1 else arr)
        rows.append(BinTableRow(names, vals))
    return rows

"""Image files of the session set-up cache (ref: io/image.py:13-103): `write_image` / `read_tiff_image` with the reference's
intensity scaling and its `scale_factor` entry in the TIFF ImageDescription.

tifffile is not part of this image, so the TIFF container is written and parsed here: baseline little-endian TIFF, one
uncompressed grey-scale strip, ImageDescription = the JSON tifffile writes for `metadata=` ({"shape": [...], **metadata}).
`read_tiff_image` also reads multi-strip uncompressed files, which is what tifffile produces for the reference's default
`compress=0`; compressed TIFFs raise.  Host-side, once per session: not on the hot path.
"""
from __future__ import annotations

import ast
import json
import os
import struct
from typing import Optional, Tuple, Union

import numpy as np

_TYPES = {1: 'B', 2: 'c', 3: 'H', 4: 'I', 16: 'Q'}


def _write_tiff(path: str, image: np.ndarray, description: str) -> None:
    img = np.ascontiguousarray(image)
    if img.ndim != 2 or img.dtype not in (np.uint8, np.uint16):
        raise ValueError(f'_write_tiff: 2-D uint8 / uint16 images only (got {img.dtype}, {img.shape})')
    desc = description.encode('utf-8') + b'\x00'
    h, w = img.shape
    bits = img.dtype.itemsize * 8
    entries = 10
    ifd_offset = 8
    desc_offset = ifd_offset + 2 + entries * 12 + 4
    data_offset = (desc_offset + len(desc) + 7) // 8 * 8
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, bits), (259, 3, 1, 1), (262, 3, 1, 1), (270, 2, len(desc), desc_offset),
            (273, 4, 1, data_offset), (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, img.nbytes)]
    with open(path, 'wb') as fh:
        fh.write(b'II' + struct.pack('<HI', 42, ifd_offset))
        fh.write(struct.pack('<H', entries))
        for tag, typ, count, value in tags:
            fh.write(struct.pack('<HHI', tag, typ, count))
            fh.write(struct.pack('<HH', value, 0) if typ == 3 else struct.pack('<I', value))
        fh.write(struct.pack('<I', 0))
        fh.write(desc)
        fh.write(b'\x00' * (data_offset - desc_offset - len(desc)))
        fh.write(img.astype('<u2' if bits == 16 else np.uint8).tobytes())


def _read_tiff(path: str) -> Tuple[np.ndarray, str]:
    with open(path, 'rb') as fh:
        buf = fh.read()
    if buf[:2] not in (b'II', b'MM'):
        raise ValueError(f'{path}: not a TIFF file')
    e = '<' if buf[:2] == b'II' else '>'
    magic, ifd = struct.unpack(e + 'HI', buf[2:8])
    if magic != 42:
        raise ValueError(f'{path}: BigTIFF / unknown TIFF flavour is not supported')
    n, = struct.unpack(e + 'H', buf[ifd:ifd + 2])
    tags = {}
    for i in range(n):
        tag, typ, count, raw = struct.unpack(e + 'HHI4s', buf[ifd + 2 + 12 * i: ifd + 14 + 12 * i])
        size = {1: 1, 2: 1, 3: 2, 4: 4}.get(typ)
        if size is None:
            continue
        data = raw[:size * count] if size * count <= 4 else buf[struct.unpack(e + 'I', raw)[0]: struct.unpack(e + 'I', raw)[0] + size * count]
        tags[tag] = data if typ == 2 else struct.unpack(e + _TYPES[typ] * count, data)
    if tags.get(259, (1,))[0] != 1:
        raise NotImplementedError(f'{path}: compressed TIFF (compression {tags[259][0]}) is not supported by this reader')
    w, h, bits = tags[256][0], tags[257][0], tags.get(258, (8,))[0]
    dtype = np.dtype(e + 'u2') if bits == 16 else np.dtype(np.uint8)
    strips = b''.join(buf[o:o + c] for o, c in zip(tags[273], tags[279]))
    image = np.frombuffer(strips, dtype=dtype, count=w * h).reshape(h, w).astype(dtype.newbyteorder('='))
    desc = tags.get(270, b'').rstrip(b'\x00').decode('utf-8', 'replace')
    return image, desc


def write_image(filename: str, image: np.ndarray, scale: bool = True, scale_factor: Optional[Union[float, Tuple[float, float]]] = None,
                dtype='uint16', metadata: Optional[dict] = None, compress: int = 0) -> None:
    """ref: io/image.py:13-62.  `scale`: stretch to the range of `dtype` (no factor: full range of the data; a (lo, hi) tuple:
    clip to that window) and record the factor in the file so that `read_tiff_image` can undo it."""
    if compress:
        raise NotImplementedError('write_image: compressed TIFF output is not implemented')
    if not filename.endswith('.tiff') and not filename.endswith('.tif'):
        raise NotImplementedError('write_image: only .tiff output is implemented (the session cache)')
    metadata = dict(metadata or {})
    image = np.asarray(image)
    if scale:
        max_int = np.iinfo(dtype).max
        image = image.astype(dtype)
        if not scale_factor:
            scale_factor = int(max_int / np.nanmax(image))
            image = image * scale_factor
        elif isinstance(scale_factor, tuple):
            image = image.astype('float32')
            image = (image - scale_factor[0]) / (scale_factor[1] - scale_factor[0])
            image = np.clip(image, 0, 1) * max_int
        metadata['scale_factor'] = str(scale_factor)
    directory = os.path.dirname(os.path.abspath(filename))
    os.makedirs(directory, exist_ok=True)
    out = image.astype(dtype)
    _write_tiff(filename, out, json.dumps({'shape': list(out.shape), **metadata}))


def read_tiff_image(filename: str, dtype='uint16', scale: bool = True, scale_key: str = 'scale_factor') -> np.ndarray:
    """ref: io/image.py:65-103."""
    image, desc = _read_tiff(filename)
    if scale:
        image_desc = json.loads(desc)
        try:
            scale_factor = int(image_desc[scale_key])
        except ValueError:
            scale_factor = ast.literal_eval(image_desc[scale_key])
        if isinstance(scale_factor, (int, float)):
            image = image / scale_factor
        elif isinstance(scale_factor, tuple):
            iinfo = np.iinfo(image.dtype)
            image = image.astype('float32') / iinfo.max
            image = image * (scale_factor[1] - scale_factor[0]) + scale_factor[0]
    return image.astype(dtype)

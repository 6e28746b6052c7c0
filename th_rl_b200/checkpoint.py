"""Packed on-disk state of a batch of runs ("packed tensors for the rest").

The reference can only persist a finished run as one directory of .npy / log.csv files (th_rl/trainer.py:101-110,
th_rl/agents.py:110-116) and has no resume.  A batch of 10^5-10^6 runs is kept as ONE file instead:

    bytes 0..7     magic  b"THRLPACK"
    bytes 8..15    little-endian uint64: length H of the JSON header
    bytes 16..16+H JSON header: format version, ABI version, the game's config (reference JSON schema), n_runs, run_id0, seed,
                   epoch, table dtype, `extra` (caller's dict, e.g. the EWM state of stats.CurveHistogram), and for every
                   array its name, dtype, shape and byte offset
    then           the raw little-endian arrays, each starting on a 4,096-byte boundary: q [R, run_stride] (padded slab layout
                   of include/thrl.h), counter, eps, price, and when the game has them hp, mlp, ring

Arrays are written straight from / read straight into page-locked host buffers (numpy.memmap views on load), so a restore
is one host->device copy per array.  The header carries the config, so a file is self-describing: `load` rebuilds the game.
"""
import json
import os
import struct

import numpy

MAGIC = b"THRLPACK"
FORMAT_VERSION = 1
ALIGN = 4096


def _align(x):
    return (x + ALIGN - 1) // ALIGN * ALIGN


def write_pack(path, header, arrays):
    """arrays: dict name -> C-contiguous numpy array.  Atomic: written to path + '.tmp', then renamed."""
    metas, off = [], 0
    hdr = dict(header, format_version=FORMAT_VERSION, arrays=metas)
    # two passes: offsets depend on the header length, which depends on the offsets' digits -> reserve a fixed-size header block
    for name, a in arrays.items():
        a = numpy.ascontiguousarray(a)
        metas.append(dict(name=name, dtype=a.dtype.str, shape=list(a.shape), offset=0, nbytes=int(a.nbytes)))
    blob = json.dumps(hdr).encode()
    data0 = _align(16 + len(blob) + 64 * len(metas))  # room for the offsets to grow to their final width
    off = data0
    for m in metas:
        m["offset"] = off
        off = _align(off + m["nbytes"])
    blob = json.dumps(hdr).encode()
    assert 16 + len(blob) <= data0
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for m, a in zip(metas, arrays.values()):
            f.seek(m["offset"])
            numpy.ascontiguousarray(a).tofile(f)
        f.truncate(max(off, data0))
    os.replace(tmp, path)
    return off


def read_pack(path, mmap=True):
    """-> (header dict, dict name -> numpy array).  With mmap the arrays are read-only views of the file."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s is not a THRLPACK file" % path)
        (hlen,) = struct.unpack("<Q", f.read(8))
        hdr = json.loads(f.read(hlen).decode())
    if hdr.get("format_version") != FORMAT_VERSION:
        raise ValueError("%s: format version %r (this reader: %d)" % (path, hdr.get("format_version"), FORMAT_VERSION))
    arrays = {}
    for m in hdr["arrays"]:
        shape, dt = tuple(m["shape"]), numpy.dtype(m["dtype"])
        if mmap and m["nbytes"]:
            arrays[m["name"]] = numpy.memmap(path, dtype=dt, mode="r", offset=m["offset"], shape=shape)
        else:
            with open(path, "rb") as f:
                f.seek(m["offset"])
                arrays[m["name"]] = numpy.fromfile(f, dtype=dt, count=int(numpy.prod(shape, dtype=numpy.int64))).reshape(shape)
    return hdr, arrays

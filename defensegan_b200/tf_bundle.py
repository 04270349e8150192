"""TensorFlow checkpoint-V2 ("tensor bundle") reader without TensorFlow.

The reference saves its GAN with `tf.train.Saver` (models/base_model.py:383-395 -> `GAN.model-<step>.index`,
`GAN.model-<step>.data-00000-of-00001` plus the `checkpoint` state file) and restores only the generator
variables from the latest checkpoint (models/gan.py:80-87, base_model.py:294-335).  This module reads that
format directly so a trained reference checkpoint can be used here without a TF1 installation:

* `<prefix>.index` is a LevelDB-style sorted string table: data blocks of prefix-compressed (key, value)
  entries with a restart array, per-block trailer (compression byte + masked CRC-32C), an index block, and a
  48-byte footer ending in the magic 0xdb4775248b80fb57.  Key "" holds a BundleHeaderProto, every other key
  is a variable name whose value is a BundleEntryProto {1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size,
  6: crc32c (masked, fixed32), 7: slices}.
* `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset + size).

`write_bundle` emits the same format (one shard, uncompressed blocks, prefix compression with restart
interval 16) and is what the tests and `save_generator(..., fmt="tf")` use.

PARITY NOTE: TensorFlow is not installable in this environment and the reference ships no checkpoint, so the
reader is validated against this writer, against TensorBoard's independent CRC-32C / TensorShapeProto
implementations (tests/test_host.py), and by hand against the format as TensorFlow documents it - not yet
against a file written by TensorFlow itself.
"""
from __future__ import annotations

import collections
import os
import re
import struct
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_FOOTER_LEN = 48
_BLOCK_TRAILER = 5

# tensorflow/core/framework/types.proto
_DTYPE_OF = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
             6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
             22: np.dtype("<u4"), 23: np.dtype("<u8")}
_ENUM_OF = {v: k for k, v in _DTYPE_OF.items()}


# ---------------------------------------------------------------------------------------------------
# CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) and TensorFlow's masking
# ---------------------------------------------------------------------------------------------------
def _make_table():
    tab = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    tab = _CRC_TABLE
    c = crc ^ 0xFFFFFFFF
    for b in memoryview(data).cast("B"):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------
# varints / minimal protobuf wire format
# ---------------------------------------------------------------------------------------------------
def _get_varint(buf, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _put_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_fields(buf) -> Iterable[Tuple[int, int, object]]:
    """Yield (field number, wire type, value) of one protobuf message (varint / fixed64 / bytes / fixed32)."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + n])
            if len(val) != n:
                raise ValueError("truncated length-delimited field")
            pos += n
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


BundleEntry = collections.namedtuple("BundleEntry", "dtype shape shard_id offset size crc32c has_slices")


def parse_tensor_shape(buf: bytes) -> Tuple[int, ...]:
    """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1; string name = 2}; bool unknown_rank = 3."""
    dims = []
    for field, _, val in _parse_fields(buf):
        if field == 2:
            size = 0
            for f2, _, v2 in _parse_fields(val):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif field == 3 and val:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def encode_tensor_shape(shape: Iterable[int]) -> bytes:
    out = b""
    for d in shape:
        dim = b"\x08" + _put_varint(int(d)) if int(d) != 0 else b""   # proto3 omits zero-valued scalars
        out += b"\x12" + _put_varint(len(dim)) + dim
    return out


def parse_bundle_entry(buf: bytes) -> BundleEntry:
    dtype = shard = offset = size = crc = 0
    shape: Tuple[int, ...] = ()
    slices = False
    for field, _, val in _parse_fields(buf):
        if field == 1:
            dtype = val
        elif field == 2:
            shape = parse_tensor_shape(val)
        elif field == 3:
            shard = val
        elif field == 4:
            offset = val
        elif field == 5:
            size = val
        elif field == 6:
            crc = val
        elif field == 7:
            slices = True
    return BundleEntry(dtype, shape, shard, offset, size, crc, slices)


def encode_bundle_entry(dtype_enum: int, shape, shard_id: int, offset: int, size: int, masked_crc: int) -> bytes:
    sh = encode_tensor_shape(shape)
    out = b"\x08" + _put_varint(dtype_enum) + b"\x12" + _put_varint(len(sh)) + sh
    if shard_id:
        out += b"\x18" + _put_varint(shard_id)
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", masked_crc)
    return out


# ---------------------------------------------------------------------------------------------------
# sorted string table
# ---------------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    end = offset + size
    if end + _BLOCK_TRAILER > len(data):
        raise ValueError("table block [%d, %d) runs past the end of the index file" % (offset, end))
    ctype = data[end]
    if verify:
        stored = struct.unpack_from("<I", data, end + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:end + 1]):
            raise ValueError("checksum mismatch in table block at offset %d" % offset)
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d); TensorFlow writes bundle indices uncompressed" % ctype)
    return data[offset:end]


def _block_entries(block: bytes):
    if len(block) < 4:
        raise ValueError("table block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise ValueError("bad restart array in table block")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > limit:
            raise ValueError("corrupt table entry")
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> "collections.OrderedDict[bytes, bytes]":
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < _FOOTER_LEN:
        raise ValueError("%s is too short to be a TensorFlow bundle index" % path)
    footer = data[-_FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, _FOOTER_LEN - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s: bad table magic (not a checkpoint-V2 .index file)" % path)
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out: "collections.OrderedDict[bytes, bytes]" = collections.OrderedDict()
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        b_off, p = _get_varint(handle, 0)
        b_size, _ = _get_varint(handle, p)
        for k, v in _block_entries(_read_block(data, b_off, b_size, verify)):
            out[k] = v
    return out


class _BlockBuilder:
    def __init__(self, restart_interval: int = 16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.interval = restart_interval

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count < self.interval:
            n = min(len(self.last), len(key))
            while shared < n and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        out = bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts)
        return out + struct.pack("<I", len(self.restarts))

    def empty(self) -> bool:
        return not self.buf


def write_table(path: str, items: "Iterable[Tuple[bytes, bytes]]", block_size: int = 262144) -> None:
    """Sorted (key, value) pairs -> table file (uncompressed blocks, as TensorFlow's BundleWriter asks for)."""
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)

    def emit(block: bytes) -> bytes:
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return handle

    cur = _BlockBuilder()
    last_key = None
    for key, value in items:
        if last_key is not None and key <= last_key:
            raise ValueError("table keys must be strictly increasing")
        if not cur.empty() and len(cur.buf) >= block_size:
            index.add(last_key, emit(cur.finish()))
            cur = _BlockBuilder()
        cur.add(key, value)
        last_key = key
    if not cur.empty():
        index.add(last_key, emit(cur.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (_FOOTER_LEN - 8 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---------------------------------------------------------------------------------------------------
# bundle level
# ---------------------------------------------------------------------------------------------------
def _data_path(prefix: str, shard: int, num_shards: int) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def _read_index(prefix: str, verify: bool):
    table = read_table(prefix + ".index", verify)
    if b"" not in table:
        raise ValueError("%s.index has no bundle header" % prefix)
    num_shards, endianness = 1, 0
    for field, _, val in _parse_fields(table[b""]):
        if field == 1:
            num_shards = val
        elif field == 2:
            endianness = val
    if endianness != 0:
        raise NotImplementedError("big-endian tensor bundle")
    entries = collections.OrderedDict((k.decode("utf-8"), parse_bundle_entry(v)) for k, v in table.items() if k != b"")
    return num_shards, entries


def list_bundle(prefix: str) -> "collections.OrderedDict[str, Tuple[np.dtype, Tuple[int, ...]]]":
    """{variable name: (dtype, shape)} of a checkpoint prefix (e.g. '.../GAN.model-20000')."""
    _, entries = _read_index(prefix, True)
    return collections.OrderedDict((k, (_DTYPE_OF.get(e.dtype), e.shape)) for k, e in entries.items())


def read_bundle(prefix: str, names: Optional[Iterable[str]] = None, verify_crc: bool = True
                ) -> "collections.OrderedDict[str, np.ndarray]":
    """Read variables of a checkpoint prefix into numpy arrays (all of them, or `names`)."""
    num_shards, entries = _read_index(prefix, verify_crc)
    wanted = list(entries) if names is None else list(names)
    out: "collections.OrderedDict[str, np.ndarray]" = collections.OrderedDict()
    files: Dict[int, object] = {}
    try:
        for name in wanted:
            if name not in entries:
                raise KeyError("variable %r is not in checkpoint %s" % (name, prefix))
            e = entries[name]
            if e.has_slices:
                raise NotImplementedError("partitioned variable %r (tensor slices) is not supported" % name)
            if e.dtype not in _DTYPE_OF:
                raise NotImplementedError("variable %r has unsupported dtype enum %d" % (name, e.dtype))
            dt = _DTYPE_OF[e.dtype]
            count = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
            if count * dt.itemsize != e.size:
                raise ValueError("variable %r: %d bytes on disk, shape %s needs %d" % (name, e.size, e.shape, count * dt.itemsize))
            if e.shard_id not in files:
                files[e.shard_id] = open(_data_path(prefix, e.shard_id, num_shards), "rb")
            f = files[e.shard_id]
            f.seek(e.offset)
            raw = f.read(e.size)
            if len(raw) != e.size:
                raise ValueError("variable %r: data file truncated" % name)
            if verify_crc and unmask_crc(e.crc32c) != crc32c(raw):
                raise ValueError("variable %r: checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e.shape).copy()
    finally:
        for f in files.values():
            f.close()
    return out


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 262144) -> None:
    """Write {name: array} as a single-shard checkpoint-V2 bundle (`prefix.index`, `prefix.data-00000-of-00001`)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = []
    offset = 0
    with open(_data_path(prefix, 0, 1), "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            arr = np.asarray(tensors[name], order="C")   # (ascontiguousarray would turn scalars into shape (1,))
            dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
            if np.dtype(dt) not in _ENUM_OF:
                raise NotImplementedError("dtype %s of %r" % (arr.dtype, name))
            raw = arr.astype(dt, copy=False).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"),
                          encode_bundle_entry(_ENUM_OF[np.dtype(dt)], arr.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    # BundleHeaderProto {num_shards = 1, endianness = LITTLE (default), version {producer = 1}}
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"
    write_table(prefix + ".index", [(b"", header)] + items, block_size=block_size)


def latest_checkpoint(ckpt_dir: str) -> Optional[str]:
    """Prefix named by the `checkpoint` state file (`model_checkpoint_path: "GAN.model-20000"`), like
    tf.train.latest_checkpoint; falls back to the highest-numbered `*.index` in the directory."""
    state = os.path.join(ckpt_dir, "checkpoint")
    if os.path.isfile(state):
        with open(state, "r") as f:
            m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', f.read(), re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(ckpt_dir, p)
            if os.path.isfile(p + ".index"):
                return p
    best, best_step = None, -1
    if os.path.isdir(ckpt_dir):
        for fn in os.listdir(ckpt_dir):
            if fn.endswith(".index"):
                m = re.search(r"-(\d+)\.index$", fn)
                step = int(m.group(1)) if m else 0
                if step > best_step:
                    best, best_step = os.path.join(ckpt_dir, fn[:-len(".index")]), step
    return best


def read_generator_variables(prefix: str, verify_crc: bool = True) -> "collections.OrderedDict[str, np.ndarray]":
    """The reference's generator restore (`slim.get_variables('Generator')`, models/gan.py:80-87): every variable
    whose name starts with 'Generator', without optimizer slots (`.../Adam`, `.../Adam_1`)."""
    names = [n for n in list_bundle(prefix) if n.startswith("Generator") and not re.search(r"/(Adam(_\d+)?|RMSProp(_\d+)?|Momentum)$", n)]
    return read_bundle(prefix, names, verify_crc)

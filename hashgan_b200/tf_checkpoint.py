"""Reader (and a minimal writer) for TensorFlow-1.x V2 checkpoints ("tensor bundles").

Reference call sites: the discriminator weights of an evaluation run come from
``tf.train.Saver(params_with_name('discriminator')).restore(session, cfg.MODEL.D_PRETRAINED_MODEL_PATH)``
(main.py:187-195, lib/params.py:38-39; ``config/cifar_evaluation.yaml:5`` names ``D_20000.ckpt``), written by
``saver.save`` every CHECKPOINT_FREQUENCY iterations (main.py:243-250).  TensorFlow 1.12 is not installable here, so the
format is restated from its published layout (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*):
**parity unpinned** -- there is no TensorFlow-written fixture in the reference; every checksum the format carries
(CRC32C of each table block and of each tensor) is verified on read instead.

    <prefix>.index                 an immutable sorted string table (LevelDB "table" format):
                                   data blocks | meta-index block | index block | 48-byte footer, magic 0xdb4775248b80fb57
        block  = entries + restart array (uint32 offsets, uint32 count) + 1 byte compression + 4 bytes masked CRC32C
        entry  = varint shared, varint non_shared, varint value_len, key suffix, value      (prefix-compressed keys)
        key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
        key variable name -> BundleEntryProto {1: dtype, 2: shape{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size,
                                               6: crc32c (fixed32, masked), 7: slices}
    <prefix>.data-SSSSS-of-NNNNN   the raw little-endian tensor bytes, row-major, at [offset, offset + size)

Only what the evaluation path needs is implemented: dense tensors (no slices), uncompressed blocks (TensorFlow writes
bundle tables uncompressed), the numeric dtypes below.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

__all__ = ["read_checkpoint", "list_variables", "write_checkpoint", "crc32c", "CheckpointError"]

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(ValueError):
    pass


# ---- CRC32C (Castagnoli), masked the LevelDB way ---------------------------------------------------------------
def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_TABLE = _make_table()
_NP_TABLE = np.array(_TABLE, dtype=np.uint32)


def _native_crc(data, crc: int) -> Optional[int]:
    """hg_crc32c from libhashgan_b200 (plain host C) when the library is built; None otherwise."""
    try:
        import ctypes as C

        from . import _native

        lib = _native.lib()
        fn = lib.hg_crc32c
    except Exception:
        return None
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    buf = np.ascontiguousarray(buf)
    return int(fn(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.c_uint32(crc))) & 0xFFFFFFFF


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C of `data` (bytes or uint8 array), continuing from `crc`.  crc32c(b"123456789") == 0xE3069283."""
    n = len(data) if not isinstance(data, np.ndarray) else data.size
    if n > 4096:
        got = _native_crc(data, crc)
        if got is not None:
            return got
    c = crc ^ 0xFFFFFFFF
    tab = _TABLE
    for byte in (bytes(data) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8).tobytes()):
        c = tab[(c ^ byte) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def _unmask(masked: int) -> int:
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format --------------------------------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_message(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field, wire_type, value)]: varint -> int, fixed32/64 -> int, length-delimited -> bytes."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            if len(v) != n:
                raise CheckpointError("truncated length-delimited field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError(f"unsupported protobuf wire type {wt}")
        out.append((field, wt, v))
    return out


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


# ---- the table -------------------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    end = offset + size
    if end + 5 > len(data):
        raise CheckpointError("table block handle points outside the file")
    block, ctype = data[offset:end], data[end]
    stored = struct.unpack_from("<I", data, end + 1)[0]
    if verify and _unmask(stored) != crc32c(data[offset:end + 1]):
        raise CheckpointError("table block checksum mismatch (corrupt .index file)")
    if ctype != 0:
        raise CheckpointError("compressed table blocks are not supported (TensorFlow writes bundle indexes uncompressed)")
    return block


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise CheckpointError("table block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt table entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_table(path: str, verify: bool) -> List[Tuple[bytes, bytes]]:
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) < 48:
        raise CheckpointError(f"{path}: too short for a table footer")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointError(f"{path}: not a TensorFlow V2 checkpoint index (bad table magic)")
    pos = 0
    _, pos = _get_varint(footer, pos)  # meta-index handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ---- bundle entries --------------------------------------------------------------------------------------------------
class _Entry:
    __slots__ = ("dtype", "shape", "shard", "offset", "size", "crc", "sliced")

    def __init__(self, value: bytes):
        self.dtype, self.shape, self.shard, self.offset, self.size, self.crc, self.sliced = 0, (), 0, 0, 0, None, False
        for field, wt, v in _parse_message(value):
            if field == 1:
                self.dtype = int(v)
            elif field == 2:
                dims = []
                for f2, _, v2 in _parse_message(v):
                    if f2 == 2:
                        size = 0
                        for f3, _, v3 in _parse_message(v2):
                            if f3 == 1:
                                size = _signed64(int(v3))
                        dims.append(size)
                    elif f2 == 3 and v2:
                        raise CheckpointError("tensor of unknown rank in checkpoint")
                self.shape = tuple(dims)
            elif field == 3:
                self.shard = int(v)
            elif field == 4:
                self.offset = int(v)
            elif field == 5:
                self.size = int(v)
            elif field == 6:
                self.crc = int(v)
            elif field == 7:
                self.sliced = True


def _index(prefix: str, verify: bool):
    path = prefix + ".index"
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} (a V2 checkpoint is <prefix>.index + <prefix>.data-00000-of-0000N)")
    num_shards, entries = 1, {}
    for key, value in _read_table(path, verify):
        if key == b"":
            for field, _, v in _parse_message(value):
                if field == 1:
                    num_shards = int(v)
                elif field == 2 and int(v) != 0:
                    raise CheckpointError("big-endian checkpoints are not supported")
        else:
            entries[key.decode("utf-8")] = _Entry(value)
    return num_shards, entries


def list_variables(prefix: str) -> Dict[str, Tuple[np.dtype, Tuple[int, ...]]]:
    """{variable name: (dtype, shape)} of the checkpoint `<prefix>.index`."""
    _, entries = _index(prefix, True)
    return {k: (np.dtype(_DTYPES.get(e.dtype, np.void)), e.shape) for k, e in entries.items()}


def read_checkpoint(prefix: str, names: Optional[Iterable[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """Tensors of a TensorFlow V2 checkpoint by variable name (all of them, or `names`).

    `verify` checks the CRC32C of every table block and of every tensor read."""
    num_shards, entries = _index(prefix, verify)
    want = list(entries) if names is None else list(names)
    out, files = {}, {}
    try:
        for name in want:
            if name not in entries:
                raise KeyError(f"{name} is not in {prefix}.index")
            e = entries[name]
            if e.sliced:
                raise CheckpointError(f"{name}: partitioned (sliced) variables are not supported")
            if e.dtype not in _DTYPES:
                raise CheckpointError(f"{name}: unsupported dtype enum {e.dtype}")
            dt = np.dtype(_DTYPES[e.dtype])
            count = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
            if count * dt.itemsize != e.size:
                raise CheckpointError(f"{name}: {e.size} bytes do not match shape {e.shape} of {dt}")
            if e.shard not in files:
                files[e.shard] = open(f"{prefix}.data-{e.shard:05d}-of-{num_shards:05d}", "rb")
            fh = files[e.shard]
            fh.seek(e.offset)
            raw = fh.read(e.size)
            if len(raw) != e.size:
                raise CheckpointError(f"{name}: data file is truncated")
            if verify and e.crc is not None and _unmask(e.crc) != crc32c(raw):
                raise CheckpointError(f"{name}: tensor checksum mismatch (corrupt data file)")
            out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt, copy=False).reshape(e.shape)
    finally:
        for fh in files.values():
            fh.close()
    return out


# ---- minimal writer (fixtures, export of converted weights) ------------------------------------------------------------
def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _field(num: int, wt: int, payload: bytes) -> bytes:
    return _put_varint((num << 3) | wt) + payload


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], entries_per_block: int = 8) -> None:
    """Write `tensors` as a single-shard V2 checkpoint in the layout described in the module docstring."""
    data, items = bytearray(), []
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda m: _put_varint(len(m)) + m)(_field(1, 0, _put_varint(1))))
    items.append((b"", header))
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        a = np.asarray(tensors[name])
        if a.ndim and not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        dt = a.dtype.newbyteorder("=")
        if np.dtype(dt) not in _DTYPE_CODES:
            raise CheckpointError(f"{name}: dtype {a.dtype} cannot be written")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        dims = b"".join(_field(2, 2, (lambda m: _put_varint(len(m)) + m)(_field(1, 0, _put_varint(int(d))))) for d in a.shape)
        msg = _field(1, 0, _put_varint(_DTYPE_CODES[np.dtype(dt)])) + _field(2, 2, _put_varint(len(dims)) + dims)
        if len(data):
            msg += _field(4, 0, _put_varint(len(data)))
        msg += _field(5, 0, _put_varint(len(raw))) + _field(6, 5, struct.pack("<I", _mask(crc32c(raw))))
        items.append((name.encode("utf-8"), msg))
        data += raw
    table, index_items = bytearray(), []

    def emit(block: bytes) -> bytes:
        off = len(table)
        table.extend(block)
        table.append(0)  # no compression
        table.extend(struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        index_items.append((chunk[-1][0], emit(_build_block(chunk))))  # separator key: >= every key of the block
    meta = emit(_build_block([]))
    index = emit(_build_block(index_items, restart_interval=1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    table.extend(footer)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(table))
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))

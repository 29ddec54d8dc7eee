"""Reader (and, for tests, writer) of TensorFlow "tensor bundle" checkpoints -- the V2 format `tf.train.Saver` writes
and `saver.restore(session, 'weights/demon_original')` reads in the reference (examples/example.py:82-83,
weights/download_weights.sh:2).  Pure Python + numpy, no TensorFlow.

A checkpoint `<prefix>` is two kinds of file:

  <prefix>.index                 an SSTable (the LevelDB table format of tensorflow/core/lib/io/table*) that maps
                                 ""            -> BundleHeaderProto  {num_shards, endianness, version}
                                 tensor name   -> BundleEntryProto   {dtype, shape, shard_id, offset, size, crc32c}
  <prefix>.data-SSSSS-of-NNNNN   the raw little-endian tensor bytes of shard SSSSS at [offset, offset + size)

SSTable (LevelDB `table_format.md`): data blocks of prefix-compressed entries
(varint shared, varint non_shared, varint value_len, key suffix, value) followed by a restart array, every block
trailed by a 1-byte compression type (0 = none, 1 = snappy) and a masked CRC32C; a metaindex block, an index block
whose values are BlockHandles (varint offset, varint size) of the data blocks, and a 48-byte footer
(metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).

The variable names are the ones TensorFlow creates for the reference graph (`netFlow1/conv1y/kernel`, ...,
demon_b200/weights.py), so `load_checkpoint(prefix)` feeds `Session.load_weights` / `demon_net_set_weight` directly;
optimizer slots and `global_step` that a training checkpoint also holds are ignored by name.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DTYPE_CODES = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}


class CheckpointError(ValueError):
    pass


# ---- CRC32C (Castagnoli), masked the LevelDB way -----------------------------------------------------------------------
def _make_crc_table():
    poly = 0x82F63B78
    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _make_crc_table()


def crc32c(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format ------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    result = shift = 0
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


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_message(buf):
    """protobuf wire format -> list of (field number, wire type, value); value is int or bytes."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise CheckpointError("truncated length-delimited field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _field(field, wt, payload):
    return _put_varint((field << 3) | wt) + payload


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_bundle_entry(buf):
    """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto): dtype=1, shape=2, shard_id=3, offset=4, size=5,
    crc32c=6 (fixed32), slices=7."""
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, _, v in _parse_message(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:   # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1; string name = 2}; unknown_rank = 3
            dims = []
            for f2, _, v2 in _parse_message(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _parse_message(v2):
                        if f3 == 1:
                            size = _signed(v3)
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["slices"] += 1
    return e


def parse_bundle_header(buf):
    """BundleHeaderProto: num_shards=1, endianness=2 (0 little, 1 big), version=3."""
    h = {"num_shards": 1, "endianness": 0}
    for field, _, v in _parse_message(buf):
        if field == 1:
            h["num_shards"] = v
        elif field == 2:
            h["endianness"] = v
    return h


# ---- snappy (the index blocks may be compressed: table::Options defaults to kSnappyCompression) ----------------------
def snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy stream (bad copy offset)")
        for _ in range(ln):   # byte-wise: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("corrupt snappy stream (length %d, header says %d)" % (len(out), n))
    return bytes(out)


def snappy_compress_literal(data):
    """A valid snappy stream made of literals only (what the test writer uses to exercise the compressed-block path)."""
    out = bytearray(_put_varint(len(data)))
    pos = 0
    while pos < len(data):
        chunk = data[pos:pos + 65536]
        ln = len(chunk) - 1
        if ln < 60:
            out.append(ln << 2)
        else:
            nb = (ln.bit_length() + 7) // 8
            out.append((59 + nb) << 2)
            out += ln.to_bytes(nb, "little")
        out += chunk
        pos += len(chunk)
    return bytes(out)


# ---- SSTable ------------------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify=True):
    end = offset + size
    if end + BLOCK_TRAILER_LEN > len(buf):
        raise CheckpointError("block handle [%d, +%d) runs past the end of the index file" % (offset, size))
    raw = buf[offset:end]
    ctype = buf[end]
    if verify:
        stored = unmask_crc(struct.unpack_from("<I", buf, end + 1)[0])
        if crc32c(buf[offset:end + 1]) != stored:
            raise CheckpointError("CRC mismatch in an index block at offset %d" % offset)
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_uncompress(raw)
    raise CheckpointError("unknown block compression type %d" % ctype)


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < FOOTER_LEN:
        raise CheckpointError("%s is too small to be a table" % path)
    footer = buf[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, FOOTER_LEN - 8)[0] != TABLE_MAGIC:
        raise CheckpointError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    _, pos = _get_varint(footer, 0)       # metaindex offset
    _, pos = _get_varint(footer, pos)     # metaindex size
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p = _get_varint(handle, 0)
        bsize, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return out


# ---- bundle reader ------------------------------------------------------------------------------------------------------
def _shard_path(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def list_variables(prefix, verify=True):
    """[(name, shape, dtype code)] of a checkpoint, like tf.train.list_variables."""
    out = []
    for key, value in read_table(prefix + ".index", verify):
        if key == b"":
            continue
        e = parse_bundle_entry(value)
        out.append((key.decode("utf-8"), e["shape"], e["dtype"]))
    return out


def load_checkpoint(prefix, names=None, verify_data_crc=False, verify_index_crc=True):
    """Reads the tensors of checkpoint `prefix` (e.g. 'weights/demon_original') into a dict name -> numpy array.
    names: restrict to these variable names (missing ones raise KeyError).  verify_data_crc: check the CRC32C of every
    tensor's bytes as well (pure Python, ~1 s per MB)."""
    index = prefix + ".index"
    if not os.path.isfile(index):
        raise FileNotFoundError("no checkpoint index at %s (expected the files %s.index and %s.data-00000-of-0000N)" % (index, prefix, prefix))
    entries, header = {}, {"num_shards": 1, "endianness": 0}
    for key, value in read_table(index, verify_index_crc):
        if key == b"":
            header = parse_bundle_header(value)
        else:
            entries[key.decode("utf-8")] = parse_bundle_entry(value)
    if header["endianness"] != 0:
        raise CheckpointError("big-endian checkpoints are not supported")
    wanted = list(entries) if names is None else list(names)
    missing = [n for n in wanted if n not in entries]
    if missing:
        raise KeyError("checkpoint %s has no variable(s) %s" % (prefix, ", ".join(missing[:5]) + (" ..." if len(missing) > 5 else "")))
    shards = {}
    out = {}
    for name in wanted:
        e = entries[name]
        if e["slices"]:
            raise CheckpointError("variable %s is stored as slices (partitioned variable): not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise CheckpointError("variable %s has unsupported dtype code %d" % (name, e["dtype"]))
        dt = _DTYPES[e["dtype"]]
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise CheckpointError("variable %s: shape %s does not match %d bytes" % (name, e["shape"], e["size"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap(_shard_path(prefix, sid, header["num_shards"]), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise CheckpointError("variable %s runs past the end of its data shard" % name)
        if verify_data_crc and e["crc32c"] is not None and crc32c(raw.tobytes()) != unmask_crc(e["crc32c"]):
            raise CheckpointError("CRC mismatch in the data of variable %s" % name)
        out[name] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e["shape"]).astype(dt.newbyteorder("="), copy=True)
    return out


def load_demon_weights(prefix, verify_data_crc=False):
    """The 242 variables of the `networks_original` graphs out of checkpoint `prefix`, ready for Session.load_weights.
    Raises KeyError naming what is missing, ValueError on a shape that differs from the graph's."""
    from . import weights as W
    specs = W.variable_specs()
    got = load_checkpoint(prefix, names=list(specs), verify_data_crc=verify_data_crc)
    for name, (_, shape) in specs.items():
        if tuple(got[name].shape) != tuple(shape):
            raise ValueError("checkpoint variable %s has shape %s, the graph needs %s" % (name, got[name].shape, shape))
        got[name] = np.ascontiguousarray(got[name], dtype=np.float32)
    return got


# ---- writer (tests; also lets a user re-save converted weights) -------------------------------------------------------
class _TableBuilder:
    def __init__(self, block_size=4096, restart_interval=16, compress=False):
        self.out = bytearray()
        self.block_size, self.restart_interval, self.compress = block_size, restart_interval, compress
        self.index_entries = []
        self._reset_block()
        self.last_key = b""

    def _reset_block(self):
        self.block = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.block_last_key = b""

    def add(self, key, value):
        assert key >= self.last_key, "keys must be added in sorted order"
        shared = 0
        if self.counter < self.restart_interval:
            m = min(len(key), len(self.block_last_key))
            while shared < m and key[shared] == self.block_last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.block))
            self.counter = 0
        self.block += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.block_last_key = key
        self.last_key = key
        self.counter += 1
        if len(self.block) >= self.block_size:
            self._flush()

    def _emit(self, contents):
        raw = bytes(contents)
        ctype = 0
        if self.compress:
            raw, ctype = snappy_compress_literal(raw), 1
        off = len(self.out)
        self.out += raw
        self.out.append(ctype)
        self.out += struct.pack("<I", mask_crc(crc32c(raw + bytes([ctype]))))
        return off, len(raw)

    def _finish_block(self, block, restarts):
        return bytes(block) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))

    def _flush(self):
        if not self.block:
            return
        off, size = self._emit(self._finish_block(self.block, self.restarts))
        self.index_entries.append((self.block_last_key, _put_varint(off) + _put_varint(size)))
        self._reset_block()

    def finish(self):
        self._flush()
        moff, msize = self._emit(self._finish_block(b"", [0]))              # empty metaindex block
        blk, restarts, = bytearray(), []
        for key, handle in self.index_entries:                              # index block: restart at every entry
            restarts.append(len(blk))
            blk += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(handle)) + key + handle
        ioff, isize = self._emit(self._finish_block(blk, restarts or [0]))
        footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
        footer += b"\0" * (FOOTER_LEN - 8 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
        self.out += footer
        return bytes(self.out)


def save_checkpoint(prefix, tensors, compress_index=False, block_size=4096, crc_limit=1 << 16):
    """Writes dict name -> numpy array as a one-shard tensor bundle `<prefix>.index` + `<prefix>.data-00000-of-00001`.
    Tensors larger than crc_limit bytes get no data CRC (the reader skips the check for entries without one)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    tb = _TableBuilder(block_size=block_size, compress=compress_index)
    version = _field(1, 0, _put_varint(1))                                              # VersionDef.producer = 1
    tb.add(b"", _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _field(3, 2, _put_varint(len(version)) + version))
    offset = 0
    with open(_shard_path(prefix, 0, 1), "wb") as data:
        for name in sorted(tensors, key=lambda n: n.encode("utf-8")):
            a = np.asarray(tensors[name])   # (ascontiguousarray would turn a scalar into shape (1,); tobytes() is C order anyway)
            if a.dtype not in _DTYPE_CODES:
                raise TypeError("unsupported dtype %s for %s" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            shape = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(s)) for s in a.shape))
            entry = _field(1, 0, _put_varint(_DTYPE_CODES[a.dtype])) + _field(2, 2, _put_varint(len(shape)) + shape)
            entry += _field(3, 0, _put_varint(0)) + _field(4, 0, _put_varint(offset)) + _field(5, 0, _put_varint(len(raw)))
            if len(raw) <= crc_limit:   # (TensorFlow always writes it; the pure-Python CRC costs ~1 s per MB)
                entry += _field(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
            tb.add(name.encode("utf-8"), entry)
            data.write(raw)
            offset += len(raw)
    with open(prefix + ".index", "wb") as f:
        f.write(tb.finish())
    return prefix

"""CPU: the pure-Python TensorFlow tensor-bundle reader (demon_b200/checkpoint.py) -- the `saver.restore(session,
'weights/demon_original')` of examples/example.py:82-83.  The pretrained checkpoint is not in the reference repository
(weights/download_weights.sh:2 is a wget), so the format is exercised on bundles written here in the documented
layout: SSTable index (prefix-compressed blocks, restart arrays, masked CRC32C trailers, footer with the table magic),
BundleHeaderProto / BundleEntryProto values, raw little-endian data shard."""
import os
import struct

import numpy as np
import pytest

from demon_b200 import checkpoint as ck
from demon_b200 import weights as W


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(b"123456789") == 0xE3069283
    for v in (0, 1, 0xDEADBEEF, 0xFFFFFFFF):
        assert ck.unmask_crc(ck.mask_crc(v)) == v


def test_snappy_decoder_literals_and_copies():
    data = bytes(range(256)) * 300
    assert ck.snappy_uncompress(ck.snappy_compress_literal(data)) == data
    # hand-made stream: literal "abcd", copy-1 (offset 4, len 8: overlapping run), copy-2 (offset 12, len 5)
    stream = ck._put_varint(17) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1 | (0 << 5), 4]) + bytes([((5 - 1) << 2) | 2, 12, 0])
    assert ck.snappy_uncompress(stream) == b"abcd" + b"abcdabcd" + b"abcda"
    with pytest.raises(ck.CheckpointError):
        ck.snappy_uncompress(ck._put_varint(4) + bytes([((4 - 1) << 2) | 2, 9, 0]))   # copy before any output


@pytest.mark.parametrize("compress", (False, True))
def test_bundle_round_trip_small(tmp_path, compress):
    rng = np.random.RandomState(0)
    tensors = {"scope/a/kernel": rng.rand(3, 1, 6, 32).astype(np.float32), "scope/a/bias": rng.rand(32).astype(np.float32),
               "global_step": np.array(1234, np.int64), "z/double": rng.rand(2, 3), "empty/x": np.zeros((0, 4), np.float32)}
    for i in range(300):     # many keys: several data blocks, restart points, shared prefixes
        tensors["netFlow1/layer_%03d/kernel" % i] = rng.rand(2, 2).astype(np.float32)
    prefix = ck.save_checkpoint(str(tmp_path / "model"), tensors, compress_index=compress, block_size=512)
    assert os.path.isfile(prefix + ".index") and os.path.isfile(prefix + ".data-00000-of-00001")
    got = ck.load_checkpoint(prefix, verify_data_crc=True)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    names = [n for n, _, _ in ck.list_variables(prefix)]
    assert names == sorted(tensors, key=lambda n: n.encode())
    with pytest.raises(KeyError):
        ck.load_checkpoint(prefix, names=["not/there"])


def test_bundle_detects_corruption(tmp_path):
    prefix = ck.save_checkpoint(str(tmp_path / "m"), {"a": np.arange(6, dtype=np.float32).reshape(2, 3)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    bad = bytearray(raw); bad[5] ^= 0x40
    open(prefix + ".index", "wb").write(bad)
    with pytest.raises(ck.CheckpointError):
        ck.load_checkpoint(prefix)
    bad = bytearray(raw); bad[-1] ^= 0xFF
    open(prefix + ".index", "wb").write(bad)
    with pytest.raises(ck.CheckpointError):
        ck.load_checkpoint(prefix)
    open(prefix + ".index", "wb").write(raw)
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read()); d[3] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(d)
    with pytest.raises(ck.CheckpointError):
        ck.load_checkpoint(prefix, verify_data_crc=True)
    with pytest.raises(FileNotFoundError):
        ck.load_checkpoint(str(tmp_path / "nothing"))


def test_all_242_graph_variables_resolve_through_session_restore(tmp_path):
    """A checkpoint with the reference's variable names and layouts (plus what a training checkpoint also holds)
    restores into the Session like examples/example.py:82-83 does; a wrong shape or a missing variable is an error."""
    from demon_b200.networks_original import Session
    w = W.synthetic_weights(3)
    assert len(w) == 242 == len(W.variable_specs())
    extra = dict(w)
    extra["global_step"] = np.array(7, np.int64)
    extra["netFlow1/conv1y/kernel/Adam"] = np.zeros((9, 1, 6, 32), np.float32)
    extra["beta1_power"] = np.array(0.9, np.float32)
    prefix = ck.save_checkpoint(str(tmp_path / "demon_original"), extra)
    sess = Session()
    sess.restore(prefix)
    assert set(sess.weights) == set(w)
    for k in w:
        assert sess.weights[k].dtype == np.float32 and np.array_equal(sess.weights[k], w[k]), k
    broken = dict(w); del broken["netRefine/conv0/bias"]
    p2 = ck.save_checkpoint(str(tmp_path / "broken"), broken)
    with pytest.raises(KeyError, match="netRefine/conv0/bias"):
        Session().restore(p2)
    broken = dict(w); broken["netDM1/motion_fc1/kernel"] = np.zeros((1024, 6144), np.float32)
    p3 = ck.save_checkpoint(str(tmp_path / "transposed"), broken)
    with pytest.raises(ValueError, match="motion_fc1"):
        Session().restore(p3)


def test_entry_proto_layout_is_the_documented_one(tmp_path):
    """Byte-level check of one index entry against tensor_bundle.proto / tensor_shape.proto field numbers."""
    prefix = ck.save_checkpoint(str(tmp_path / "m"), {"v": np.zeros((9, 1, 6, 32), np.float32)})
    kv = dict(ck.read_table(prefix + ".index"))
    hdr = ck.parse_bundle_header(kv[b""])
    assert hdr == {"num_shards": 1, "endianness": 0}
    e = kv[b"v"]
    # dtype: field 1 varint DT_FLOAT(1) -> 08 01 ; shape: field 2 length-delimited, four dims {size}
    assert e[:2] == b"\x08\x01" and e[2] == 0x12
    parsed = ck.parse_bundle_entry(e)
    assert parsed["shape"] == (9, 1, 6, 32) and parsed["size"] == 9 * 6 * 32 * 4 and parsed["offset"] == 0 and parsed["shard_id"] == 0
    footer = open(prefix + ".index", "rb").read()[-8:]
    assert struct.unpack("<Q", footer)[0] == 0xDB4775248B80FB57

"""Read tf.train.Checkpoint files WITHOUT TensorFlow (checkpoint interop; the reference saves
tf.train.Checkpoint(model=model) through a CheckpointManager, mogpe/training/utils.py:228-242).

A checkpoint is a tensor bundle: `<prefix>.index` is a LevelDB-format sorted table (data blocks with prefix-compressed
keys, an index block, a 48-byte footer) whose values are BundleEntryProto messages (1: dtype, 2: shape, 3: shard_id,
4: offset, 5: size, 6: crc32c); `<prefix>.data-00000-of-00001` holds the raw little-endian tensors at those offsets.
"""
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(buf, offset, size):
    """Yield (key, value) of one table block (restart array at the end is not needed for a linear scan)."""
    block = buf[offset:offset + size]
    num_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _parse_proto(buf):
    """Minimal protobuf wire-format reader -> {field: [values]} (varint and length-delimited fields only + fixed32/64)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack("<I", buf[pos:pos + 4])[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack("<Q", buf[pos:pos + 8])[0]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _shape(shape_msg):
    dims = []
    for d in _parse_proto(shape_msg).get(2, []):  # TensorShapeProto.dim
        dims.append(_parse_proto(d).get(1, [0])[0])
    return tuple(dims)


def read_index(prefix):
    """-> {tensor name: (dtype enum, shape, shard, offset, size)} from `<prefix>.index`."""
    buf = open(prefix + ".index", "rb").read()
    if struct.unpack("<Q", buf[-8:])[0] != _MAGIC:
        raise ValueError(f"{prefix}.index is not a tensor-bundle table (bad magic)")
    footer = buf[-48:]
    _, p = _varint(footer, 0)       # metaindex handle
    _, p = _varint(footer, p)
    ioff, p = _varint(footer, p)    # index block handle
    isize, p = _varint(footer, p)
    entries = {}
    for _, handle in _block_entries(buf, ioff, isize):
        boff, q = _varint(handle, 0)
        bsize, q = _varint(handle, q)
        for key, value in _block_entries(buf, boff, bsize):
            if key == b"":
                continue  # BundleHeaderProto
            m = _parse_proto(value)
            entries[key.decode()] = (m.get(1, [0])[0], _shape(m[2][0]) if 2 in m else (), m.get(3, [0])[0],
                                     m.get(4, [0])[0], m.get(5, [0])[0])
    return entries


def read_tf_checkpoint(prefix, numeric_only=True):
    """-> {tensor name: np.ndarray} for every numeric tensor of a tf.train.Checkpoint (e.g. '.../ckpt-200')."""
    entries = read_index(prefix)
    shards = {}
    out = {}
    for name, (dtype, shape, shard, offset, size) in entries.items():
        if dtype not in _DTYPES:
            if numeric_only:
                continue
            raise ValueError(f"{name}: unsupported dtype enum {dtype}")
        if shard not in shards:
            shards[shard] = open(f"{prefix}.data-{shard:05d}-of-00001", "rb").read()
        raw = shards[shard][offset:offset + size]
        out[name] = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dtype]).newbyteorder("<")).reshape(shape).copy()
    return out

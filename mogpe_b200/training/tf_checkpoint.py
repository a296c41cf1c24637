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


# ---- writer (reference: tf.train.Checkpoint(model=model) saved through a CheckpointManager,
# mogpe/training/utils.py:228-242) -----------------------------------------------------------------------------------
# The same container the reader above parses: a LevelDB-format table `<prefix>.index` whose values are BundleEntryProto
# messages, raw little-endian tensors in `<prefix>.data-00000-of-00001`, and the object graph tf.train.Checkpoint.restore
# walks (`_CHECKPOINTABLE_OBJECT_GRAPH`, a TrackableObjectGraph proto in a scalar string tensor).  Checksums are CRC-32C
# (Castagnoli), masked as LevelDB does.  Every primitive is pinned against the reference's own checkpoint in
# tests/test_checkpoint_interop.py.
_DTYPE_ENUM = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9,
               np.dtype(np.bool_): 10}
_DT_STRING = 7


def _crc_table():
    poly, table = 0x82F63B78, []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _crc_table()


def crc32c(data, crc=0):
    """CRC-32C of `data`, continuing from `crc` (crc32c::Extend)."""
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c_np(arr_bytes, crc=0):
    """The same over a large buffer, 8 table steps per Python iteration would still be slow: use zlib-free slicing-by-1 in
    numpy chunks (tensors here are a few MB at most)."""
    return crc32c(arr_bytes, crc)


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


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


def _pb_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field, payload):
    return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


def _pb_fixed32(field, v):
    return _put_varint((field << 3) | 5) + struct.pack("<I", v)


def _entry_proto(dtype_enum, shape, offset, size, crc):
    shape_msg = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)  # TensorShapeProto.dim { size }
    # proto3: zero-valued scalar fields (offset 0, shard 0) are not serialised
    return (_pb_varint(1, dtype_enum) + _pb_bytes(2, shape_msg) + (_pb_varint(4, offset) if offset else b"") + _pb_varint(5, size) +
            _pb_fixed32(6, crc))


def encode_string_tensor(value: bytes):
    """Scalar DT_STRING tensor as the bundle stores it: varint length, masked CRC of the lengths (each as a little-endian
    uint32), the bytes.  -> (payload, masked crc32c for the BundleEntryProto)."""
    lengths = _put_varint(len(value))
    crc = crc32c(struct.pack("<I", len(value)))
    cks = struct.pack("<I", mask_crc(crc))
    crc = crc32c(cks, crc)
    crc = crc32c(value, crc)
    return lengths + cks + value, mask_crc(crc)


def _table_block(items, restart_interval=16):
    """One LevelDB block: prefix-compressed (key, value) entries, restart offsets, restart count."""
    out, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block):
    """block + 1 byte compression type (0 = none) + masked CRC-32C of both."""
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


def object_graph_proto(variable_keys):
    """TrackableObjectGraph for a set of checkpoint keys `a/b/c/.ATTRIBUTES/VARIABLE_VALUE`: one node per path prefix in
    breadth-first order (node 0 = the Checkpoint object), children by local name, one VARIABLE_VALUE attribute on every
    variable node.  Objects without variables further down (bijectors, mean functions without parameters, ...) are not
    listed: tf.train.Checkpoint.restore matches the live object graph against this one by child names and leaves
    unlisted children untouched."""
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    tree = {}
    for key in sorted(variable_keys):
        if not key.endswith(suffix):
            raise ValueError(f"not a checkpoint variable key: {key}")
        node = tree
        for part in key[:-len(suffix)].split("/"):
            node = node.setdefault(part, {})
        node[None] = key
    order, queue = [], [("", tree)]
    while queue:  # breadth first
        path, node = queue.pop(0)
        order.append((path, node))
        for name in node:
            if name is not None:
                queue.append((f"{path}/{name}" if path else name, node[name]))
    index = {path: i for i, (path, _) in enumerate(order)}
    nodes = b""
    for path, node in order:
        msg = b""
        for name in node:
            if name is not None:
                child = f"{path}/{name}" if path else name
                msg += _pb_bytes(1, _pb_varint(1, index[child]) + _pb_bytes(2, name.encode()))
        if None in node:
            full_name = path.split("/")[-1] if "/" in path else path
            msg += _pb_bytes(2, _pb_bytes(1, b"VARIABLE_VALUE") + _pb_bytes(2, full_name.encode()) + _pb_bytes(3, node[None].encode()))
        nodes += _pb_bytes(1, msg)
    return nodes


def parse_object_graph(payload):
    """TrackableObjectGraph bytes -> {checkpoint key: path of child names from the root} (for tests / inspection)."""
    g = _parse_proto(payload)
    nodes = [_parse_proto(n) for n in g.get(1, [])]
    out = {}

    def walk(i, path, seen):
        if i in seen:
            return
        n = nodes[i]
        for a in n.get(2, []):
            am = _parse_proto(a)
            out[am[3][0].decode()] = "/".join(path)
        for c in n.get(1, []):
            cm = _parse_proto(c)
            walk(cm.get(1, [0])[0], path + [cm[2][0].decode()], seen | {i})

    walk(0, [], frozenset())
    return out


def write_tf_checkpoint(prefix, tensors, save_counter=None):
    """Write {checkpoint key: np.ndarray} as `<prefix>.index` + `<prefix>.data-00000-of-00001` with the object graph of those
    keys.  Keys carry the '/.ATTRIBUTES/VARIABLE_VALUE' suffix (as read_tf_checkpoint returns them)."""
    tensors = dict(tensors)
    if save_counter is not None:
        tensors["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.asarray(save_counter, dtype=np.int64)
    data, entries = bytearray(), {}
    for key in sorted(tensors):
        a = np.ascontiguousarray(tensors[key])
        if a.dtype not in _DTYPE_ENUM:
            raise TypeError(f"{key}: dtype {a.dtype} cannot be stored")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        entries[key] = _entry_proto(_DTYPE_ENUM[a.dtype], a.shape, len(data), len(raw), mask_crc(crc32c(raw)))
        data += raw
    og, og_crc = encode_string_tensor(object_graph_proto(tensors.keys()))
    entries["_CHECKPOINTABLE_OBJECT_GRAPH"] = _entry_proto(_DT_STRING, (), len(data), len(og), og_crc)
    data += og
    # BundleHeaderProto { num_shards = 1; endianness = LITTLE (0, default); version { producer = 1 } } under the empty key
    items = [(b"", _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))] + [(k.encode(), entries[k]) for k in sorted(entries)]
    block = _with_trailer(_table_block(items))
    data_size = len(block) - 5
    meta = _with_trailer(_table_block([]))
    meta_off = len(block)
    index_block = _with_trailer(_table_block([(items[-1][0] + b"\x00", _put_varint(0) + _put_varint(data_size))]))
    index_off = meta_off + len(meta)
    footer = _put_varint(meta_off) + _put_varint(len(meta) - 5) + _put_varint(index_off) + _put_varint(len(index_block) - 5)
    footer = footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(block + meta + index_block + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    return sorted(entries)

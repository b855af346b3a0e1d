"""Reader / writer of TensorFlow `Saver` V2 checkpoints (TensorBundle) for the LCN variables -- SURVEY 8(f) rank 4.

The reference saves with tf.compat.v1.train.Saver(max_to_keep=1) (network/models_att.py:256-260,321-322) and restores
with tf.train.latest_checkpoint + Saver.restore (:445-463); tools/checkpoint_analysis.py lists a checkpoint's tensors.
TensorFlow is not installable here, so this module restates the published on-disk format [TF-sem: tensorflow 2.13,
core/util/tensor_bundle/tensor_bundle.{h,cc}, core/lib/io/{table_builder,block_builder,format}.cc,
core/protobuf/tensor_bundle.proto]:

  <prefix>.index                   an SSTable (LevelDB table format, no compression): sorted keys -> serialized protos.
                                   key ""        -> BundleHeaderProto {num_shards = 1, endianness = LITTLE, version}
                                   key <varname> -> BundleEntryProto  {dtype, shape, shard_id, offset, size, crc32c}
  <prefix>.data-00000-of-00001     the tensors' raw little-endian bytes, concatenated in key order
  <dir>/checkpoint                 text CheckpointState: model_checkpoint_path: "model-<step>"

  SSTable: data block(s) of prefix-compressed entries (varint32 shared, non_shared, value_len; key delta; value) with a
  restart array every 16 entries, each block followed by a 5-byte trailer (compression type 0 + masked CRC32C of block +
  type); an empty metaindex block; an index block mapping a separator key to each data block's (offset, size) handle;
  a 48-byte footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).

Variable names are the reference's: mask, linear_model/w1 ..., <bn layer>/{gamma,beta,moving_mean,
moving_variance}, global_step, the Adam slots <var>/Adam and <var>/Adam_1 and the optimizer's beta1_power /
beta2_power.  Parity status: UNPINNED against real TensorFlow output (no TF here, no checkpoint in the reference tree);
pinned against itself (round trip) and against the format's own checksums.  Pure host I/O: not on the hot path.
"""
import glob
import os
import re
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_NP_OF = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT_OF = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}
RESTART_INTERVAL = 16

# ---- CRC32C (Castagnoli), table driven, vectorised over 8-byte strides is not needed: checkpoints here are < 200 MB ----
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        # slicing-by-8 tables
        tabs = np.zeros((8, 256), dtype=np.uint32)
        tabs[0] = t
        for k in range(1, 8):
            tabs[k] = (tabs[k - 1] >> 8) ^ t[tabs[k - 1] & 0xFF]
        _CRC_TABLE = tabs
    return _CRC_TABLE


def crc32c(data, crc=0):
    """CRC32C of bytes-like `data`: lcn_crc32c of liblcn_b200.so (host code, GB/s); crc32c_py when the library is absent."""
    buf = np.frombuffer(memoryview(data), dtype=np.uint8)
    try:
        from .. import _lib
        lib = _lib.load()
    except Exception:
        return crc32c_py(buf, crc)
    if buf.size == 0:
        return int(crc)
    buf = np.ascontiguousarray(buf)
    return int(lib.lcn_crc32c(buf.ctypes.data, buf.size, int(crc)))


def crc32c_py(data, crc=0):
    """The same checksum in NumPy / Python (slicing-by-8): the cross-check of lcn_crc32c in the CPU tests."""
    tabs = _crc_table()
    buf = np.frombuffer(memoryview(data), dtype=np.uint8)
    c = np.uint32(crc ^ 0xFFFFFFFF)
    n8 = len(buf) // 8
    if n8:
        words = buf[: n8 * 8].reshape(n8, 8)
        t = [tabs[k] for k in range(8)]
        c = int(c)
        # the dependency chain is serial; keep it in Python ints but 8 bytes per iteration
        w = words.astype(np.uint32)
        lo = (w[:, 0] | (w[:, 1] << 8) | (w[:, 2] << 16) | (w[:, 3] << 24)).tolist()
        b4, b5, b6, b7 = w[:, 4].tolist(), w[:, 5].tolist(), w[:, 6].tolist(), w[:, 7].tolist()
        t0, t1, t2, t3, t4, t5, t6, t7 = (x.tolist() for x in t)
        for i in range(n8):
            x = c ^ lo[i]
            c = (t7[x & 0xFF] ^ t6[(x >> 8) & 0xFF] ^ t5[(x >> 16) & 0xFF] ^ t4[x >> 24] ^
                 t3[b4[i]] ^ t2[b5[i]] ^ t1[b6[i]] ^ t0[b7[i]])
    else:
        c = int(c)
    t0 = tabs[0].tolist() if len(buf) % 8 else None
    for b in buf[n8 * 8:].tolist():
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    """crc32c::Mask: rotate right by 15 and add a constant (stored CRCs of data that embeds CRCs)."""
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf wire format ----
def _varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _read_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def _parse_fields(buf):
    """-> list of (field number, wire type, value) with value = int (varint / fixed) or bytes (length delimited)."""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            v = bytes(buf[pos: pos + ln])
            pos += ln
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.append((num, wire, v))
    return out


def _encode_header():
    # BundleHeaderProto: num_shards (1) = 1, endianness (2) = LITTLE (0, default: omitted), version (3) = VersionDef{producer (1) = 1}
    return _field(1, 0, _varint(1)) + _field(3, 2, _varint(2) + _field(1, 0, _varint(1)))


def _encode_entry(dtype, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _varint(len(d)) + d) for d in (_field(1, 0, _varint(s)) for s in shape))
    out = _field(1, 0, _varint(dtype)) + _field(2, 2, _varint(len(dims)) + dims)
    # shard_id (3) = 0 is the proto3 default and is omitted, like TensorFlow's serializer does
    if offset:
        out += _field(4, 0, _varint(offset))
    out += _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


def _decode_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    for num, _, v in _parse_fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            for n2, _, d in _parse_fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, x in _parse_fields(d):
                        if n3 == 1:
                            size = x
                    e["shape"].append(size)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = v
    return e


# ---- SSTable ----
def _build_block(items):
    """items: sorted list of (key bytes, value bytes) -> block contents (without the trailer)."""
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % RESTART_INTERVAL == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block):
    trailer_type = b"\x00"                                       # kNoCompression
    return block + trailer_type + struct.pack("<I", mask_crc(crc32c(block + trailer_type)))


def _parse_block(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos: pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos: pos + vlen])))
        pos += vlen
    return out


def write_table(path, items):
    """items: dict or list of (key str/bytes, value bytes); written as a one-data-block-per-256-KiB SSTable."""
    items = sorted(((k.encode() if isinstance(k, str) else k, v) for k, v in (items.items() if isinstance(items, dict) else items)))
    blocks, cur, cur_bytes = [], [], 0
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= (256 << 10):
            blocks.append(cur)
            cur, cur_bytes = [], 0
    if cur or not blocks:
        blocks.append(cur)
    out, index = bytearray(), []
    for blk in blocks:
        body = _build_block(blk)
        index.append((blk[-1][0] if blk else b"", _varint(len(out)) + _varint(len(body))))
        out += _with_trailer(body)
    meta_off = len(out)
    meta = _build_block([])
    out += _with_trailer(meta)
    idx_off = len(out)
    idx = _build_block(index)
    out += _with_trailer(idx)
    footer = _varint(meta_off) + _varint(len(meta)) + _varint(idx_off) + _varint(len(idx))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out += footer
    with open(path, "wb") as f:
        f.write(bytes(out))


def read_table(path, verify=True):
    """-> list of (key bytes, value bytes) in key order.  Verifies the magic and (verify=True) every block CRC."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s: not an SSTable (bad magic)" % path)
    footer = buf[-48:]
    _, pos = _read_varint(footer, 0)
    _, pos = _read_varint(footer, pos)
    idx_off, pos = _read_varint(footer, pos)
    idx_size, pos = _read_varint(footer, pos)

    def block_at(off, size):
        body = buf[off: off + size]
        if buf[off + size] != 0:
            raise ValueError("%s: compressed blocks are not supported" % path)
        if verify:
            want = unmask_crc(struct.unpack_from("<I", buf, off + size + 1)[0])
            if crc32c(buf[off: off + size + 1]) != want:
                raise ValueError("%s: block checksum mismatch at offset %d" % (path, off))
        return body
    out = []
    for _, handle in _parse_block(block_at(idx_off, idx_size)):
        off, p2 = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p2)
        out += _parse_block(block_at(off, size))
    return out


# ---- TensorBundle ----
def write_bundle(prefix, tensors):
    """tensors: dict name -> ndarray (float32 / int32 / int64).  Writes <prefix>.index and <prefix>.data-00000-of-00001."""
    entries, offset = {"": _encode_header()}, 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])                     # (ascontiguousarray would turn a scalar into shape (1,))
            dt = _DT_OF[a.dtype]
            raw = a.astype(_NP_OF[dt], copy=False).tobytes(order="C")
            f.write(raw)
            entries[name] = _encode_entry(dt, a.shape, offset, len(raw), mask_crc(crc32c(raw)))
            offset += len(raw)
    write_table(prefix + ".index", entries)


def list_bundle(prefix):
    """tools/checkpoint_analysis.py: -> dict name -> (dtype code, shape).  Reads only the index."""
    out = {}
    for k, v in read_table(prefix + ".index"):
        if k:
            e = _decode_entry(v)
            out[k.decode()] = (e["dtype"], tuple(e["shape"]))
    return out


def read_bundle(prefix, names=None, verify=True):
    """-> dict name -> ndarray.  names: optional subset.  Verifies every tensor's CRC32C when verify=True."""
    out = {}
    table = read_table(prefix + ".index", verify)
    if not table or table[0][0] != b"":
        raise ValueError("%s.index: no bundle header" % prefix)
    hdr = {num: v for num, _, v in _parse_fields(table[0][1])}
    n_shards = hdr.get(1, 0)
    if hdr.get(2, 0) != 0:
        raise ValueError("big-endian bundles are not supported")
    shards = {}
    for k, v in table[1:]:
        name = k.decode()
        if names is not None and name not in names:
            continue
        e = _decode_entry(v)
        if e["dtype"] not in _NP_OF:
            continue                                         # e.g. strings of a SavedModel: not LCN variables
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, n_shards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]: e["offset"] + e["size"]]
        if verify and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError("%s: tensor %s checksum mismatch" % (prefix, name))
        out[name] = np.frombuffer(raw.tobytes(), dtype=_NP_OF[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ---- checkpoint state file + the LCN variable set ----
def write_checkpoint_state(directory, name):
    with open(os.path.join(directory, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by <dir>/checkpoint, else the highest model-<step>.index present."""
    state = os.path.join(directory, "checkpoint")
    if os.path.exists(state):
        m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(directory, p)
            if os.path.exists(p + ".index"):
                return p
    files = glob.glob(os.path.join(directory, "model-*.index"))
    if not files:
        return None
    return max(files, key=lambda f: int(re.search(r"model-(\d+)\.index$", f).group(1)))[: -len(".index")]


ADAM_SCOPES = ("training/", "")          # where beta1_power / beta2_power live: inside name_scope("training") [TF-sem]


def save_model(directory, step, params, state, tensor_table, beta1=0.9, beta2=0.999):
    """Saver(max_to_keep=1).save(sess, <dir>/model, global_step=step) for the LCN variable set.
    params: name -> ndarray (LcnEngine.get_params); state: LcnEngine.get_state() (flat Adam slots, global_step)."""
    os.makedirs(directory, exist_ok=True)
    for old in glob.glob(os.path.join(directory, "model-*")):                 # max_to_keep=1
        os.remove(old)
    t = {}
    gs = int(state["global_step"]) if state is not None else int(step)
    for name, (off, rows, cols) in tensor_table.items():
        shape = (rows, cols) if rows > 1 else (cols,)
        t[name] = np.asarray(params[name], dtype=np.float32).reshape(shape)
        if state is not None:
            t[name + "/Adam"] = np.asarray(state["adam_m"][off: off + rows * cols], dtype=np.float32).reshape(shape)
            t[name + "/Adam_1"] = np.asarray(state["adam_v"][off: off + rows * cols], dtype=np.float32).reshape(shape)
        if name.endswith("/gamma"):      # Keras BN's non-trainable pair: never updated by the reference
            t[name[: -len("gamma")] + "moving_mean"] = np.zeros(shape, np.float32)
            t[name[: -len("gamma")] + "moving_variance"] = np.ones(shape, np.float32)
    t["global_step"] = np.asarray(gs, dtype=np.int32)
    if state is not None:
        t["training/beta1_power"] = np.asarray(beta1 ** (gs + 1), dtype=np.float32)   # TF1 Adam: beta^(t+1) after t steps
        t["training/beta2_power"] = np.asarray(beta2 ** (gs + 1), dtype=np.float32)
        if "loss_ema" in state:
            t["lcn_b200/loss_ema"] = np.asarray(state["loss_ema"], dtype=np.float32)
    name = "model-%d" % step
    write_bundle(os.path.join(directory, name), t)
    write_checkpoint_state(directory, name)
    return os.path.join(directory, name)


def load_model(prefix, tensor_table, n_params):
    """-> (params dict, state dict or None).  Accepts checkpoints written by save_model and by the reference's Saver
    (same variable names); Adam slots are optional (a weights-only bundle restores the variables and leaves the
    optimizer state untouched)."""
    t = read_bundle(prefix)
    params = {}
    for name, (off, rows, cols) in tensor_table.items():
        if name not in t:
            raise KeyError("checkpoint %s has no variable %r" % (prefix, name))
        params[name] = t[name]
    have_slots = all((name + "/Adam") in t and (name + "/Adam_1") in t for name in tensor_table)
    if not have_slots:
        return params, None
    m, v = np.zeros(n_params, np.float32), np.zeros(n_params, np.float32)
    for name, (off, rows, cols) in tensor_table.items():
        m[off: off + rows * cols] = t[name + "/Adam"].reshape(-1)
        v[off: off + rows * cols] = t[name + "/Adam_1"].reshape(-1)
    state = {"adam_m": m, "adam_v": v, "global_step": int(t["global_step"]) if "global_step" in t else 0}
    if "lcn_b200/loss_ema" in t:
        state["loss_ema"] = t["lcn_b200/loss_ema"]
    return params, state

"""Reader (and a small writer) for TensorFlow's checkpoint format V2 ("tensor bundle"), in pure Python / NumPy.

The reference trains with TF1 Estimators, whose checkpoints are `model.ckpt-<step>.index` + `.data-00000-of-0000N`;
its evaluation hand-off loads such a checkpoint into the exported generator (compare_gan/gans/modular_gan.py:266-285,
eval_gan_lib.py:156-163).  TensorFlow is not installable here, so the format is restated from its specification:

* `.index` is an SSTable in LevelDB's table format (tensorflow/core/lib/io/table*.cc): data blocks of prefix-compressed
  entries `varint32 shared | varint32 non_shared | varint32 value_len | key_delta | value` followed by a restart array
  and its length, each block trailed by 1 compression byte (0 = none, 1 = snappy) and a masked CRC-32C; an index block
  mapping separator keys to block handles (varint64 offset, varint64 size); a 48-byte footer = metaindex handle, index
  handle, padding, magic 0xdb4775248b80fb57.
* the empty key holds a BundleHeaderProto (num_shards = 1, endianness = 2, version = 3), every other key is a variable name
  mapped to a BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto): dtype = 1, shape = 2 (TensorShapeProto,
  dim = 2 {size = 1}), shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32, masked).
* `.data-XXXXX-of-YYYYY` holds the raw little-endian tensor bytes.

Variable names are the checkpoint key space this package already uses (`generator/B1/up_conv1/kernel`,
`.../kernel/u_var`, `<var>/Adam`, `<var>/Adam_1`, `<var>/ExponentialMovingAverage`, `global_step`, `global_step_disc`),
so `ModularGAN.load_checkpoint` accepts a TF checkpoint prefix as well as its own `.npz`.  PARITY UNPINNED: no
TF-written file is available offline; the reader is exercised against files produced by the writer below (same spec) and
by hand-assembled byte strings in tests/test_tf_checkpoint.py.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_CODES = {np.dtype(v): k for k, v in DTYPES.items()}


# ---- CRC-32C (Castagnoli), masked as LevelDB / TF do -----------------------------------------------------------------

def _crc_table():
  poly = 0x82F63B78
  t = np.zeros(256, np.uint32)
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ poly if c & 1 else c >> 1
    t[i] = c
  return t


_TABLE = _crc_table()


def crc32c(data, crc=0):
  crc ^= 0xFFFFFFFF
  table = _TABLE
  for b in bytes(data):
    crc = int(table[(crc ^ b) & 0xFF]) ^ (crc >> 8)
  return crc ^ 0xFFFFFFFF


def mask_crc(crc):
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---- varints / protobuf wire format ---------------------------------------------------------------------------------

def _read_varint(buf, pos):
  out = shift = 0
  while True:
    b = buf[pos]
    pos += 1
    out |= (b & 0x7F) << shift
    if not b & 0x80:
      return out, pos
    shift += 7


def _varint(v):
  out = bytearray()
  v &= (1 << 64) - 1
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _parse_message(buf):
  """{field: [values]} of one protobuf message (varint, fixed32/64 and length-delimited fields)."""
  fields, pos = {}, 0
  while pos < len(buf):
    key, pos = _read_varint(buf, pos)
    field, wire = key >> 3, key & 7
    if wire == 0:
      v, pos = _read_varint(buf, pos)
    elif wire == 1:
      v = struct.unpack_from("<Q", buf, pos)[0]
      pos += 8
    elif wire == 2:
      n, pos = _read_varint(buf, pos)
      v = bytes(buf[pos:pos + n])
      pos += n
    elif wire == 5:
      v = struct.unpack_from("<I", buf, pos)[0]
      pos += 4
    else:
      raise ValueError("unsupported protobuf wire type %d" % wire)
    fields.setdefault(field, []).append(v)
  return fields


def _signed64(v):
  return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(buf):
  f = _parse_message(buf)
  shape = []
  if 2 in f:
    for dim in _parse_message(f[2][0]).get(2, []):
      shape.append(_signed64(_parse_message(dim).get(1, [0])[0]))
  if 7 in f:
    raise ValueError("sliced (partitioned) variables are not supported")
  return {"dtype": f.get(1, [0])[0], "shape": tuple(shape), "shard_id": f.get(3, [0])[0], "offset": f.get(4, [0])[0],
          "size": f.get(5, [0])[0], "crc32c": f.get(6, [None])[0]}


# ---- SSTable ---------------------------------------------------------------------------------------------------------

def _read_block(data, offset, size, verify):
  raw = data[offset:offset + size]
  ctype = data[offset + size]
  if verify:
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
      raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
  if ctype == 1:
    raw = _snappy_decompress(raw)
  elif ctype != 0:
    raise ValueError("checkpoint index: unknown block compression %d" % ctype)
  n_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
  limit = len(raw) - 4 - 4 * n_restarts
  entries, pos, key = [], 0, b""
  while pos < limit:
    shared, pos = _read_varint(raw, pos)
    non_shared, pos = _read_varint(raw, pos)
    vlen, pos = _read_varint(raw, pos)
    key = key[:shared] + bytes(raw[pos:pos + non_shared])
    pos += non_shared
    entries.append((key, bytes(raw[pos:pos + vlen])))
    pos += vlen
  return entries


def _snappy_decompress(buf):
  n, pos = _read_varint(buf, 0)
  out = bytearray()
  while pos < len(buf):
    tag = buf[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(buf[pos:pos + nb], "little")
        pos += nb
      ln += 1
      out += buf[pos:pos + ln]
      pos += ln
    else:
      if kind == 1:
        ln = ((tag >> 2) & 7) + 4
        off = ((tag >> 5) << 8) | buf[pos]
        pos += 1
      elif kind == 2:
        ln = (tag >> 2) + 1
        off = int.from_bytes(buf[pos:pos + 2], "little")
        pos += 2
      else:
        ln = (tag >> 2) + 1
        off = int.from_bytes(buf[pos:pos + 4], "little")
        pos += 4
      for _ in range(ln):
        out.append(out[-off])
  if len(out) != n:
    raise ValueError("snappy: length mismatch")
  return bytes(out)


def read_index(prefix, verify=True):
  """{name: entry dict} plus the header under the key ""."""
  data = open(prefix + ".index", "rb").read()
  if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != MAGIC:
    raise ValueError("%s.index is not a TensorFlow V2 checkpoint index (bad magic)" % prefix)
  footer = data[len(data) - 48:]
  _, pos = _read_varint(footer, 0)        # metaindex handle (unused)
  _, pos = _read_varint(footer, pos)
  ioff, pos = _read_varint(footer, pos)
  isize, pos = _read_varint(footer, pos)
  out = {}
  for _, handle in _read_block(data, ioff, isize, verify):
    boff, p = _read_varint(handle, 0)
    bsize, p = _read_varint(handle, p)
    for key, value in _read_block(data, boff, bsize, verify):
      if key == b"":
        h = _parse_message(value)
        out[""] = {"num_shards": h.get(1, [1])[0], "endianness": h.get(2, [0])[0]}
      else:
        out[key.decode()] = _parse_entry(value)
  return out


def list_variables(prefix):
  """[(name, shape)] like tf.train.list_variables."""
  return [(k, list(v["shape"])) for k, v in sorted(read_index(prefix).items()) if k]


def load_checkpoint(prefix, names=None, verify_tensors=False):
  """{name: numpy array} for `names` (default: all) of the checkpoint `prefix` (path without `.index`)."""
  index = read_index(prefix)
  header = index.pop("", {"num_shards": 1, "endianness": 0})
  if header.get("endianness", 0) not in (0,):
    raise ValueError("big-endian checkpoints are not supported")
  shards, out = {}, {}
  for name, e in index.items():
    if names is not None and name not in names:
      continue
    if e["dtype"] not in DTYPES:
      continue                                      # strings / resources: not model state
    sid = e["shard_id"]
    if sid not in shards:
      path = "%s.data-%05d-of-%05d" % (prefix, sid, header.get("num_shards", 1))
      shards[sid] = np.memmap(path, dtype=np.uint8, mode="r")
    raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
    if verify_tensors and e["crc32c"] is not None and mask_crc(crc32c(raw.tobytes())) != e["crc32c"]:
      raise ValueError("checkpoint tensor %s: checksum mismatch" % name)
    dt = np.dtype(DTYPES[e["dtype"]])
    n = int(np.prod(e["shape"])) if e["shape"] else 1
    if n * dt.itemsize != e["size"]:
      raise ValueError("checkpoint tensor %s: %d bytes for shape %s %s" % (name, e["size"], e["shape"], dt))
    out[name] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e["shape"]).copy()
  return out


def latest_checkpoint(model_dir):
  """tf.train.latest_checkpoint without the `checkpoint` state file: the `model.ckpt-<step>` with the largest step."""
  best = None
  for f in os.listdir(model_dir):
    if f.startswith("model.ckpt-") and f.endswith(".index"):
      try:
        step = int(f[len("model.ckpt-"):-len(".index")])
      except ValueError:
        continue
      if best is None or step > best[0]:
        best = (step, os.path.join(model_dir, f[:-len(".index")]))
  return None if best is None else best[1]


# ---- writer (tests, and exporting a trained model for the reference's tooling) ------------------------------------------

def _field(num, wire, payload):
  return _varint((num << 3) | wire) + payload


def _entry_proto(dtype_code, shape, offset, size, crc):
  dims = b"".join(_field(2, 2, _varint(len(d)) + d) for d in (_field(1, 0, _varint(s)) for s in shape))
  msg = _field(1, 0, _varint(dtype_code))
  msg += _field(2, 2, _varint(len(dims)) + dims)
  if offset:
    msg += _field(4, 0, _varint(offset))
  msg += _field(5, 0, _varint(size))
  msg += _field(6, 5, struct.pack("<I", crc))
  return msg


def _block(entries, restart_interval=16):
  out, restarts, prev = bytearray(), [], b""
  for i, (key, value) in enumerate(entries):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(out))
    else:
      while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
        shared += 1
    out += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
    prev = key
  if not restarts:
    restarts = [0]
  for r in restarts:
    out += struct.pack("<I", r)
  out += struct.pack("<I", len(restarts))
  return bytes(out)


def save_checkpoint(prefix, tensors, block_entries=64):
  """Writes {name: array} as `prefix.index` + `prefix.data-00000-of-00001` (uncompressed blocks, checksums filled in)."""
  names = sorted(tensors, key=lambda s: s.encode())
  data, entries = bytearray(), [(b"", _field(1, 0, _varint(1)) + _field(3, 2, _varint(2) + _field(1, 0, _varint(1))))]
  for name in names:
    a = np.asarray(tensors[name])
    if not a.flags.c_contiguous:        # (np.ascontiguousarray would turn a scalar into shape [1])
      a = np.ascontiguousarray(a)
    if a.dtype not in DTYPE_CODES:
      raise ValueError("unsupported dtype %s for %s" % (a.dtype, name))
    raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    entries.append((name.encode(), _entry_proto(DTYPE_CODES[a.dtype], a.shape, len(data), len(raw), mask_crc(crc32c(raw)))))
    data += raw
  out, index_entries = bytearray(), []

  def emit(block):
    off = len(out)
    trailer = b"\x00"
    out.extend(block + trailer + struct.pack("<I", mask_crc(crc32c(block + trailer))))
    return _varint(off) + _varint(len(block))
  for i in range(0, len(entries), block_entries):
    chunk = entries[i:i + block_entries]
    handle = emit(_block(chunk))
    index_entries.append((chunk[-1][0] + b"\x00" if i + block_entries < len(entries) else chunk[-1][0] + b"\xff", handle))
  meta = emit(_block([]))
  index = emit(_block(index_entries, restart_interval=1))
  footer = meta + index
  footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
  out.extend(footer)
  with open(prefix + ".index", "wb") as f:
    f.write(bytes(out))
  with open(prefix + ".data-00000-of-00001", "wb") as f:
    f.write(bytes(data))
  return prefix

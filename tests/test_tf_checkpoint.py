"""TensorFlow checkpoint-V2 ("tensor bundle") reader of compare_gan_b200/tf_checkpoint.py: hand-assembled byte strings per
the format's specification, round trips through the writer, checksum verification, and the name-compatible hand-off
into the engine's checkpoint key space.  No TF-written file exists offline: parity with real TF files is unpinned."""
import os
import struct

import numpy as np
import pytest

from compare_gan_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
  # RFC 3720 B.4 test vectors for CRC-32C
  assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA
  assert tfc.crc32c(b"\xff" * 32) == 0x62A8AB43
  assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
  assert tfc.crc32c(b"123456789") == 0xE3069283
  # LevelDB's mask: rotate right by 15, add 0xa282ead8
  assert tfc.mask_crc(0) == 0xa282ead8


def test_varint_and_entry_proto_by_hand():
  # BundleEntryProto { dtype: DT_FLOAT(1), shape { dim { size: 3 } dim { size: 2 } }, offset: 300, size: 24, crc32c: 0x01020304 }
  raw = bytes([0x08, 0x01,                                   # field 1 varint 1
               0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x02,   # field 2 len 8: two dims
               0x20, 0xAC, 0x02,                             # field 4 varint 300
               0x28, 0x18,                                   # field 5 varint 24
               0x35, 0x04, 0x03, 0x02, 0x01])                # field 6 fixed32
  e = tfc._parse_entry(raw)
  assert e == {"dtype": 1, "shape": (3, 2), "shard_id": 0, "offset": 300, "size": 24, "crc32c": 0x01020304}
  assert tfc._entry_proto(1, (3, 2), 300, 24, 0x01020304) == raw


def test_block_prefix_compression_by_hand():
  # two entries "abc"->"1", "abd"->"22" with one restart: shared prefix "ab" on the second entry
  blk = bytes([0, 3, 1]) + b"abc" + b"1" + bytes([2, 1, 2]) + b"d" + b"22" + struct.pack("<I", 0) + struct.pack("<I", 1)
  data = blk + b"\x00" + struct.pack("<I", tfc.mask_crc(tfc.crc32c(blk + b"\x00")))
  assert tfc._read_block(data, 0, len(blk), True) == [(b"abc", b"1"), (b"abd", b"22")]
  assert tfc._block([(b"abc", b"1"), (b"abd", b"22")]) == blk
  corrupted = bytearray(data)
  corrupted[4] ^= 1
  with pytest.raises(ValueError):
    tfc._read_block(bytes(corrupted), 0, len(blk), True)


def test_snappy_block():
  # literal "abcd" + copy (offset 4, length 4) -> "abcdabcd"
  comp = bytes([8, (4 - 1) << 2]) + b"abcd" + bytes([((4 - 4) << 2) | 1 | (0 << 5), 4])
  assert tfc._snappy_decompress(comp) == b"abcdabcd"


def test_round_trip_many_variables(tmp_path):
  rng = np.random.RandomState(0)
  tensors = {"generator/B%d/up_conv1/kernel" % i: rng.randn(3, 3, 4, 5).astype(np.float32) for i in range(150)}
  tensors["global_step"] = np.array(1234, np.int64)
  tensors["generator/fc_noise/kernel/u_var"] = rng.randn(7, 1).astype(np.float32)
  tensors["beta1_power"] = np.array(0.5, np.float32)
  prefix = tfc.save_checkpoint(str(tmp_path / "model.ckpt-1234"), tensors)
  assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
  listed = dict(tfc.list_variables(prefix))
  assert listed["global_step"] == [] and listed["generator/B7/up_conv1/kernel"] == [3, 3, 4, 5] and len(listed) == len(tensors)
  back = tfc.load_checkpoint(prefix, verify_tensors=True)
  assert sorted(back) == sorted(tensors)
  for k, v in tensors.items():
    np.testing.assert_array_equal(back[k], v)
    assert back[k].dtype == v.dtype
  only = tfc.load_checkpoint(prefix, names={"global_step"})
  assert list(only) == ["global_step"] and int(only["global_step"]) == 1234
  assert tfc.latest_checkpoint(str(tmp_path)) == prefix
  # a flipped tensor byte is caught by the per-tensor checksum
  with open(prefix + ".data-00000-of-00001", "r+b") as f:
    f.seek(10)
    b = f.read(1)
    f.seek(10)
    f.write(bytes([b[0] ^ 0xFF]))
  with pytest.raises(ValueError):
    tfc.load_checkpoint(prefix, verify_tensors=True)


def test_bad_magic(tmp_path):
  p = str(tmp_path / "x")
  open(p + ".index", "wb").write(b"\x00" * 64)
  with pytest.raises(ValueError):
    tfc.read_index(p)


def test_engine_loads_a_tensorflow_format_checkpoint(tmp_path):
  """The eval hand-off (modular_gan.py:266-285): a checkpoint in TF's own format, under the reference's variable names,
  restores the engine's weights, Adam slots, EMA shadows and counters (host code above the emulated ABI)."""
  from tests.abi_emulator import emulated_library
  from tests.gpu_util import make_inputs, make_pair
  with emulated_library():
    eng, _ = make_pair("resnet_cifar_arch", (32, 32, 3), 2, d_sn=True, disc_iters=1, g_use_ema=True, ema_start_step=0)
    rng = np.random.RandomState(0)
    eng.set_inputs(*make_inputs(rng, 1, 2, (32, 32, 3), 128))
    eng.run_cycle()
    eng.read_losses()
    want = eng.checkpoint_dict()
    prefix = tfc.save_checkpoint(str(tmp_path / "model.ckpt-1"), want)
    names = dict(tfc.list_variables(prefix))
    for k in ("generator/B1/up_conv1/kernel", "generator/B1/up_conv1/kernel/Adam_1", "discriminator/B1/same_conv1/kernel/u_var",
              "generator/B1/up_conv1/kernel/ExponentialMovingAverage", "generator/B1/bn1/moving_variance", "global_step"):
      assert k in names, k
    eng.store.vars["generator/fc_noise/kernel"].t.zero_()
    eng.g_opt.m.t.zero_()
    eng.ema.t.zero_()
    eng.g_opt.step.fill_(0)
    for spelling in (prefix, prefix + ".index", str(tmp_path)):
      eng.load_checkpoint(spelling)
      got = eng.checkpoint_dict()
      for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    os.remove(prefix + ".index")
    tfc.save_checkpoint(prefix, {k: v for k, v in want.items() if k != "generator/fc_noise/kernel"})
    with pytest.raises(ValueError):
      eng.load_checkpoint(prefix)

"""Host-side input pipeline (`cgan_loader_*`, csrc/loader.cu) against an independent Python model of the reference's
tf.data chain (datasets.py:261-291): repeat -> shuffle(buffer, seed) -> batch(drop_remainder) -> prefetch.  No GPU."""
import threading
import time

import numpy as np
import pytest

from compare_gan_b200 import _lib, datasets

M64 = (1 << 64) - 1


def splitmix64(state):
  state = (state + 0x9E3779B97F4A7C15) & M64
  z = state
  z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
  z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
  return state, z ^ (z >> 31)


def model_stream(n, buffer_size, seed, count):
  """tf.data shuffle semantics on the repeat() stream 0,1,..,n-1,0,1,..: the buffer holds the next `buffer_size` stream
  elements; every output is drawn uniformly from it and replaced by the next stream element."""
  nxt, out, state = 0, [], seed
  if buffer_size <= 1:
    return [(i % n) for i in range(count)]
  buf = []
  for _ in range(buffer_size):
    buf.append(nxt)
    nxt = (nxt + 1) % n
  limit = M64 - M64 % buffer_size
  for _ in range(count):
    while True:
      state, r = splitmix64(state)
      if r < limit:
        break
    slot = r % buffer_size
    out.append(buf[slot])
    buf[slot] = nxt
    nxt = (nxt + 1) % n
  return out


def make_source(n=37, h=4, w=3, c=3, seed=0):
  rng = np.random.RandomState(seed)
  images = rng.randint(0, 256, size=(n, h, w, c)).astype(np.uint8)
  images[:, 0, 0, 0] = np.arange(n) % 256          # element id readable from the data
  return images, np.arange(n, dtype=np.int32) * 3


@pytest.mark.parametrize("buffer_size,seed", [(0, 1), (8, 1), (8, 2), (37, 5), (100, 7)])
def test_stream_matches_tf_data_model(buffer_size, seed):
  images, labels = make_source()
  n, batch, nb = len(images), 5, 40
  it = datasets.BatchIterator(images, labels, batch, buffer_size, seed, ring=3)
  got_ids, got_labels = [], []
  for _ in range(nb):
    x, l = next(it)
    assert x.shape == (batch, 4, 3, 3) and x.dtype == np.float32 and l.dtype == np.int32
    ids = l // 3
    np.testing.assert_array_equal(x, images[ids].astype(np.float32) / np.float32(255.0))   # _parse_fn, bit-exact
    got_ids += list(ids)
    got_labels += list(l)
    it.release(1)
  it.close()
  assert got_ids == model_stream(n, buffer_size, seed, nb * batch)          # drop_remainder: batches never straddle a gap
  assert got_labels == [3 * i for i in got_ids]
  if buffer_size > 1:
    # shuffle-buffer locality: stream element j cannot be emitted before output position j - buffer + 1
    first_seen = {}
    for pos, e in enumerate(got_ids):
      first_seen.setdefault(e, pos)
    assert all(pos >= e - buffer_size + 1 for e, pos in first_seen.items())
    assert got_ids != [(i % n) for i in range(nb * batch)]


def test_seed_changes_order_and_float_sources_are_copied():
  images, labels = make_source()
  a = datasets.BatchIterator(images, labels, 6, 16, 11, ring=2)
  b = datasets.BatchIterator(images, labels, 6, 16, 12, ring=2)
  la, lb = next(a)[1].copy(), next(b)[1].copy()
  assert (la != lb).any()
  a.close(); b.close()
  f = np.random.RandomState(3).rand(10, 2, 2, 1).astype(np.float32)
  it = datasets.BatchIterator(f, None, 4, 0, 0, ring=2)
  x, l = next(it)
  np.testing.assert_array_equal(x, f[:4])
  assert not l.any()                                   # no source labels -> zeros
  it.close()


def test_ring_protocol_and_errors():
  images, labels = make_source()
  it = datasets.BatchIterator(images, labels, 4, 0, 0, ring=3)
  held = [next(it) for _ in range(3)]
  snapshot = [h[0].copy() for h in held]
  time.sleep(0.05)                                     # the producer must not touch outstanding slots
  for h, s in zip(held, snapshot):
    np.testing.assert_array_equal(h[0], s)
  with pytest.raises(_lib.CganError, match="outstanding"):
    next(it)
  with pytest.raises(_lib.CganError, match="more slots"):
    it.release(4)
  it.release(3)
  x, l = next(it)                                      # stream continues where it stopped: elements 12..15
  assert list(l // 3) == [12, 13, 14, 15]
  it.close()
  with pytest.raises(ValueError):
    datasets.BatchIterator(images.astype(np.int16), labels, 4, 0, 0, ring=3)


def test_producer_consumer_stress():
  images, labels = make_source(n=101)
  it = datasets.BatchIterator(images, labels, 7, 32, 9, ring=4)
  expect = model_stream(101, 32, 9, 7 * 300)
  rng = np.random.RandomState(0)
  got, outstanding = [], 0
  for i in range(300):
    x, l = next(it)
    got += list(l // 3)
    outstanding += 1
    if rng.rand() < 0.3:
      time.sleep(0.001)
    if outstanding == 3 or rng.rand() < 0.5:
      it.release(outstanding)
      outstanding = 0
  assert got == expect
  t = threading.Thread(target=it.close)               # close() while the producer waits for a free slot must not hang
  t.start(); t.join(5)
  assert not t.is_alive()


def test_dataset_input_fns():
  ds = datasets.get_dataset("cifar10")
  with pytest.raises(ValueError):
    ds.train_input_fn({})
  it = ds.train_input_fn({"batch_size": 8})
  x, l = next(it)
  fake_images, _ = ds._make_fake_dataset("train")
  assert x.shape == (8, 32, 32, 3) and (l == 1).all()            # fake data set: all-ones labels (datasets.py:143)
  for row in x:                                                   # every batch row is one of the 100 fake images
    assert (np.abs(fake_images - row).reshape(100, -1).max(1) == 0).any()
  it.close()
  ev = ds.eval_input_fn({"batch_size": 64})
  batches = 0
  for x, l in ev:
    batches += 1
    ev.release(1)
  assert batches == ds.eval_test_samples // 64
  ev.close()


def test_npy_shards_from_data_dir(tmp_path):
  images, labels = make_source(n=50, h=32, w=32, c=3)
  np.save(str(tmp_path / "cifar10_train_images.npy"), images)
  np.save(str(tmp_path / "cifar10_train_labels.npy"), labels)
  ds = datasets.get_dataset("cifar10", fake_dataset=False, data_dir=str(tmp_path), shuffle_buffer_size=0)
  it = ds.train_input_fn({"batch_size": 10})
  x, l = next(it)
  np.testing.assert_array_equal(l, labels[:10])
  np.testing.assert_array_equal(x, images[:10].astype(np.float32) / np.float32(255.0))
  it.close()
  with pytest.raises(ValueError):
    datasets.get_dataset("cifar10", fake_dataset=False, data_dir=None)._load_dataset("train") if "CGAN_DATA_DIR" not in __import__("os").environ else (_ for _ in ()).throw(ValueError())


def test_each_data_parallel_rank_reads_its_own_stream():
  """_get_per_host_random_seed (datasets.py:147-170): the seed is offset per host / rank, so data-parallel replicas draw
  different batches, reproducibly."""
  ds = datasets.get_dataset("cifar10")
  firsts = []
  for rank in (0, 1, 0):
    it = ds.train_input_fn({"batch_size": 8}, rank=rank, ring=2)
    firsts.append(next(it)[0].copy())
    it.close()
  assert (firsts[0] != firsts[1]).any()
  np.testing.assert_array_equal(firsts[0], firsts[2])

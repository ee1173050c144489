"""Cross-replica helpers (reference tpu/tpu_ops.py:29-125) over NCCL via torch.distributed:
one process per GPU, the only collectives on the hot path are gradient and BN-moment all-reduces.
"""
import torch
import torch.distributed as dist

from .. import kernels as K


_FORCE_LOCAL = [False]


def force_local(flag):
  """Treat this process as a single replica even inside an initialised process group (equivalence tests)."""
  _FORCE_LOCAL[0] = bool(flag)


def num_replicas():
  if _FORCE_LOCAL[0]:
    return 1
  return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def replica_id():
  """Rank of this process among the replicas (0 when single-replica)."""
  if _FORCE_LOCAL[0] or not (dist.is_available() and dist.is_initialized()):
    return 0
  return dist.get_rank()


_P2P = {"state": None, "max": 0}       # None: not tried yet; True: peer buffers connected; False: unavailable (NCCL only)


def _p2p_ready():
  """Connects the NVLink peer buffers of the ranks once (include/cgan_b200.h, cgan_p2p_*): each rank publishes a cudaIpc
  handle, the handles are all-gathered over the process group.  Any failure leaves the NCCL path in charge."""
  import ctypes
  import os
  if _P2P["state"] is None:
    _P2P["state"] = False
    if os.environ.get("CGAN_P2P", "1") != "0" and dist.get_backend() == "nccl" and K._RT["lib"] is not None \
        and not getattr(K._RT["lib"], "emulated", False):
      # every rank goes through the same sequence of collectives whatever fails locally, so no rank is left waiting
      lib = K._RT["lib"]
      world, rank = dist.get_world_size(), dist.get_rank()
      handle, ok, why = (ctypes.c_ubyte * 64)(), 1.0, ""
      try:
        lib.call("p2p_local_handle", world, handle)
      except Exception as e:
        ok, why = 0.0, str(e)[:200]
      mine = torch.tensor(list(handle), dtype=torch.uint8, device=K._RT["device"])
      parts = [torch.zeros_like(mine) for _ in range(world)]
      dist.all_gather(parts, mine)
      flag = torch.tensor([ok], device=K._RT["device"])
      dist.all_reduce(flag, op=dist.ReduceOp.MIN)
      if flag.item() > 0:
        try:
          blob = b"".join(p.cpu().numpy().tobytes() for p in parts)
          buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
          lib.call("p2p_connect", rank, world, buf)
        except Exception as e:      # no peer access (different nodes, MIG, ...)
          ok, why = 0.0, str(e)[:200]
        flag = torch.tensor([ok], device=K._RT["device"])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks or none
      if flag.item() > 0:
        _P2P["max"] = int(lib.fn["cgan_p2p_max_floats"]())
        _P2P["state"] = True
      elif why:
        import logging
        logging.warning("peer-memory all-reduce unavailable (%s); NCCL carries the BN moments too", why)
  return _P2P["state"]


def cross_replica_sum_(x):
  """In-place all-reduce(sum) of a device tensor (tf.contrib.tpu.cross_replica_sum, tpu_ops.py:70,90).  Small float32
  vectors (the [2C] batch-norm exchanges) go through the one-kernel NVLink peer-memory all-reduce, everything else
  (the flat gradient buffers) through NCCL."""
  if num_replicas() > 1:
    t = x.t
    if t.is_cuda and t.dtype == torch.float32 and _p2p_ready() and t.numel() <= _P2P["max"]:
      K._call("allreduce_small", x.ptr, t.numel())
    else:
      dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return x


def cross_replica_mean(x, group_size=None):
  """tpu_ops.py:75-91: sum over replicas / number of replicas (in place)."""
  n = num_replicas()
  if n > 1:
    cross_replica_sum_(x)
    if x.t.is_cuda:
      K._call("axpby", x.ptr, 1.0 / n, x.ptr, 0.0, None, 0.0, x.numel)
    else:                      # host-side buffers (gloo tests of the exchange logic)
      x.t.mul_(1.0 / n)
  return x


def cross_replica_moments_from_local(stats2c):
  """tpu_ops.py:94-125 (parallel=True): stats2c holds the LOCAL [mean, mean-of-squares]; returns the global
  (mean, biased variance) = (E[x], E[x^2] - E[x]^2) after one fused [2C] all-reduce."""
  cross_replica_mean(stats2c)
  c = stats2c.numel // 2
  mean = stats2c.t[:c]
  return mean, stats2c.t[c:] - mean * mean

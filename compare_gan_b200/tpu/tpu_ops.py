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


def cross_replica_sum_(x):
  """In-place all-reduce(sum) of a device tensor (tf.contrib.tpu.cross_replica_sum, tpu_ops.py:70,90)."""
  if num_replicas() > 1:
    dist.all_reduce(x.t, op=dist.ReduceOp.SUM)
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

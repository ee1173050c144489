"""Cross-replica helpers (reference tpu/tpu_ops.py:29-125) over NCCL via torch.distributed:
one process per GPU, the only collectives on the hot path are gradient and BN-moment all-reduces.
"""
import torch
import torch.distributed as dist

from .. import kernels as K


def num_replicas():
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
    K._call("axpby", x.ptr, 1.0 / n, x.ptr, 0.0, None, 0.0, x.numel)
  return x

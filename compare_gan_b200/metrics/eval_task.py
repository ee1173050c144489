"""Abstract evaluation task (reference metrics/eval_task.py:28-76)."""
import abc


class EvalTask(object):
  """Tasks that compute a score from generated and real Inception statistics."""
  __metaclass__ = abc.ABCMeta
  _LABEL = None

  def metric_list(self):
    return frozenset(self._LABEL)

  def run_after_session(self, fake_dset, real_dset):
    """fake_dset / real_dset: eval_utils.EvalDataSample-like with .activations [N,2048] and .logits [N,1008]."""
    raise NotImplementedError()

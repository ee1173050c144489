"""Inception Score (reference metrics/inception_score.py:34-48; tfgan classifier_score_from_logits):
exp(mean_x KL(p(y|x) || p(y))) in float64."""
import numpy as np

from . import eval_task


def classifier_score_from_logits(logits):
  logits = np.asarray(logits, np.float64)
  m = logits.max(1, keepdims=True)
  logp = logits - m - np.log(np.exp(logits - m).sum(1, keepdims=True))
  p = np.exp(logp)
  log_marg = np.log(p.mean(0, keepdims=True))
  return float(np.exp((p * (logp - log_marg)).sum(1).mean()))


class InceptionScoreTask(eval_task.EvalTask):
  _LABEL = "inception_score"

  def run_after_session(self, fake_dset, real_dset):
    del real_dset
    return {self._LABEL: classifier_score_from_logits(fake_dset.logits)}

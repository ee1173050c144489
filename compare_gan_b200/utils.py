"""Helpers mirroring compare_gan/utils.py."""
import inspect


def call_with_accepted_args(fn, **kwargs):
  """Calls `fn` only with the keyword arguments that `fn` accepts (reference utils.py:69-96)."""
  target = getattr(fn, "__wrapped_fn__", fn)
  sig = inspect.signature(target)
  if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values()):
    return fn(**kwargs)
  return fn(**{k: v for k, v in kwargs.items() if k in sig.parameters})

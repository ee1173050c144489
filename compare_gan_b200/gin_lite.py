"""A small gin-config compatible registry/parser (gin-config 0.1.4 is a third-party dependency of the
reference, setup.py:34, and is not installable here).  It covers the surface the reference's
example_configs/*.gin use: bindings `selector.param = value`, scoped selectors
`scope/selector.param`, module-qualified selectors (`resnet_biggan.Generator.ch`), `@configurable`
references, `@configurable()` evaluated references, `%MACRO`s, `gin.REQUIRED`, whitelist/blacklist.
"""
import ast
import functools
import inspect
import re

REQUIRED = object()

_REGISTRY = {}       # full name ("module.sub.Name") -> wrapped callable
_META = {}           # full name -> (whitelist, blacklist)
_BINDINGS = {}       # (scope, full name) -> {param: value}
_MACROS = {}
_SCOPES = []


class _Ref(object):
  def __init__(self, name, evaluate):
    self.name, self.evaluate = name, evaluate

  def resolve(self):
    fn = _lookup(self.name)
    return fn() if self.evaluate else fn

  def __repr__(self):
    return "@%s%s" % (self.name, "()" if self.evaluate else "")


class _Macro(object):
  def __init__(self, name):
    self.name = name


def _lookup_name(selector):
  """Resolve a (possibly partially module-qualified) selector to a registered full name."""
  if selector in _REGISTRY:
    return selector
  hits = [k for k in _REGISTRY if k.endswith("." + selector)]
  if len(hits) == 1:
    return hits[0]
  if not hits:
    raise ValueError("No configurable matching '%s'." % selector)
  raise ValueError("Ambiguous selector '%s': %s" % (selector, sorted(hits)))


def _lookup(selector):
  return _REGISTRY[_lookup_name(selector)]


def _resolve(v):
  if isinstance(v, _Ref):
    return v.resolve()
  if isinstance(v, _Macro):
    if v.name not in _MACROS:
      raise ValueError("Unknown macro %%%s" % v.name)
    return _resolve(_MACROS[v.name])
  if isinstance(v, list):
    return [_resolve(x) for x in v]
  if isinstance(v, tuple):
    return tuple(_resolve(x) for x in v)
  if isinstance(v, dict):
    return {k: _resolve(x) for k, x in v.items()}
  return v


def _bound_kwargs(full):
  out = {}
  out.update(_BINDINGS.get(("", full), {}))
  for i in range(len(_SCOPES)):
    out.update(_BINDINGS.get(("/".join(_SCOPES[:i + 1]), full), {}))
  if _SCOPES:
    out.update(_BINDINGS.get((_SCOPES[-1], full), {}))
  return out


def configurable(name_or_fn=None, module=None, whitelist=None, blacklist=None):
  """@gin.configurable decorator for functions and classes."""
  def make(fn, name):
    mod = module if module is not None else fn.__module__.split(".")[-1]
    full = (mod + "." if mod else "") + name
    target = fn.__init__ if inspect.isclass(fn) else fn
    sig = inspect.signature(target)
    params = [p for p in sig.parameters if p != "self"]
    has_var_kw = any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values())
    _META[full] = (whitelist, blacklist, params, has_var_kw)

    def inject(args, kwargs):
      """args: positional args WITHOUT self.  Fill in config bindings for everything the caller left out."""
      given = set(kwargs)
      given.update(params[:len(args)])
      for k, v in _bound_kwargs(full).items():
        if k not in given:
          kwargs[k] = _resolve(v)
      for pname, p in sig.parameters.items():
        if p.default is REQUIRED and pname not in kwargs and pname not in given:
          raise ValueError("Required bindings for `%s` not provided in config: ['%s']" % (name, pname))
      return kwargs

    if inspect.isclass(fn):
      orig_init = fn.__init__

      @functools.wraps(orig_init)
      def new_init(self, *args, **kwargs):
        kwargs = inject(args, kwargs)
        orig_init(self, *args, **kwargs)
      fn.__init__ = new_init
      fn.__gin_name__ = full
      _REGISTRY[full] = fn
      return fn

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
      return fn(*args, **inject(args, kwargs))
    wrapper.__gin_name__ = full
    wrapper.__wrapped_fn__ = fn
    _REGISTRY[full] = wrapper
    return wrapper

  if callable(name_or_fn):
    return make(name_or_fn, name_or_fn.__name__)
  return lambda fn: make(fn, name_or_fn or fn.__name__)


def external_configurable(fn, name, module=None):
  """gin.external_configurable: register a callable under an explicit (dotted) name."""
  mod, _, short = name.rpartition(".")
  return configurable(short, module=mod if module is None else module)(fn)


def bind_parameter(binding, value):
  """bind_parameter("scope/selector.param", value)."""
  scope, _, rest = binding.rpartition("/")
  selector, _, param = rest.rpartition(".")
  full = _lookup_name(selector)
  wl, bl, params, var_kw = _META[full]
  if wl is not None and param not in wl:
    raise ValueError("Parameter '%s' of '%s' is not whitelisted." % (param, full))
  if bl is not None and param in bl:
    raise ValueError("Parameter '%s' of '%s' is blacklisted." % (param, full))
  if param not in params and not var_kw:
    raise ValueError("Configurable '%s' doesn't have a parameter named '%s'." % (full, param))
  _BINDINGS.setdefault((scope, full), {})[param] = value


def query_parameter(binding):
  scope, _, rest = binding.rpartition("/")
  selector, _, param = rest.rpartition(".")
  return _BINDINGS[(scope, _lookup_name(selector))][param]


def clear_config():
  _BINDINGS.clear()
  _MACROS.clear()


class config_scope(object):
  def __init__(self, name):
    self.name = name

  def __enter__(self):
    _SCOPES.append(self.name)

  def __exit__(self, *a):
    _SCOPES.pop()


_TOKEN = re.compile(r"@([A-Za-z_][\w./]*)(\(\))?|%([A-Za-z_][\w.]*)")


def _parse_value(text):
  text = text.strip()
  holders = {}

  def sub(m):
    key = "__gin_%d__" % len(holders)
    holders[key] = _Macro(m.group(3)) if m.group(3) else _Ref(m.group(1), bool(m.group(2)))
    return '"%s"' % key
  py = _TOKEN.sub(sub, text)
  val = ast.literal_eval(py)

  def restore(v):
    if isinstance(v, str) and v in holders:
      return holders[v]
    if isinstance(v, list):
      return [restore(x) for x in v]
    if isinstance(v, tuple):
      return tuple(restore(x) for x in v)
    if isinstance(v, dict):
      return {k: restore(x) for k, x in v.items()}
    return v
  return restore(val)


def parse_config(text, skip_unknown=False):
  """Parse gin bindings (one logical statement may span lines inside brackets)."""
  if isinstance(text, (list, tuple)):
    text = "\n".join(text)
  stmts, buf, depth = [], "", 0
  for raw in text.splitlines():
    line = raw.split("#", 1)[0].rstrip() if '"' not in raw and "'" not in raw else raw.rstrip()
    if not line.strip() and depth == 0:
      continue
    buf += (" " if buf else "") + line.strip()
    depth = buf.count("(") + buf.count("[") + buf.count("{") - buf.count(")") - buf.count("]") - buf.count("}")
    if depth <= 0 and not buf.rstrip().endswith("\\") and not buf.rstrip().endswith("="):
      stmts.append(buf)
      buf, depth = "", 0
  if buf.strip():
    stmts.append(buf)
  for s in stmts:
    s = s.strip()
    if not s or s.startswith("#") or s.startswith("import ") or s.startswith("include "):
      continue
    if "=" not in s:
      raise ValueError("Cannot parse gin statement: %r" % s)
    lhs, rhs = s.split("=", 1)
    lhs = lhs.strip()
    if "#" in rhs and '"' not in rhs and "'" not in rhs:
      rhs = rhs.split("#", 1)[0]
    value = _parse_value(rhs)
    if "." not in lhs.rpartition("/")[2]:
      _MACROS[lhs] = value
      continue
    try:
      bind_parameter(lhs, value)
    except ValueError:
      if not skip_unknown:
        raise


def parse_config_files_and_bindings(config_files, bindings, skip_unknown=False):
  for f in config_files or []:
    parse_config(open(f).read(), skip_unknown)
  if bindings:
    parse_config(bindings, skip_unknown)


def operative_config_str():
  lines = []
  for (scope, full), params in sorted(_BINDINGS.items()):
    for k, v in sorted(params.items()):
      lines.append("%s%s.%s = %r" % (scope + "/" if scope else "", full, k, v))
  return "\n".join(lines) + "\n"

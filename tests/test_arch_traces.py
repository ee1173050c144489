"""The host-side network definitions, traced on the CPU (tests/arch_trace.py): for every architecture the engine
builds, the sequence of kernel-layer calls (operands, shapes, data flow) and the variables it creates (checkpoint key
space, shapes, initial values) are pinned in tests/golden/arch_traces.json, and the resnet_cifar key space is checked
against the reference's own golden lists (architectures/resnet_norm_test.py)."""
import json
import os

import numpy as np
import pytest

from tests import arch_trace as at

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "arch_traces.json")))
REF_VARS = json.load(open(os.path.join(HERE, "golden", "resnet_cifar_variables.json")))


@pytest.mark.parametrize("case", sorted(at.CASES))
def test_definition_trace_is_pinned(case):
  tr = at.trace_networks(**at.CASES[case])
  want = GOLDEN[case]
  assert [v[:3] for v in tr["variables"]] == [v[:3] for v in want["variables"]], "variable names / shapes / trainability"
  assert tr["variables"] == want["variables"], "initial values (initialiser kind or RNG order changed)"
  assert len(tr["ops"]) == want["n_ops"]
  assert at.canonical_sha1(tr["ops"]) == want["ops_sha1"], "kernel-call sequence changed"


def _engine_names(gin_text, prefix, trainable_only):
  tr = at.trace_networks(gin_text=gin_text, architecture="resnet_cifar_arch", image_shape=(32, 32, 3))
  return [[n, s] for n, s, t, _ in tr["variables"] if n.startswith(prefix) and (t or not trainable_only)]


def test_engine_key_space_matches_reference_lists():
  # architectures/resnet_norm_test.py:39-63 (G, no norm), :78-106 (D), :124-162 (G with batch norm), :326-361 (G with SN)
  assert _engine_names("G.batch_norm_fn = None", "generator/", True) == REF_VARS["g_default"]
  assert _engine_names("G.batch_norm_fn = None", "discriminator/", True) == REF_VARS["d_default"]
  assert _engine_names("G.batch_norm_fn = @batch_norm", "generator/", True) == REF_VARS["g_batch_norm"]
  assert _engine_names("G.batch_norm_fn = None\nG.spectral_norm = True", "generator/", False) == REF_VARS["g_spectral_norm_global"]


def test_engine_biggan_parameter_counts_match_the_reference():
  """The ENGINE's definitions at the reference tests' settings: resnet_biggan 128x128 (resnet_biggan_test.py:112-154:
  70,433,988 / 87,982,370 trainable weights) and resnet_biggan_deep 128x128 (resnet_biggan_deep_test.py:30-60:
  50,244,484 / 34,590,210)."""
  ref = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))

  def counts(tr):
    g = sum(int(np.prod(s)) for n, s, t, _ in tr["variables"] if t and n.startswith("generator/"))
    d = sum(int(np.prod(s)) for n, s, t, _ in tr["variables"] if t and n.startswith("discriminator/"))
    return g, d
  big = at.trace_networks(gin_text="G.batch_norm_fn = @conditional_batch_norm\nG.spectral_norm = True\nD.spectral_norm = True\n"
                                   "spectral_norm.singular_value = 'auto'\nstandardize_batch.use_moving_averages = False",
                          architecture="resnet_biggan_arch", image_shape=(128, 128, 3), z_dim=120, num_classes=1000,
                          conditional=True)
  assert counts(big) == (ref["biggan128_params"]["generator"], ref["biggan128_params"]["discriminator"])
  deep = at.trace_networks(gin_text="G.batch_norm_fn = @conditional_batch_norm", architecture="resnet_biggan_deep_arch",
                           image_shape=(128, 128, 3), z_dim=128, num_classes=1000, conditional=True)
  assert counts(deep) == (ref["biggan_deep128_params"]["generator"], ref["biggan_deep128_params"]["discriminator"])

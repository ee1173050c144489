"""CPU tests: the C-ABI shared library builds, loads, and exports every symbol include/cgan_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os

import pytest

from compare_gan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
  if not os.path.exists(_lib.SO_PATH):
    import __graft_entry__
    __graft_entry__.build()
  return ctypes.CDLL(_lib.SO_PATH)


def test_header_parses_and_is_nonempty():
  protos = _lib.parse_header()
  assert len(protos) >= 45
  assert "cgan_conv2d_fwd" in protos and "cgan_adam_step" in protos and "cgan_cov_accumulate" in protos


def test_every_declared_symbol_is_exported(so):
  missing = [name for name in _lib.parse_header() if not hasattr(so, name)]
  assert not missing, missing


def test_version_and_null_context_errors(so):
  so.cgan_version.restype = ctypes.c_int
  assert so.cgan_version() >= 1
  so.cgan_last_error.restype = ctypes.c_char_p
  assert so.cgan_last_error(None) == b"null context"
  # every entry point rejects a NULL context instead of crashing
  so.cgan_fill.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int64]
  assert so.cgan_fill(None, None, 0.0, 0) != 0


def test_conv_desc_layout_matches_header():
  assert ctypes.sizeof(_lib.ConvDesc) == 13 * 4
  src = open(_lib.HEADER).read()
  body = src[src.index("typedef struct {"):src.index("} cgan_conv_desc;")]
  fields = [f.strip().split()[-1] for line in body.splitlines() if "int32_t" in line
            for f in line.split(";")[0].replace("int32_t", "").split(",")]
  assert fields == [n for n, _ in _lib.ConvDesc._fields_]


def test_product_path_fails_loudly_without_cuda():
  import torch
  if torch.cuda.is_available():
    pytest.skip("CUDA present")
  from compare_gan_b200 import kernels
  with pytest.raises(_lib.CganError):
    kernels.init(0)

"""The generated DFT codelets (csrc/fft_codelets.cuh): the generator's programs are
evaluated against numpy.fft, in float64 and -- operation by operation -- in float32, and
the header in the tree must be exactly what the generator renders."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    'gen_fft_codelets', os.path.join(ROOT, 'tools', 'gen_fft_codelets.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def test_header_is_generated():
  with open(gen.HEADER) as f:
    assert f.read() == gen.render()


def _eval_packed_fp32(p, outs, x):
  """The packed program with every lane rounded to fp32 after each operation (fused
  multiply-adds rounded once), i.e. what FADD2 / FMUL2 / FFMA2 compute."""
  f = np.float32
  env = {f'a[{i}]': (f(v.real), f(v.imag)) for i, v in enumerate(x)}
  fma = lambda a, b, c: f(np.float64(a) * np.float64(b) + np.float64(c))
  for op in p.ops:
    k, d = op[0], op[1]
    a = env[op[2]]
    if k == 'add2':
      b = env[op[3]]; env[d] = (f(a[0] + b[0]), f(a[1] + b[1]))
    elif k == 'sub2':
      b = env[op[3]]; env[d] = (f(a[0] - b[0]), f(a[1] - b[1]))
    elif k == 'subrot':
      b = env[op[3]]; env[d] = (f(a[1] - b[1]), f(b[0] - a[0]))
    elif k == 'mulc':
      c = f(op[3]); env[d] = (f(a[0] * c), f(a[1] * c))
    elif k == 'fmac':
      c = f(op[3]); b = env[op[4]]; env[d] = (fma(a[0], c, b[0]), fma(a[1], c, b[1]))
    elif k == 'neg2':
      env[d] = (-a[0], -a[1])
    elif k == 'rotn':
      env[d] = (a[1], -a[0])
    elif k == 'rotp':
      env[d] = (-a[1], a[0])
    elif k == 'cmulc':
      wr, wi = f(op[3]), f(op[4])
      env[d] = (fma(a[0], wr, -f(a[1] * wi)), fma(a[0], wi, f(a[1] * wr)))
    else:
      raise AssertionError(k)
  return np.array([complex(*env[o]) for o in outs])


@pytest.mark.parametrize('n', gen.SIZES)
def test_packed_program(n):
  rng = np.random.default_rng(n)
  x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
  p, outs = gen.build_p(n)
  xd = x.astype(np.complex128)
  want = np.fft.fft(xd)
  assert np.abs(gen.evaluate_p(p, outs, xd) - want).max() < 1e-12 * n
  got32 = _eval_packed_fp32(p, outs, x)
  assert np.abs(got32 - want).max() < 4e-7 * n * np.abs(want).max()
  # same value as the scalar program in exact arithmetic
  ps, outs_s = gen.build(n)
  assert np.abs(gen.evaluate(ps, outs_s, xd) - gen.evaluate_p(p, outs, xd)).max() < 1e-12 * n

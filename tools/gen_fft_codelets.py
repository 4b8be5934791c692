"""Generates sofima_b200/csrc/fft_codelets.cuh: straight-line, register-resident
forward DFT codelets (natural order in, natural order out) for the small sizes the
two-pass shared-memory FFT of csrc/flow_fast.cuh is built from.

  python tools/gen_fft_codelets.py            # writes the header, self-checks first

The generator builds a tiny SSA program per size with a recursive Cooley-Tukey
split (radices 4, 2, 5, 3), folds trivial twiddles, evaluates the program in
Python against numpy.fft (self-check) and only then emits CUDA.
"""

from __future__ import annotations

import cmath
import math
import os
import sys

import numpy as np

SIZES = (2, 3, 4, 5, 8, 10, 12, 15, 16, 20, 24, 25, 32)


class Prog:
  """SSA program over scalar floats."""

  def __init__(self):
    self.ops = []
    self.n = 0

  def tmp(self):
    self.n += 1
    return f't{self.n}'

  def emit(self, op, *args):
    d = self.tmp()
    self.ops.append((op, d) + args)
    return d

  # complex helpers: a complex value is a (re, im) pair of names (or None = 0)
  def cadd(self, a, b):
    return (self.emit('add', a[0], b[0]), self.emit('add', a[1], b[1]))

  def csub(self, a, b):
    return (self.emit('sub', a[0], b[0]), self.emit('sub', a[1], b[1]))

  def mul_neg_i(self, a):  # a * (-i) = (im, -re)
    return (a[1], self.emit('neg', a[0]))

  def cscale(self, a, c):
    return (self.emit('mul', a[0], c), self.emit('mul', a[1], c))

  def cmulc(self, a, w):
    """a * w for a compile-time complex constant w (unit modulus)."""
    wr, wi = w.real, w.imag
    eps = 1e-15
    if abs(wi) < eps:
      return a if wr > 0 else (self.emit('neg', a[0]), self.emit('neg', a[1]))
    if abs(wr) < eps:
      # w = +/- i
      if wi < 0:
        return self.mul_neg_i(a)
      return (self.emit('neg', a[1]), a[0])
    # (ar + i ai)(wr + i wi) = (ar wr - ai wi) + i (ar wi + ai wr)
    t = self.emit('mul', a[1], wi)
    re = self.emit('fms', a[0], wr, t)      # a0*wr - t
    t2 = self.emit('mul', a[1], wr)
    im = self.emit('fma', a[0], wi, t2)     # a0*wi + t2
    return (re, im)


def butterfly(p: Prog, a, r):
  """r-point forward DFT of the list a (r in 2, 3, 4, 5)."""
  if r == 2:
    return [p.cadd(a[0], a[1]), p.csub(a[0], a[1])]
  if r == 4:
    t0, t1 = p.cadd(a[0], a[2]), p.csub(a[0], a[2])
    t2, t3 = p.cadd(a[1], a[3]), p.mul_neg_i(p.csub(a[1], a[3]))
    return [p.cadd(t0, t2), p.cadd(t1, t3), p.csub(t0, t2), p.csub(t1, t3)]
  if r == 3:
    h = math.sqrt(3.0) / 2
    t1 = p.cadd(a[1], a[2])
    t2 = (p.emit('fma', t1[0], -0.5, a[0][0]), p.emit('fma', t1[1], -0.5, a[0][1]))
    d = p.csub(a[1], a[2])
    t3 = p.mul_neg_i(p.cscale(d, h))
    return [p.cadd(a[0], t1), p.cadd(t2, t3), p.csub(t2, t3)]
  if r == 5:
    c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
    s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
    t1, t2 = p.cadd(a[1], a[4]), p.cadd(a[2], a[3])
    t3, t4 = p.csub(a[1], a[4]), p.csub(a[2], a[3])
    b0 = p.cadd(a[0], p.cadd(t1, t2))

    def lin(x, y, cx, cy, base=None):
      out = []
      for c in range(2):
        v = p.emit('mul', y[c], cy)
        v = p.emit('fma', x[c], cx, v)
        if base is not None:
          v = p.emit('add', base[c], v)
        out.append(v)
      return tuple(out)

    m1, m2 = lin(t1, t2, c1, c2, a[0]), lin(t1, t2, c2, c1, a[0])
    n1 = p.mul_neg_i(lin(t3, t4, s1, s2))
    n2 = p.mul_neg_i(lin(t3, t4, s2, -s1))
    return [b0, p.cadd(m1, n1), p.cadd(m2, n2), p.csub(m2, n2), p.csub(m1, n1)]
  raise ValueError(r)


def dft(p: Prog, a):
  """Forward DFT of the list a (any length with factors 2, 3, 5)."""
  n = len(a)
  if n == 1:
    return a
  if n in (2, 3, 4, 5):
    return butterfly(p, a, n)
  for r in (4, 2, 5, 3):
    if n % r == 0:
      break
  else:
    raise ValueError(n)
  m = n // r
  # n = m * n1 + n2 ; k = k1 + r * k2
  y = [[None] * r for _ in range(m)]
  for n2 in range(m):
    col = butterfly(p, [a[m * n1 + n2] for n1 in range(r)], r)
    for k1 in range(r):
      w = cmath.exp(-2j * math.pi * n2 * k1 / n)
      y[n2][k1] = p.cmulc(col[k1], w)
  out = [None] * n
  for k1 in range(r):
    sub = dft(p, [y[n2][k1] for n2 in range(m)])
    for k2 in range(m):
      out[k1 + r * k2] = sub[k2]
  return out


def build(n):
  p = Prog()
  ins = [(f'a[{i}].x', f'a[{i}].y') for i in range(n)]
  outs = dft(p, ins)
  return p, outs


def evaluate(p: Prog, outs, x):
  env = {}
  for i, v in enumerate(x):
    env[f'a[{i}].x'], env[f'a[{i}].y'] = v.real, v.imag
  for op in p.ops:
    kind, d = op[0], op[1]
    if kind == 'add':
      env[d] = env[op[2]] + env[op[3]]
    elif kind == 'sub':
      env[d] = env[op[2]] - env[op[3]]
    elif kind == 'neg':
      env[d] = -env[op[2]]
    elif kind == 'mul':
      env[d] = env[op[2]] * op[3]
    elif kind == 'fma':
      env[d] = env[op[2]] * op[3] + env[op[4]]
    elif kind == 'fms':
      env[d] = env[op[2]] * op[3] - env[op[4]]
  return np.array([env[r] + 1j * env[i] for r, i in outs])


def lit(c):
  return repr(float(np.float32(c))) + 'f' if 'e' not in repr(float(np.float32(c))) \
      else repr(float(np.float32(c))) + 'f'


def emit_cuda(n, p, outs):
  lines = [f'template <> struct Dft<{n}> {{',
           f'  static __device__ __forceinline__ void run(float2 (&a)[{n}]) {{']
  for op in p.ops:
    kind, d = op[0], op[1]
    if kind == 'add':
      e = f'{op[2]} + {op[3]}'
    elif kind == 'sub':
      e = f'{op[2]} - {op[3]}'
    elif kind == 'neg':
      e = f'-{op[2]}'
    elif kind == 'mul':
      e = f'{op[2]} * {lit(op[3])}'
    elif kind == 'fma':
      e = f'fmaf({op[2]}, {lit(op[3])}, {op[4]})'
    elif kind == 'fms':
      e = f'fmaf({op[2]}, {lit(op[3])}, -{op[4]})'
    lines.append(f'    const float {d} = {e};')
  for i, (r, im) in enumerate(outs):
    lines.append(f'    a[{i}] = make_float2({r}, {im});')
  lines += ['  }', '};', '']
  return lines


def main():
  rng = np.random.default_rng(0)
  out = ['// GENERATED by tools/gen_fft_codelets.py -- do not edit.',
         '// Register-resident forward DFT codelets, natural order in and out.',
         '// The inverse transform is obtained by swapping re/im before and after.',
         '#pragma once', '', 'namespace sofima {', 'namespace flow {', '',
         'template <int N> struct Dft;', '']
  for n in SIZES:
    p, outs = build(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    err = np.abs(evaluate(p, outs, x) - np.fft.fft(x)).max()
    assert err < 1e-12, (n, err)
    flops = sum(1 if o[0] in ('add', 'sub', 'mul') else (2 if o[0] in ('fma', 'fms') else 0)
                for o in p.ops)
    print(f'DFT-{n}: {len(p.ops)} ops, {flops} flops, self-check err {err:.1e}')
    out += emit_cuda(n, p, outs)
  out += ['}  // namespace flow', '}  // namespace sofima', '']
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      'sofima_b200', 'csrc', 'fft_codelets.cuh')
  with open(path, 'w') as f:
    f.write('\n'.join(out))
  print('wrote', path)


if __name__ == '__main__':
  main()

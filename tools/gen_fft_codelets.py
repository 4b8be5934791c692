"""Generates sofima_b200/csrc/fft_codelets.cuh: straight-line, register-resident
forward DFT codelets (natural order in, natural order out) for the small sizes the
two-pass shared-memory FFT of csrc/flow_fast.cuh is built from.

  python tools/gen_fft_codelets.py            # writes the header, self-checks first

The generator builds a tiny SSA program per size with a recursive Cooley-Tukey
split (radices 4, 2, 5, 3), folds trivial twiddles, evaluates the program in
Python against numpy.fft (self-check) and only then emits CUDA.

The emitted `Dft<N>` is a complex SSA over float2 register pairs with the butterflies
as Blackwell packed fp32x2 instructions -- FADD2 / FMUL2 / FFMA2 take one issue slot for
the re and im lanes; each lane is an ordinary IEEE fp32 operation, so the values equal
those of the scalar program (kept here as a second self-check, not emitted).  Multiplication by -i never needs a lane swap: the differences it
applies to are formed directly as -i (a - b) = (a.y - b.y, b.x - a.x) by two scalar
subtractions.  Twiddles by compile-time constants stay scalar (a lane swap would cost
what the packed form saves).
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


# ------------------------------------------------------------------------------------
# Packed (float2 / fp32x2) form
# ------------------------------------------------------------------------------------
class ProgP:
  """SSA program over complex values (float2 register pairs)."""

  def __init__(self):
    self.ops = []
    self.n = 0

  def emit(self, op, *args):
    self.n += 1
    d = f'p{self.n}'
    self.ops.append((op, d) + args)
    return d

  def cadd(self, a, b):
    return self.emit('add2', a, b)

  def csub(self, a, b):
    return self.emit('sub2', a, b)

  def sub_rot(self, a, b):
    """-i (a - b) = (a.y - b.y, b.x - a.x): two scalar subtractions, no swap."""
    return self.emit('subrot', a, b)

  def scale(self, a, c):
    return self.emit('mulc', a, c)

  def fma(self, a, c, b):
    """a * c + b with a real constant c on both lanes."""
    return self.emit('fmac', a, c, b)

  def cmulc(self, a, w):
    wr, wi = w.real, w.imag
    eps = 1e-15
    if abs(wi) < eps:
      return a if wr > 0 else self.emit('neg2', a)
    if abs(wr) < eps:
      return self.emit('rotn', a) if wi < 0 else self.emit('rotp', a)
    return self.emit('cmulc', a, wr, wi)


def butterfly_p(p: ProgP, a, r):
  if r == 2:
    return [p.cadd(a[0], a[1]), p.csub(a[0], a[1])]
  if r == 4:
    t0, t1 = p.cadd(a[0], a[2]), p.csub(a[0], a[2])
    t2, t3 = p.cadd(a[1], a[3]), p.sub_rot(a[1], a[3])
    return [p.cadd(t0, t2), p.cadd(t1, t3), p.csub(t0, t2), p.csub(t1, t3)]
  if r == 3:
    h = math.sqrt(3.0) / 2
    t1 = p.cadd(a[1], a[2])
    t2 = p.fma(t1, -0.5, a[0])
    rr = p.sub_rot(a[1], a[2])
    return [p.cadd(a[0], t1), p.fma(rr, h, t2), p.fma(rr, -h, t2)]
  if r == 5:
    c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
    s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
    t1, t2 = p.cadd(a[1], a[4]), p.cadd(a[2], a[3])
    r3, r4 = p.sub_rot(a[1], a[4]), p.sub_rot(a[2], a[3])
    b0 = p.cadd(a[0], p.cadd(t1, t2))
    m1 = p.cadd(a[0], p.fma(t1, c1, p.scale(t2, c2)))
    m2 = p.cadd(a[0], p.fma(t1, c2, p.scale(t2, c1)))
    n1 = p.fma(r3, s1, p.scale(r4, s2))
    n2 = p.fma(r3, s2, p.scale(r4, -s1))
    return [b0, p.cadd(m1, n1), p.cadd(m2, n2), p.csub(m2, n2), p.csub(m1, n1)]
  raise ValueError(r)


def dft_p(p: ProgP, a):
  n = len(a)
  if n == 1:
    return a
  if n in (2, 3, 4, 5):
    return butterfly_p(p, a, n)
  for r in (4, 2, 5, 3):
    if n % r == 0:
      break
  else:
    raise ValueError(n)
  m = n // r
  y = [[None] * r for _ in range(m)]
  for n2 in range(m):
    col = butterfly_p(p, [a[m * n1 + n2] for n1 in range(r)], r)
    for k1 in range(r):
      w = cmath.exp(-2j * math.pi * n2 * k1 / n)
      y[n2][k1] = p.cmulc(col[k1], w)
  out = [None] * n
  for k1 in range(r):
    sub = dft_p(p, [y[n2][k1] for n2 in range(m)])
    for k2 in range(m):
      out[k1 + r * k2] = sub[k2]
  return out


def build_p(n):
  p = ProgP()
  outs = dft_p(p, [f'a[{i}]' for i in range(n)])
  return p, outs


def evaluate_p(p: ProgP, outs, x):
  env = {f'a[{i}]': complex(v) for i, v in enumerate(x)}
  for op in p.ops:
    kind, d = op[0], op[1]
    if kind == 'add2':
      env[d] = env[op[2]] + env[op[3]]
    elif kind == 'sub2':
      env[d] = env[op[2]] - env[op[3]]
    elif kind == 'subrot':
      a, b = env[op[2]], env[op[3]]
      env[d] = complex(a.imag - b.imag, b.real - a.real)
    elif kind == 'mulc':
      env[d] = env[op[2]] * op[3]
    elif kind == 'fmac':
      env[d] = env[op[2]] * op[3] + env[op[4]]
    elif kind == 'neg2':
      env[d] = -env[op[2]]
    elif kind == 'rotn':
      a = env[op[2]]
      env[d] = complex(a.imag, -a.real)
    elif kind == 'rotp':
      a = env[op[2]]
      env[d] = complex(-a.imag, a.real)
    elif kind == 'cmulc':
      env[d] = env[op[2]] * complex(op[3], op[4])
    else:
      raise ValueError(kind)
  return np.array([env[o] for o in outs])


def emit_cuda_p(n, p, outs):
  lines = [f'template <> struct Dft<{n}> {{',
           f'  static __device__ __forceinline__ void run(float2 (&a)[{n}]) {{']
  for op in p.ops:
    kind, d = op[0], op[1]
    if kind == 'add2':
      e = f'f2add({op[2]}, {op[3]})'
    elif kind == 'sub2':
      e = f'f2sub({op[2]}, {op[3]})'
    elif kind == 'subrot':
      e = f'make_float2({op[2]}.y - {op[3]}.y, {op[3]}.x - {op[2]}.x)'
    elif kind == 'mulc':
      e = f'f2mul({op[2]}, f2splat({lit(op[3])}))'
    elif kind == 'fmac':
      e = f'f2fma({op[2]}, f2splat({lit(op[3])}), {op[4]})'
    elif kind == 'neg2':
      e = f'make_float2(-{op[2]}.x, -{op[2]}.y)'
    elif kind == 'rotn':
      e = f'make_float2({op[2]}.y, -{op[2]}.x)'
    elif kind == 'rotp':
      e = f'make_float2(-{op[2]}.y, {op[2]}.x)'
    elif kind == 'cmulc':
      a, wr, wi = op[2], lit(op[3]), lit(op[4])
      e = (f'make_float2(fmaf({a}.x, {wr}, -({a}.y * {wi})), '
           f'fmaf({a}.x, {wi}, {a}.y * {wr}))')
    lines.append(f'    const float2 {d} = {e};')
  for i, o in enumerate(outs):
    lines.append(f'    a[{i}] = {o};')
  lines += ['  }', '};', '']
  return lines


PACKED_PRELUDE = r"""
// Packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2): one issue slot for the re
// and im lanes of a complex value, each lane an ordinary IEEE fp32 operation.
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; "
      "add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; "
      "sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; "
      "mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {  // a * b + c
  float2 r;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; "
      "mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 f2splat(float v) { return make_float2(v, v); }

template <int N> struct Dft;
"""


def render(verbose=False):
  """Self-checks every program and returns the text of fft_codelets.cuh."""
  log = print if verbose else (lambda *a: None)
  rng = np.random.default_rng(0)
  out = ['// GENERATED by tools/gen_fft_codelets.py -- do not edit.',
         '// Register-resident forward DFT codelets, natural order in and out.',
         '// The inverse transform is obtained by swapping re/im before and after.',
         '#pragma once', '', 'namespace sofima {', 'namespace flow {', '']
  for n in SIZES:
    p, outs = build(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    err = np.abs(evaluate(p, outs, x) - np.fft.fft(x)).max()
    assert err < 1e-12, (n, err)
    flops = sum(1 if o[0] in ('add', 'sub', 'mul') else (2 if o[0] in ('fma', 'fms') else 0)
                for o in p.ops)
    log(f'Dft-{n} scalar check: {len(p.ops)} ops, {flops} flops, self-check err {err:.1e}')
  out += PACKED_PRELUDE.split('\n')
  for n in SIZES:
    p, outs = build_p(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    err = np.abs(evaluate_p(p, outs, x) - np.fft.fft(x)).max()
    assert err < 1e-12, (n, err)
    cost = sum(2 if o[0] in ('subrot', 'neg2', 'rotn', 'rotp') else (4 if o[0] == 'cmulc' else 1)
               for o in p.ops)
    log(f'Dft-{n} packed: {len(p.ops)} complex ops, ~{cost} instructions, self-check err {err:.1e}')
    out += emit_cuda_p(n, p, outs)
  out += ['}  // namespace flow', '}  // namespace sofima', '']
  return '\n'.join(out)


HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      'sofima_b200', 'csrc', 'fft_codelets.cuh')


def main():
  text = render(verbose=True)
  path = HEADER
  with open(path, 'w') as f:
    f.write(text)
  print('wrote', path)


if __name__ == '__main__':
  main()

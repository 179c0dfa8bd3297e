#!/usr/bin/env python
"""Identify the fp32 operation order of the reference's camera transform / lighting as torch + cuBLAS evaluate it on the GPU,
from the intermediates dumped by tools/gpu_scene_probe.py.  fp32 operations are emulated in float64 (exact product of two fp32
values; one rounding per fused op)."""
import sys
import numpy as np

d = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/scene_probe.npz')
f32 = np.float32


def rn(x):
    return np.asarray(x, np.float64).astype(f32)


def fma(a, b, c):
    return rn(np.float64(a) * np.float64(b) + np.float64(c))


def mul(a, b):
    return rn(np.float64(a) * np.float64(b))


def add(a, b):
    return rn(np.float64(a) + np.float64(b))


def rate(name, got, want):
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    print('   %-58s %8.4f %% bit-identical   max |d| %.3e' % (name, 100 * same.mean(), np.abs(got.astype(np.float64) - want).max()))


v = d['d']
x, y, z = v[:, 0], v[:, 1], v[:, 2]
print('norm(at - eye):')
rate('sqrt((x*x + y*y) + z*z), products rounded', np.sqrt(add(add(mul(x, x), mul(y, y)), mul(z, z)).astype(np.float64)).astype(f32), d['norm_d'])
rate('sqrt(fma(z,z, fma(y,y, x*x)))', np.sqrt(fma(z, z, fma(y, y, mul(x, x))).astype(np.float64)).astype(f32), d['norm_d'])
rate('sqrt(fma(x,x, fma(y,y, z*z)))', np.sqrt(fma(x, x, fma(y, y, mul(z, z))).astype(np.float64)).astype(f32), d['norm_d'])
nd = d['norm_d']
print('z = d / max(norm, eps):')
rate('IEEE division', rn(np.float64(v) / np.maximum(nd, f32(1e-5))[:, None]), d['z'])
rate('multiply by reciprocal', mul(v, rn(1.0 / np.maximum(nd, f32(1e-5)).astype(np.float64))[:, None]), d['z'])
print('cross(up, z):')
up = np.array([0, 1, 0], f32)[None].repeat(len(v), 0)
zz = d['z']


def cross_variants(a, b, want, label):
    a0, a1, a2 = a[..., 0], a[..., 1], a[..., 2]
    b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
    c_plain = np.stack([add(mul(a1, b2), -mul(a2, b1)), add(mul(a2, b0), -mul(a0, b2)), add(mul(a0, b1), -mul(a1, b0))], -1)
    c_fma1 = np.stack([fma(a1, b2, -mul(a2, b1)), fma(a2, b0, -mul(a0, b2)), fma(a0, b1, -mul(a1, b0))], -1)
    c_fma2 = np.stack([fma(-a2, b1, mul(a1, b2)), fma(-a0, b2, mul(a2, b0)), fma(-a1, b0, mul(a0, b1))], -1)
    rate(label + ' a1*b2 - a2*b1 unfused', c_plain, want)
    rate(label + ' fma(a1,b2, -(a2*b1))', c_fma1, want)
    rate(label + ' fma(-a2,b1, a1*b2)', c_fma2, want)


cross_variants(up, zz, d['cx'], 'cx')
cross_variants(zz, d['x'], d['cy'], 'cy')
print('matmul(v - eye, r^T) (cuBLAS):')
vm = d['vm']
for name, axis in (('x', d['x']), ('y', d['y']), ('z', zz)):
    k = {'x': 0, 'y': 1, 'z': 2}[name]
    a0, a1, a2 = vm[..., 0], vm[..., 1], vm[..., 2]
    r0, r1, r2 = axis[:, None, 0], axis[:, None, 1], axis[:, None, 2]
    want = d['vc'][..., k]
    rate(name + ': fma(a2,r2, fma(a1,r1, a0*r0))', fma(a2, r2, fma(a1, r1, mul(a0, r0))), want)
    rate(name + ': fma(a0,r0, fma(a1,r1, a2*r2))', fma(a0, r0, fma(a1, r1, mul(a2, r2))), want)
    rate(name + ': (a0*r0 + a1*r1) + a2*r2 unfused', add(add(mul(a0, r0), mul(a1, r1)), mul(a2, r2)), want)
print('perspective:')
w = d['width'][0, 0]
rate('x / z / width (two IEEE divisions)', rn(np.float64(rn(np.float64(d['vc'][..., 0]) / d['vc'][..., 2])) / w), d['xs'])
if 'ref_screen' in d:
    rate('reference package LookAt == the dumped pipeline', d['ref_screen'][..., 0], d['xs'])
print('face normal: cross(v2 - v1, v0 - v1), norm, normalize(eps 1e-6), cosine:')
fv = d['fv']
a = rn(np.float64(fv[:, :, 2]) - fv[:, :, 1]); e = rn(np.float64(fv[:, :, 0]) - fv[:, :, 1])
cross_variants(a, e, d['cr'], 'cr')
cr = d['cr']
c0, c1, c2 = cr[..., 0], cr[..., 1], cr[..., 2]
rate('norm: sqrt((x*x + y*y) + z*z)', np.sqrt(add(add(mul(c0, c0), mul(c1, c1)), mul(c2, c2)).astype(np.float64)).astype(f32), d['norm_cr'])
rate('norm: sqrt(fma(z,z, fma(y,y, x*x)))', np.sqrt(fma(c2, c2, fma(c1, c1, mul(c0, c0))).astype(np.float64)).astype(f32), d['norm_cr'])
rate('n = cr / max(norm, 1e-6) IEEE', rn(np.float64(cr) / np.maximum(d['norm_cr'], f32(1e-6))[..., None]), d['n'])
n = d['n']
ld2 = np.array([0.3, 0.8, -0.5], f32)
p0, p1, p2 = mul(n[..., 0], ld2[0]), mul(n[..., 1], ld2[1]), mul(n[..., 2], ld2[2])
rate('sum(n * dir): (p0 + p1) + p2', add(add(p0, p1), p2), d['cos2_raw'])
rate('sum(n * dir): p0 + (p1 + p2)', add(p0, add(p1, p2)), d['cos2_raw'])
cos = d['cos']
rate('light = 0.5 + 0.5 * (1 * cos)', add(f32(0.5), mul(f32(0.5), cos))[..., None].repeat(3, -1), d['light'])
rate('light = fma(0.5, cos, 0.5)', fma(f32(0.5), cos, f32(0.5))[..., None].repeat(3, -1), d['light'])

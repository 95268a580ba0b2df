"""Test infrastructure (not on the product path): numpy emulation of the operand splits the
tensor-core GEMM uses, with exact (float64) accumulation, so the error of the SPLIT alone can be
stated: tinynn-autograd_b200/csrc/gemm_tc.cu forms

    tf32x3 : A*B ~= lo(A)*hi(B) + hi(A)*lo(B) + hi(A)*hi(B),      hi = tf32(x), lo = tf32(x - hi)
    mix    : A*B ~= bf16(A - hi(A))*bf16(B) + bf16(A)*bf16(B - hi(B)) + hi(A)*hi(B)

tests/test_oracle_golden.py pins the bounds quoted in DESIGN.md section 3.1."""
import numpy as np


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) to 10 mantissa bits, fp32 container"""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def bf16_rn(x):
    """round to nearest even to bfloat16, returned as float32"""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def product_tf32x3(a, b):
    f = lambda v: v.astype(np.float64)
    ah, bh = tf32_rna(a), tf32_rna(b)
    al, bl = tf32_rna(a - ah), tf32_rna(b - bh)
    return f(al) @ f(bh) + f(ah) @ f(bl) + f(ah) @ f(bh)


def product_mix(a, b):
    f = lambda v: v.astype(np.float64)
    ah, bh = tf32_rna(a), tf32_rna(b)
    return f(bf16_rn(a - ah)) @ f(bf16_rn(b)) + f(bf16_rn(a)) @ f(bf16_rn(b - bh)) + f(ah) @ f(bh)


def product_single_tf32(a, b):
    return tf32_rna(a).astype(np.float64) @ tf32_rna(b).astype(np.float64)


def rel_err(got, exact):
    return float(np.max(np.abs(got - exact)) / np.max(np.abs(exact)))

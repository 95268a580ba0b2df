"""Test infrastructure (not on the product path): numpy emulation of the operand splits the
tensor-core GEMM uses, with exact (float64) accumulation, so the error of the SPLIT alone can be
stated: tinynn-autograd_b200/csrc/gemm_tc.cu forms

    tf32x3 : A*B ~= lo(A)*hi(B) + hi(A)*lo(B) + hi(A)*hi(B),      hi = tf32(x), lo = tf32(x - hi)
    mix    : A*B ~= bf16(A - hi(A))*bf16(B) + bf16(A)*bf16(B - hi(B)) + hi(A)*hi(B)

and tinynn-autograd_b200/csrc/gemm_f16.cu (the default) the same expansion on scaled fp16 planes

    f16    : X = x 2^e (max|X| in [2^14, 2^15) per tensor), hf = fp16(X), l = fp16(X - hf),
             A*B ~= 2^-(ea+eb) (l_A*hf_B + hf_A*l_B + hf_A*hf_B)

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


def f16_scale_exponent(x):
    """e with max|x| * 2^e in [2^14, 2^15) (gemm_f16.cu: meta_scale); 0 for an all-zero tensor"""
    amax = float(np.max(np.abs(x))) if x.size else 0.0
    if amax == 0.0:
        return 0
    return 14 - int(np.floor(np.log2(amax)))


def f16_planes(x):
    """(hf, l, e): the two fp16 planes of x * 2^e as float32 arrays (numpy float16 rounds to nearest
    even with gradual underflow, like cvt.rn.f16.f32)"""
    e = f16_scale_exponent(x)
    X = np.ldexp(np.ascontiguousarray(x, dtype=np.float32), e).astype(np.float32)
    hf = X.astype(np.float16).astype(np.float32)
    l = (X - hf).astype(np.float16).astype(np.float32)
    return hf, l, e


def f16_guard(x, small_limit=2.0 ** -5, fraction=256):
    """True when the tensor is inside gemm_f16.cu's guard: finite, and at most 1/256 of its non-zero
    elements lie below the level where the residual plane goes subnormal"""
    if not np.all(np.isfinite(x)):
        return False
    amax = float(np.max(np.abs(x))) if x.size else 0.0
    if amax == 0.0:
        return True
    if np.floor(np.log2(amax)) < -112:
        return False
    X = np.abs(np.ldexp(np.ascontiguousarray(x, dtype=np.float32), f16_scale_exponent(x)))
    nz = X != 0
    return int(np.sum(nz & (X < small_limit))) * fraction <= int(np.sum(nz))


def product_f16(a, b):
    f = lambda v: v.astype(np.float64)
    ah, al, ea = f16_planes(a)
    bh, bl, eb = f16_planes(b)
    return np.ldexp(f(al) @ f(bh) + f(ah) @ f(bl) + f(ah) @ f(bh), -(ea + eb))


def product_single_tf32(a, b):
    return tf32_rna(a).astype(np.float64) @ tf32_rna(b).astype(np.float64)


def rel_err(got, exact):
    return float(np.max(np.abs(got - exact)) / np.max(np.abs(exact)))

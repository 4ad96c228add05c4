"""The generated thread-resident IMDCT (pycricodecs_b200/csrc/hca_dct_thread_gen.inc, what hca_imdct_fast_kernel
runs per lane) compiled for the host with the CUDA intrinsics mapped to plain, uncontracted fp32 arithmetic, against
the oracle's imdct: PCM16 must be bit-equal, including the 32768 factor folded into the window constants."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r'''
#include <cstdint>
#include <cstring>
#include <cmath>
#define __device__
#define __forceinline__ inline
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
// the two-wide helpers of hca_fast_kernels.cu, as the scalar operations they stand for
static inline void hca_bfly2(float& a0, float& a1, float& b0, float& b1) {
    float s0 = __fadd_rn(a0, b0), s1 = __fadd_rn(a1, b1), d0 = __fsub_rn(a0, b0), d1 = __fsub_rn(a1, b1);
    a0 = s0; a1 = s1; b0 = d0; b1 = d1;
}
static inline void hca_sum2(unsigned long long, float p0, float p1, float q0, float q1, float& d0, float& d1) {
    d0 = __fadd_rn(p0, q0); d1 = __fadd_rn(p1, q1);
}
#include "hca_dct_thread_gen.inc"
extern "C" void run(const float* spectra, int n, int16_t* pcm) {
    float4 carry[16];
    memset(carry, 0, sizeof carry);
    for (int s = 0; s < n; s++) {
        float x[128];
        for (int i = 0; i < 128; i++) x[i] = spectra[s * 128 + i];
        hca_dct4_dec(x, 0ull, [](int) {});
        hca_window_thread<1>(x, carry, 0ull, [](float v) {
            float t = truncf(v);                       // cvt.rzi.s16.f32: truncate, saturate
            if (t > 32767.f) t = 32767.f;
            if (t < -32768.f) t = -32768.f;
            return (int16_t)t;
        }, [&](int i, int16_t v) { pcm[s * 128 + i] = v; }, [&](int c) { x[4 * c] = x[4 * c + 1] = x[4 * c + 2] = x[4 * c + 3] = 1e30f; },   // refilled registers must be dead
            [](int) {});
    }
}
'''


def test_generated_thread_imdct_matches_oracle(port):
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "shim.cpp")
        with open(src, "w") as f:
            f.write(SHIM)
        so = os.path.join(tmp, "shim.so")
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                        os.path.join(ROOT, "pycricodecs_b200", "csrc"), "-o", so, src], check=True)
        lib = ctypes.CDLL(so)
        rng = np.random.default_rng(7)
        n = 48
        for scale in (1e-30, 1e-4, 0.05, 1.0, 11.0):
            spec = (rng.standard_normal((n, 128)) * scale).astype(np.float32)
            spec[3] = 0
            spec[4, ::3] = 0
            pcm = np.zeros((n, 128), np.int16)
            lib.run(spec.ctypes.data_as(ctypes.c_void_p), n, pcm.ctypes.data_as(ctypes.c_void_p))
            prev = np.zeros(128, np.float32)
            for s in range(n):
                wave, prev, _ = port.imdct(spec[s].copy(), prev)
                v = wave.astype(np.float32) * np.float32(32768.0)       # clHCA_ReadSamples16 (hca.cpp:339-360)
                want = np.clip(np.trunc(v), -32768, 32767).astype(np.int16)
                assert np.array_equal(pcm[s], want), f"scale {scale}, subframe {s}"


PAIR_SHIM = SHIM.split('#include "hca_dct_thread_gen.inc"')[0] + r'''
#include "hca_dct_pair_gen.inc"
// The two halves of a warp pair, one after the other on the host: what a half publishes depends only on its own first six
// passes, so the partner's values are taken from a dry run of the partner (loads answered with zeros).
template <int H, class Load>
static void dct_half(float (&x)[64], float (&pub)[64], Load load) {
    auto store = [&](int q, float a, float b, float c, float d) { pub[4 * q] = a; pub[4 * q + 1] = b; pub[4 * q + 2] = c; pub[4 * q + 3] = d; };
    if (H == 0) hca_dct4_dec_h0(x, 0ull, store, [] {}, load, [](int) {});
    else hca_dct4_dec_h1(x, 0ull, store, [] {}, load, [](int) {});
}
extern "C" void run_pair(const float* spectra, int n, int16_t* pcm) {
    float4 carry[2][8];
    memset(carry, 0, sizeof carry);
    auto cvt = [](float v) {
        float t = truncf(v);
        if (t > 32767.f) t = 32767.f;
        if (t < -32768.f) t = -32768.f;
        return (int16_t)t;
    };
    for (int s = 0; s < n; s++) {
        float pub[2][64], x[2][64], scratch[64], dummy[64];
        for (int h = 0; h < 2; h++) {                                  // dry runs: what each half publishes
            for (int i = 0; i < 64; i++) scratch[i] = spectra[s * 128 + 64 * h + i];
            auto zero = [](int) { return float4{0, 0, 0, 0}; };
            if (h == 0) dct_half<0>(scratch, pub[0], zero); else dct_half<1>(scratch, pub[1], zero);
        }
        for (int h = 0; h < 2; h++) {
            for (int i = 0; i < 64; i++) x[h][i] = spectra[s * 128 + 64 * h + i];
            auto partner = [&](int q) { const float* p = pub[1 - h] + 4 * q; return float4{p[0], p[1], p[2], p[3]}; };
            if (h == 0) dct_half<0>(x[0], dummy, partner); else dct_half<1>(x[1], dummy, partner);
            int emitted = 0;
            auto emit = [&](int i, int16_t v) { pcm[s * 128 + i] = v; emitted++; };
            auto refill = [&](int c) { x[h][4 * c] = x[h][4 * c + 1] = x[h][4 * c + 2] = x[h][4 * c + 3] = 1e30f; };
            if (h == 0) hca_window_h0<1>(x[0], carry[0], 0ull, cvt, emit, refill);
            else hca_window_h1<1>(x[1], carry[1], 0ull, cvt, emit, refill);
            if (emitted != 64) pcm[s * 128] = 12345;
        }
    }
}
'''


def test_generated_pair_imdct_matches_oracle(port):
    """The warp-pair split of the same transform (hca_dct_pair_gen.inc, what hca_imdct_pair_kernel runs): two halves that
    exchange their 64 values once; PCM16 bit-equal with the oracle."""
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "shim.cpp")
        with open(src, "w") as f:
            f.write(PAIR_SHIM)
        so = os.path.join(tmp, "shim.so")
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                        os.path.join(ROOT, "pycricodecs_b200", "csrc"), "-o", so, src], check=True)
        lib = ctypes.CDLL(so)
        rng = np.random.default_rng(11)
        n = 40
        for scale in (1e-30, 1e-4, 0.05, 1.0, 11.0):
            spec = (rng.standard_normal((n, 128)) * scale).astype(np.float32)
            spec[2] = 0
            spec[5, ::5] = 0
            pcm = np.zeros((n, 128), np.int16)
            lib.run_pair(spec.ctypes.data_as(ctypes.c_void_p), n, pcm.ctypes.data_as(ctypes.c_void_p))
            prev = np.zeros(128, np.float32)
            for s in range(n):
                wave, prev, _ = port.imdct(spec[s].copy(), prev)
                v = wave.astype(np.float32) * np.float32(32768.0)
                want = np.clip(np.trunc(v), -32768, 32767).astype(np.int16)
                assert np.array_equal(pcm[s], want), f"scale {scale}, subframe {s}"


def test_generated_transform_is_current():
    """The tracked .inc files are what tools/gen_dct.py emits now (the kernels include them; a stale file would be
    compiled silently)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_dct", os.path.join(ROOT, "tools", "gen_dct.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    dct = g.gen_imdct(128, packed=True)
    end = dct.index("}")
    want = "\n".join(["// GENERATED by tools/gen_dct.py -- do not edit.", "#pragma once", ""] + dct[: end + 1] + [""]
                     + g.gen_thread_window(128, packed=True)) + "\n"
    have = open(os.path.join(ROOT, "pycricodecs_b200", "csrc", "hca_dct_thread_gen.inc")).read()
    assert have == want, "hca_dct_thread_gen.inc is stale: run tools/gen_dct.py"
    # every sum of two products goes through hca_sum2 (whose multiplier the compiler cannot see through): no packed
    # add of products that ptxas could contract, and every product is an explicitly rounded scalar multiply
    assert "add.rn.f32x2" not in have and "__fmaf" not in have and "fmaf(" not in have
    assert have.count("hca_sum2(") == 2 * 6 * 32 + 64 and have.count("hca_bfly2(") == 6 * 32
    pair = open(os.path.join(ROOT, "pycricodecs_b200", "csrc", "hca_dct_pair_gen.inc")).read()
    assert pair == "\n".join(g.gen_pair_file(128)) + "\n", "hca_dct_pair_gen.inc is stale: run tools/gen_dct.py"
    assert "add.rn.f32x2" not in pair and "__fmaf" not in pair and "fmaf(" not in pair

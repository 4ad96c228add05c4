#!/usr/bin/env python3
"""Emit the register-resident 128-point DCT-IV / IMDCT / MDCT networks used by
the HCA kernels as straight-line CUDA (pycricodecs_b200/csrc/hca_dct_gen.inc).

Why generated code: one thread owns one whole 128-point transform in registers.
Every stage of the network is a set of 64 two-in/two-out butterflies, so each
butterfly can overwrite its own inputs and the stage's index shuffle becomes a
compile-time renaming of registers -- no shared memory, no shuffles, no moves;
the instruction stream is the 4 k fp32 operations of the transform and nothing
else. That only works if every array index is a literal, hence this generator:
it tracks the logical->physical register permutation through the stages and
bakes the twiddle factors in as immediates (IEEE-754 bit patterns).

Arithmetic contract: the network, the operand order and the rounding points
are those of the reference (decoder: imdct_transform, CriCodecs/hca.cpp:1898-
1992; encoder: mdct_transform + DCT4, hca.cpp:2481-2553). Every product and
sum is a separate __fmul_rn / __fadd_rn / __fsub_rn, so nothing is contracted
into FMA and results are bit-identical to the reference's scalar SSE2 build.
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_tables as T  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def f(bits: int) -> str:
    return f"__uint_as_float(0x{int(bits):08X}u)"


def neg(bits: int) -> int:
    return int(bits) ^ 0x80000000


def gen_imdct(lockstep: int = 0, packed: bool = False) -> list[str]:
    """lockstep > 0: the DCT takes a functor and calls it after every `lockstep` fp32 instructions (the transform
    kernel uses it for a named barrier that keeps the warps of one scheduler on the same instruction-cache lines).

    packed: butterflies whose two registers differ in a bit above bit 0 come in pairs (a, b), (a + 1, b + 1); their
    sums and differences are emitted as two-wide operations on the register pairs (hca_bfly2 / hca_sum2, f32x2 on
    sm_100a: half the instructions for the same, separately rounded, results). Products stay scalar multiplies by
    immediates; a difference t0 - t1 becomes t0 + (x * -c), which is the same number."""
    sin, cos = T.imdct_trig()
    sin = sin.reshape(7, 64)
    cos = cos.reshape(7, 64)
    win = T.window()
    out = []
    phys = list(range(128))  # phys[logical index] = register index
    out.append("// 128-point DCT-IV of the HCA decoder, in place on registers x[0..127].")
    out.append("// Input x[i] = spectra[i]; afterwards dct[i] lives in x[kImdctPerm[i]] (see hca_imdct_window).")
    if lockstep:
        out.append("template <class Sync>")
        if packed:
            out.append("__device__ __forceinline__ void hca_dct4_dec(float (&x)[128], const unsigned long long one, Sync sync) {")
        else:
            out.append("__device__ __forceinline__ void hca_dct4_dec(float (&x)[128], Sync sync) {")
    else:
        out.append("__device__ __forceinline__ void hca_dct4_dec(float (&x)[128]) {")
    out.append("    float t0, t1, t2, t3;")
    if packed:
        out.append("    float u0, u1, u2, u3, r0, r1, r2, r3, w0, w1, w2, w3, y0, y1, y2, y3, z0, z1, z2, z3;")
    pending, flip, DEPTH = [], [0], 2
    since = 0
    nsync = [0]
    def sync_line():
        nsync[0] += 1
        return f"    sync({nsync[0]});"
    # sum/difference passes, half = 64 .. 1  (hca.cpp:1907-1935)
    half = 64
    while half >= 1:
        blocks = 64 // half
        nxt = [None] * 128
        pairs = set()
        for j in range(blocks):
            for k in range(half):
                pairs.add((phys[j * 2 * half + 2 * k], phys[j * 2 * half + 2 * k + 1]))
        for j in range(blocks):
            for k in range(half):
                a = phys[j * 2 * half + 2 * k]
                b = phys[j * 2 * half + 2 * k + 1]
                nxt[j * 2 * half + k] = a
                nxt[j * 2 * half + half + k] = b
                if packed and a % 2 == 0 and b % 2 == 0 and (a + 1, b + 1) in pairs:
                    out.append(f"    hca_bfly2(x[{a}], x[{a + 1}], x[{b}], x[{b + 1}]);")
                elif packed and a % 2 == 1 and b % 2 == 1 and (a - 1, b - 1) in pairs:
                    continue                                   # done with its even partner
                else:
                    out.append(f"    t0 = __fadd_rn(x[{a}], x[{b}]); x[{b}] = __fsub_rn(x[{a}], x[{b}]); x[{a}] = t0;")
                since += 2
                if lockstep and since >= lockstep:
                    out.append(sync_line())
                    since = 0
        phys = nxt
        half //= 2
    # rotation passes, half = 1 .. 64  (hca.cpp:1937-1972)
    for stage in range(7):
        half = 1 << stage
        blocks = 64 >> stage
        nxt = [None] * 128
        rot = {}
        for j in range(blocks):
            for k in range(half):
                rot[(phys[j * 2 * half + k], phys[j * 2 * half + half + k])] = (sin[stage, j * half + k], cos[stage, j * half + k])
        for j in range(blocks):
            for k in range(half):
                a = phys[j * 2 * half + k]
                b = phys[j * 2 * half + half + k]
                nxt[j * 2 * half + k] = a
                nxt[j * 2 * half + 2 * half - 1 - k] = b
                sb, cb = rot[(a, b)]
                if packed and a % 2 == 0 and b % 2 == 0 and (a + 1, b + 1) in rot:
                    sb2, cb2 = rot[(a + 1, b + 1)]
                    # software pipelining in the source: the sums of a group are emitted after the NEXT group's
                    # products (alternating temporaries), so no sum directly follows the products it waits for
                    t, u = (("t", "u"), ("r", "w"), ("y", "z"))[flip[0] % 3]
                    flip[0] += 1
                    out.append(f"    {t}0 = __fmul_rn(x[{a}], {f(sb)}); {t}1 = __fmul_rn(x[{b}], {f(neg(cb))}); "
                               f"{t}2 = __fmul_rn(x[{a}], {f(cb)}); {t}3 = __fmul_rn(x[{b}], {f(sb)});")
                    out.append(f"    {u}0 = __fmul_rn(x[{a + 1}], {f(sb2)}); {u}1 = __fmul_rn(x[{b + 1}], {f(neg(cb2))}); "
                               f"{u}2 = __fmul_rn(x[{a + 1}], {f(cb2)}); {u}3 = __fmul_rn(x[{b + 1}], {f(sb2)});")
                    if len(pending) >= DEPTH:
                        out.append(pending.pop(0))
                    pending.append(f"    hca_sum2(one, {t}0, {u}0, {t}1, {u}1, x[{a}], x[{a + 1}]); "
                                   f"hca_sum2(one, {t}2, {u}2, {t}3, {u}3, x[{b}], x[{b + 1}]);")
                    since += 10
                elif packed and a % 2 == 1 and b % 2 == 1 and (a - 1, b - 1) in rot:
                    continue
                else:
                    s_, c_ = f(sb), f(cb)
                    out.append(f"    t0 = __fmul_rn(x[{a}], {s_}); t1 = __fmul_rn(x[{b}], {c_}); "
                               f"t2 = __fmul_rn(x[{a}], {c_}); t3 = __fmul_rn(x[{b}], {s_}); "
                               f"x[{a}] = __fsub_rn(t0, t1); x[{b}] = __fadd_rn(t2, t3);")
                    since += 6
                if lockstep and since >= lockstep:
                    out.append(sync_line())
                    since = 0
        out += pending
        del pending[:]
        phys = nxt
    out.append("}")
    out.append("")
    # window + overlap (hca.cpp:1983-1992). dprev holds dct[0..63] of the previous subframe of this channel.
    out.append("// Window + overlap-add. `dprev[i]` = previous subframe's dct[i], i < 64 (imdct_previous is its windowed")
    out.append("// form: prev[i] = w[127-i]*dct[63-i], prev[64+i] = w[63-i]*dct[i]). Calls emit(i, wave[i]) for i = 0..127")
    out.append("// in sample order, then replaces dprev by this subframe's dct[0..63].")
    out.append("template <class Emit>")
    out.append("__device__ __forceinline__ void hca_imdct_window(const float (&x)[128], float (&dprev)[64], Emit emit) {")
    for i in range(64):
        out.append(f"    emit({i}, __fadd_rn(__fmul_rn({f(win[i])}, x[{phys[i + 64]}]), __fmul_rn({f(win[127 - i])}, dprev[{63 - i}])));")
    for i in range(64):
        out.append(f"    emit({i + 64}, __fsub_rn(__fmul_rn({f(win[i + 64])}, x[{phys[127 - i]}]), __fmul_rn({f(win[63 - i])}, dprev[{i}])));")
    for i in range(64):
        out.append(f"    dprev[{i}] = x[{phys[i]}];")
    out.append("}")
    out.append("")
    out.append("// Only the carry of hca_imdct_window (used for the look-back subframe in front of a run of frames).")
    out.append("__device__ __forceinline__ void hca_imdct_carry(const float (&x)[128], float (&dprev)[64]) {")
    for i in range(64):
        out.append(f"    dprev[{i}] = x[{phys[i]}];")
    out.append("}")
    return out



WSLOTS, CVT_LAG, STORE_LAG = 6, 2, 2


def gen_thread_window(lockstep: int = 0, packed: bool = False) -> list[str]:
    """Window + overlap-add + carry for the thread-resident transform (hca_imdct_fast_kernel), straight to PCM.

    The reference (hca.cpp:1983-1992) computes  wave[i] = w[i]*dct[64+i] + prev[i],  wave[64+i] = w[64+i]*dct[127-i]
    - prev[64+i]  with  prev[i] = w[127-i]*dprev[63-i],  prev[64+i] = w[63-i]*dprev[i]  (dprev = the previous
    subframe's dct[0..63]). Samples i and 127-i share dct[64+i] and dprev[63-i], so the code walks dprev in order,
    four values (one 16-byte shared-memory word) at a time, and overwrites each carry word with this subframe's
    dct[0..63] as soon as it has been consumed.

    The window constants carry the PCM scale 32768 (an exact power-of-two bump of the exponent). Products and sums
    scale exactly with it unless a product is subnormal, and a subnormal term cannot move the truncated integer
    (|term| < 2^-111 against a sum that is either >= 2^-102 * 32768, where it is below half an ulp, or so small
    that both versions truncate to 0), so  trunc(wave * 32768)  is unchanged.
    """
    win = T.window()
    # register of dct[k] after hca_dct4_dec: replay the permutation bookkeeping of gen_imdct
    phys = list(range(128))
    half = 64
    while half >= 1:
        nxt = [None] * 128
        for j in range(64 // half):
            for k in range(half):
                a, b = phys[j * 2 * half + 2 * k], phys[j * 2 * half + 2 * k + 1]
                nxt[j * 2 * half + k], nxt[j * 2 * half + half + k] = a, b
        phys = nxt
        half //= 2
    for stage in range(7):
        half = 1 << stage
        nxt = [None] * 128
        for j in range(64 >> stage):
            for k in range(half):
                a, b = phys[j * 2 * half + k], phys[j * 2 * half + half + k]
                nxt[j * 2 * half + k], nxt[j * 2 * half + 2 * half - 1 - k] = a, b
        phys = nxt

    def scaled(bits: int) -> str:
        bits = int(bits)
        e = (bits >> 23) & 0xFF
        assert 1 <= e <= 200, "window constant must be a normal float"
        return f(bits + (15 << 23))

    # which carry group (four consecutive dprev values = one float4) frees which register
    inv = [0] * 128
    for m, r in enumerate(phys):
        inv[r] = m
    group_of = [(inv[r] // 4) if inv[r] < 64 else ((127 - inv[r]) // 4) for r in range(128)]
    out = []
    out.append("// Window + overlap-add of one subframe held in registers after hca_dct4_dec, emitted as PCM-scaled floats:")
    out.append("// emit(i, v) receives v = wave[i] * 32768 for every sample i (in the order the carry is walked). `carry` is")
    out.append("// this thread's column of the shared carry array ([16][CARRY_STRIDE] float4 = dct[0..63] of the previous")
    out.append("// subframe); it is replaced by this subframe's dct[0..63]. The carry groups are walked in the order 0, 15, 1, 14,")
    out.append("// ...: after each pair exactly four aligned register quads x[4c..4c+3] are dead, and refill(c) is called for")
    out.append("// each so that the caller can already load chunk c of the NEXT subframe's spectra into them (the registers")
    out.append("// are full during the transform, so this is the only place a prefetch can live).")
    if lockstep and packed:
        out.append("// Packed form: cvt(v) turns a PCM-scaled float into the value emit(i, s) stores.")
        out.append("template <int CARRY_STRIDE, class Cvt, class Emit, class Refill, class Sync>")
        out.append("__device__ __forceinline__ void hca_window_thread(float (&x)[128], float4* carry, const unsigned long long one, Cvt cvt, Emit emit, Refill refill, Sync sync) {")
        out.append("    float p0, p1, q0, q1;")
        out.append("    float " + ", ".join(f"va{k}, vb{k}" for k in range(WSLOTS)) + ";")
        out.append("    decltype(cvt(0.f)) " + ", ".join(f"sa{k}, sb{k}" for k in range(WSLOTS)) + ";")
    elif lockstep:
        out.append("template <int CARRY_STRIDE, class Emit, class Refill, class Sync>")
        out.append("__device__ __forceinline__ void hca_window_thread(float (&x)[128], float4* carry, Emit emit, Refill refill, Sync sync) {")
    else:
        out.append("template <int CARRY_STRIDE, class Emit, class Refill>")
        out.append("__device__ __forceinline__ void hca_window_thread(float (&x)[128], float4* carry, Emit emit, Refill refill) {")
    out.append("    float4 c;")
    wpend, wcount = [], [0]
    done = set()
    order = []
    for p_ in range(8):
        order += [p_, 15 - p_]
    refilled = set()
    for q in order:
        out.append(f"    c = carry[{q} * CARRY_STRIDE];")
        for e, comp in enumerate("xyzw"):
            k = 4 * q + e
            i = 63 - k
            d = f"x[{phys[64 + i]}]"
            wa, wb = scaled(win[i]), scaled(win[127 - i])
            if packed:
                # (wave[i], wave[127-i]) = (wa*d, wb*d) + (wb*c, -wa*c): two products each, one two-wide sum
                wan = scaled(neg(win[i]))
                out.append(f"    p0 = __fmul_rn({wa}, {d}); p1 = __fmul_rn({wb}, {d}); q0 = __fmul_rn({wb}, c.{comp}); q1 = __fmul_rn({wan}, c.{comp});")
                # the conversion to int16 (F2I, a long-latency unit) runs CVT_LAG elements behind the sums and the
                # shared-memory stores STORE_LAG elements behind the conversions, on rotating temporaries
                slot = wcount[0] % WSLOTS
                wcount[0] += 1
                out.append(f"    hca_sum2(one, p0, p1, q0, q1, va{slot}, vb{slot});")
                wpend.append((i, slot))
                if len(wpend) > CVT_LAG:
                    j, sl = wpend[-1 - CVT_LAG]
                    out.append(f"    sa{sl} = cvt(va{sl}); sb{sl} = cvt(vb{sl});")
                if len(wpend) > CVT_LAG + STORE_LAG:
                    j, sl = wpend[-1 - CVT_LAG - STORE_LAG]
                    out.append(f"    emit({j}, sa{sl}); emit({127 - j}, sb{sl});")
            else:
                out.append(f"    emit({i}, __fadd_rn(__fmul_rn({wa}, {d}), __fmul_rn({wb}, c.{comp})));")
                out.append(f"    emit({127 - i}, __fsub_rn(__fmul_rn({wb}, {d}), __fmul_rn({wa}, c.{comp})));")
        out.append(f"    carry[{q} * CARRY_STRIDE] = make_float4(x[{phys[4 * q]}], x[{phys[4 * q + 1]}], x[{phys[4 * q + 2]}], x[{phys[4 * q + 3]}]);")
        done.add(q)
        for ch in range(32):
            if ch not in refilled and all(group_of[4 * ch + e] in done for e in range(4)):
                out.append(f"    refill({ch});")
                refilled.add(ch)
        if lockstep and q in order[1::2]:
            out.append(f"    sync({100 + order.index(q) // 2});")
    assert len(refilled) == 32
    if packed:                                   # drain the conversion / store pipeline
        n = len(wpend)
        for k in range(n - CVT_LAG, n):
            j, sl = wpend[k]
            out.append(f"    sa{sl} = cvt(va{sl}); sb{sl} = cvt(vb{sl});")
        for k in range(n - CVT_LAG - STORE_LAG, n):
            j, sl = wpend[k]
            out.append(f"    emit({j}, sa{sl}); emit({127 - j}, sb{sl});")
    out.append("}")
    out.append("")
    out.append("// Carry only (the look-back subframe in front of a run of frames).")
    out.append("template <int CARRY_STRIDE>")
    out.append("__device__ __forceinline__ void hca_carry_thread(const float (&x)[128], float4* carry) {")
    for q in range(16):
        out.append(f"    carry[{q} * CARRY_STRIDE] = make_float4(x[{phys[4 * q]}], x[{phys[4 * q + 1]}], x[{phys[4 * q + 2]}], x[{phys[4 * q + 3]}]);")
    out.append("}")
    return out


def _final_mapping():
    """phys[m] = register (= input coefficient index) that holds dct[m] after the fourteen passes."""
    phys = list(range(128))
    half = 64
    while half >= 1:
        nxt = [None] * 128
        for j in range(64 // half):
            for k in range(half):
                a, b = phys[j * 2 * half + 2 * k], phys[j * 2 * half + 2 * k + 1]
                nxt[j * 2 * half + k], nxt[j * 2 * half + half + k] = a, b
        phys = nxt
        half //= 2
    for stage in range(7):
        half = 1 << stage
        nxt = [None] * 128
        for j in range(64 >> stage):
            for k in range(half):
                a, b = phys[j * 2 * half + k], phys[j * 2 * half + half + k]
                nxt[j * 2 * half + k], nxt[j * 2 * half + 2 * half - 1 - k] = a, b
        phys = nxt
    return phys


def gen_imdct_half(h: int, lockstep: int = 0) -> list[str]:
    """One HALF of the decoder's 128-point DCT-IV for the warp-pair transform kernel (hca_imdct_pair_kernel): a lane
    of warp h of a pair keeps the coefficients with register index 64 h .. 64 h + 63 (= input coefficients, the network
    never moves a value out of its register). The fourteen passes pair registers that differ in bit 0, 1, .., 6 (sums
    and differences) and then 6, 5, .., 0 (rotations), so only the two passes in the middle cross the halves -- and
    they work on the SAME pairs (r, r + 64): sum / difference a' = a + b, b' = a - b, then the rotation a'' = a' sin -
    b' cos, b'' = a' cos + b' sin. A lane therefore publishes its 64 values once (store), the pair meets (barrier), and
    each lane recomputes a' and b' from its own and its partner's value and keeps the output it owns. The two halves
    run different instruction streams (every rotation factor is an immediate), which is why the split is by warp.

    Arithmetic contract as gen_imdct: the reference network (hca.cpp:1898-1972), operand order and rounding points
    unchanged; two-wide forms (hca_bfly2 / hca_sum2) for sums whose registers are an even/odd pair."""
    sin, cos = T.imdct_trig()
    sin = sin.reshape(7, 64)
    cos = cos.reshape(7, 64)
    base = 64 * h
    out = []
    out.append(f"// Half {h} (registers {base}..{base + 63}) of the 128-point DCT-IV of the HCA decoder, in place on x[0..63] = coefficient {base} + r.")
    out.append("// store(q, a, b, c, d): publish own values 4q..4q+3 after the first six passes; barrier(): the pair meets;")
    out.append("// load(q): the partner's values 4q..4q+3. Afterwards dct[m] lives in x[phys(m) - base] (see hca_window_half).")
    out.append("template <class Store, class Barrier, class Load, class Sync>")
    out.append(f"__device__ __forceinline__ void hca_dct4_dec_h{h}(float (&x)[64], const unsigned long long one, Store store, Barrier barrier, Load load, Sync sync) {{")
    out.append("    float t0, t1, t2, t3, u0, u1, u2, u3, r0, r1, r2, r3, w0, w1, w2, w3, y0, y1, y2, y3, z0, z1, z2, z3;")
    out.append("    float4 p;")
    since = 0
    nsync = [0]

    def sync_line():
        nsync[0] += 1
        return f"    sync({nsync[0]});"
    phys = list(range(128))
    half = 64
    cross = None
    while half >= 1:
        blocks = 64 // half
        nxt = [None] * 128
        pairs = set()
        for j in range(blocks):
            for k in range(half):
                pairs.add((phys[j * 2 * half + 2 * k], phys[j * 2 * half + 2 * k + 1]))
        if half == 1:
            cross = sorted(pairs)
            assert all(b == a + 64 and a < 64 for a, b in cross)
        for j in range(blocks):
            for k in range(half):
                a = phys[j * 2 * half + 2 * k]
                b = phys[j * 2 * half + 2 * k + 1]
                nxt[j * 2 * half + k] = a
                nxt[j * 2 * half + half + k] = b
                if half == 1 or a // 64 != h:
                    continue
                assert b // 64 == h
                la, lb = a - base, b - base
                if a % 2 == 0 and b % 2 == 0 and (a + 1, b + 1) in pairs:
                    out.append(f"    hca_bfly2(x[{la}], x[{la + 1}], x[{lb}], x[{lb + 1}]);")
                elif a % 2 == 1 and b % 2 == 1 and (a - 1, b - 1) in pairs:
                    continue
                else:
                    out.append(f"    t0 = __fadd_rn(x[{la}], x[{lb}]); x[{lb}] = __fsub_rn(x[{la}], x[{lb}]); x[{la}] = t0;")
                since += 2
                if lockstep and since >= lockstep:
                    out.append(sync_line())
                    since = 0
        phys = nxt
        half //= 2
    # ---- the two passes on bit 6: publish, meet, recompute both sums from own + partner, keep the own output
    for q in range(16):
        out.append(f"    store({q}, x[{4 * q}], x[{4 * q + 1}], x[{4 * q + 2}], x[{4 * q + 3}]);")
    out.append("    barrier();")
    rot0 = {}
    for j in range(64):
        a, b = phys[j * 2], phys[j * 2 + 1]
        assert b == a + 64 and a < 64
        rot0[a] = (sin[0, j], cos[0, j])
    nxt = [None] * 128
    for j in range(64):
        a, b = phys[j * 2], phys[j * 2 + 1]
        nxt[j * 2], nxt[j * 2 + 1] = a, b
    phys = nxt
    for q in range(16):
        out.append(f"    p = load({q});")
        for e, (c0, c1) in enumerate((("x", "y"), ("z", "w"))):
            r = 4 * q + 2 * e
            (s_a, c_a), (s_b, c_b) = rot0[r], rot0[r + 1]
            if h == 0:
                # own = a, partner = b:  a' = a + b -> x, b' = a - b -> p;  keep a'' = a' sin - b' cos
                out.append(f"    hca_bfly2(x[{r}], x[{r + 1}], p.{c0}, p.{c1});")
                out.append(f"    t0 = __fmul_rn(x[{r}], {f(s_a)}); t1 = __fmul_rn(p.{c0}, {f(neg(c_a))}); "
                           f"u0 = __fmul_rn(x[{r + 1}], {f(s_b)}); u1 = __fmul_rn(p.{c1}, {f(neg(c_b))});")
            else:
                # own = b, partner = a:  a' = a + b -> p, b' = a - b -> x;  keep b'' = a' cos + b' sin
                out.append(f"    hca_bfly2(p.{c0}, p.{c1}, x[{r}], x[{r + 1}]);")
                out.append(f"    t0 = __fmul_rn(p.{c0}, {f(c_a)}); t1 = __fmul_rn(x[{r}], {f(s_a)}); "
                           f"u0 = __fmul_rn(p.{c1}, {f(c_b)}); u1 = __fmul_rn(x[{r + 1}], {f(s_b)});")
            out.append(f"    hca_sum2(one, t0, u0, t1, u1, x[{r}], x[{r + 1}]);")
            since += 7
            if lockstep and since >= lockstep:
                out.append(sync_line())
                since = 0
    # ---- rotation passes on bits 5 .. 0 (stages 1 .. 6), local to the half
    pending, flip, DEPTH = [], [0], 2
    for stage in range(1, 7):
        half = 1 << stage
        blocks = 64 >> stage
        nxt = [None] * 128
        rot = {}
        for j in range(blocks):
            for k in range(half):
                rot[(phys[j * 2 * half + k], phys[j * 2 * half + half + k])] = (sin[stage, j * half + k], cos[stage, j * half + k])
        for j in range(blocks):
            for k in range(half):
                a = phys[j * 2 * half + k]
                b = phys[j * 2 * half + half + k]
                nxt[j * 2 * half + k] = a
                nxt[j * 2 * half + 2 * half - 1 - k] = b
                if a // 64 != h:
                    continue
                assert b // 64 == h
                la, lb = a - base, b - base
                sb, cb = rot[(a, b)]
                if a % 2 == 0 and b % 2 == 0 and (a + 1, b + 1) in rot:
                    sb2, cb2 = rot[(a + 1, b + 1)]
                    t, u = (("t", "u"), ("r", "w"), ("y", "z"))[flip[0] % 3]
                    flip[0] += 1
                    out.append(f"    {t}0 = __fmul_rn(x[{la}], {f(sb)}); {t}1 = __fmul_rn(x[{lb}], {f(neg(cb))}); "
                               f"{t}2 = __fmul_rn(x[{la}], {f(cb)}); {t}3 = __fmul_rn(x[{lb}], {f(sb)});")
                    out.append(f"    {u}0 = __fmul_rn(x[{la + 1}], {f(sb2)}); {u}1 = __fmul_rn(x[{lb + 1}], {f(neg(cb2))}); "
                               f"{u}2 = __fmul_rn(x[{la + 1}], {f(cb2)}); {u}3 = __fmul_rn(x[{lb + 1}], {f(sb2)});")
                    if len(pending) >= DEPTH:
                        out.append(pending.pop(0))
                    pending.append(f"    hca_sum2(one, {t}0, {u}0, {t}1, {u}1, x[{la}], x[{la + 1}]); "
                                   f"hca_sum2(one, {t}2, {u}2, {t}3, {u}3, x[{lb}], x[{lb + 1}]);")
                    since += 10
                elif a % 2 == 1 and b % 2 == 1 and (a - 1, b - 1) in rot:
                    continue
                else:
                    s_, c_ = f(sb), f(cb)
                    out.append(f"    t0 = __fmul_rn(x[{la}], {s_}); t1 = __fmul_rn(x[{lb}], {c_}); "
                               f"t2 = __fmul_rn(x[{la}], {c_}); t3 = __fmul_rn(x[{lb}], {s_}); "
                               f"x[{la}] = __fsub_rn(t0, t1); x[{lb}] = __fadd_rn(t2, t3);")
                    since += 6
                if lockstep and since >= lockstep:
                    out.append(sync_line())
                    since = 0
        out += pending
        del pending[:]
        phys = nxt
    assert phys == _final_mapping()
    out.append("}")
    return out


def gen_window_half(h: int) -> list[str]:
    """Window + overlap-add + carry of one half (hca.cpp:1983-1992), straight to PCM-scaled floats. After the DCT an odd
    register of the half holds dct[64 + i] and its even neighbour dct[63 - i] for the same i, so a lane produces the
    samples i and 127 - i of its 32 values of i and carries its own 32 values of dct[0..63]: nothing crosses the halves.
    carry[g] (g = 0..7, stride CARRY_STRIDE float4) holds the previous subframe's even registers 8g, 8g+2, 8g+4, 8g+6;
    after group g the registers 8g..8g+7 are dead and refill(2g), refill(2g+1) may load the next subframe's chunks."""
    win = T.window()
    phys = _final_mapping()
    final = [0] * 128
    for m, r in enumerate(phys):
        final[r] = m
    base = 64 * h

    def scaled(bits: int) -> str:
        bits = int(bits)
        e = (bits >> 23) & 0xFF
        assert 1 <= e <= 200, "window constant must be a normal float"
        return f(bits + (15 << 23))
    out = []
    out.append(f"// Window + overlap-add of half {h}: emit(i, s) for the lane's 64 samples, carry replaced, refill(c) for dead quads.")
    out.append("template <int CARRY_STRIDE, class Cvt, class Emit, class Refill>")
    out.append(f"__device__ __forceinline__ void hca_window_h{h}(float (&x)[64], float4* carry, const unsigned long long one, Cvt cvt, Emit emit, Refill refill) {{")
    out.append("    float p0, p1, q0, q1;")
    out.append("    float " + ", ".join(f"va{k}, vb{k}" for k in range(WSLOTS)) + ";")
    out.append("    decltype(cvt(0.f)) " + ", ".join(f"sa{k}, sb{k}" for k in range(WSLOTS)) + ";")
    out.append("    float4 c;")
    wpend, wcount = [], [0]
    for g in range(8):
        out.append(f"    c = carry[{g} * CARRY_STRIDE];")
        for e, comp in enumerate("xyzw"):
            ev = 8 * g + 2 * e
            od = ev + 1
            j = final[base + od]
            assert j >= 64 and final[base + ev] == 127 - j
            i = j - 64
            d = f"x[{od}]"
            wa, wb, wan = scaled(win[i]), scaled(win[127 - i]), scaled(neg(win[i]))
            out.append(f"    p0 = __fmul_rn({wa}, {d}); p1 = __fmul_rn({wb}, {d}); q0 = __fmul_rn({wb}, c.{comp}); q1 = __fmul_rn({wan}, c.{comp});")
            slot = wcount[0] % WSLOTS
            wcount[0] += 1
            out.append(f"    hca_sum2(one, p0, p1, q0, q1, va{slot}, vb{slot});")
            wpend.append((i, slot))
            if len(wpend) > CVT_LAG:
                jj, sl = wpend[-1 - CVT_LAG]
                out.append(f"    sa{sl} = cvt(va{sl}); sb{sl} = cvt(vb{sl});")
            if len(wpend) > CVT_LAG + STORE_LAG:
                jj, sl = wpend[-1 - CVT_LAG - STORE_LAG]
                out.append(f"    emit({jj}, sa{sl}); emit({127 - jj}, sb{sl});")
        out.append(f"    carry[{g} * CARRY_STRIDE] = make_float4(x[{8 * g}], x[{8 * g + 2}], x[{8 * g + 4}], x[{8 * g + 6}]);")
        out.append(f"    refill({2 * g}); refill({2 * g + 1});")
    n = len(wpend)
    for k in range(n - CVT_LAG, n):
        jj, sl = wpend[k]
        out.append(f"    sa{sl} = cvt(va{sl}); sb{sl} = cvt(vb{sl});")
    for k in range(n - CVT_LAG - STORE_LAG, n):
        jj, sl = wpend[k]
        out.append(f"    emit({jj}, sa{sl}); emit({127 - jj}, sb{sl});")
    out.append("}")
    out.append("")
    out.append(f"// Carry only (the look-back subframe in front of a run of frames), half {h}.")
    out.append("template <int CARRY_STRIDE>")
    out.append(f"__device__ __forceinline__ void hca_carry_h{h}(const float (&x)[64], float4* carry) {{")
    for g in range(8):
        out.append(f"    carry[{g} * CARRY_STRIDE] = make_float4(x[{8 * g}], x[{8 * g + 2}], x[{8 * g + 4}], x[{8 * g + 6}]);")
    out.append("}")
    return out


def gen_pair_file(lockstep: int = 128) -> list[str]:
    lines = ["// GENERATED by tools/gen_dct.py -- do not edit.", "#pragma once", ""]
    for h in (0, 1):
        lines += gen_imdct_half(h, lockstep) + [""] + gen_window_half(h) + [""]
    return lines


def gen_mdct() -> list[str]:
    sin, cos = T.mdct_trig()
    sin = sin.reshape(8, 128)
    cos = cos.reshape(8, 128)
    win = T.window()
    shuffle = T.enc_shuffle()
    out = []
    out.append("// Forward MDCT of the HCA encoder (hca.cpp:2529-2553 + DCT4 :2481-2527), on registers.")
    out.append("// cur[i] = this subframe's 128 input samples (as float), prv[i] = the previous 128; t[] is scratch.")
    out.append("// Calls emit(k, spectra[k]) for k = 0..127.")
    out.append("template <class Emit>")
    out.append("__device__ __forceinline__ void hca_mdct_enc(const float (&cur)[128], const float (&prv)[128], float (&t)[128], Emit emit) {")
    out.append("    float a, b, c0, c1, c2, c3;")
    # windowing into in[] (kept in t[] with identity placement), hca.cpp:2537-2546
    # in[i] = W[63-i]*(-cur[64+i]) - (-W[64+i])*cur[63-i];  in[64+i] = W[i]*prv[i] - (-W[127-i])*prv[127-i]
    for i in range(64):
        out.append(f"    t[{i}] = __fsub_rn(__fmul_rn({f(win[63 - i])}, -cur[{64 + i}]), __fmul_rn({f(int(win[64 + i]) ^ 0x80000000)}, cur[{63 - i}]));")
        out.append(f"    t[{64 + i}] = __fsub_rn(__fmul_rn({f(win[i])}, prv[{i}]), __fmul_rn({f(int(win[127 - i]) ^ 0x80000000)}, prv[{127 - i}]));")
    # pre-rotation: dct[2i] = a*cos + b*sin, dct[2i+1] = a*sin - b*cos with a = in[2i], b = in[127-2i]  (row 7)
    # both outputs depend on in[2i] and in[127-2i]; outputs go to positions 2i and 2i+1. Pairs (2i, 127-2i) for i<64 cover
    # all 128 inputs exactly once, so do it in place: out[2i] -> reg of in[2i], out[2i+1] -> reg of in[127-2i].
    phys = list(range(128))
    nxt = [None] * 128
    for i in range(64):
        ra, rb = phys[2 * i], phys[127 - 2 * i]
        s = f(sin[7, i]); c = f(cos[7, i])
        out.append(f"    a = t[{ra}]; b = t[{rb}]; c0 = __fmul_rn(a, {c}); c1 = __fmul_rn(b, {s}); c2 = __fmul_rn(a, {s}); c3 = __fmul_rn(b, {c}); "
                   f"t[{ra}] = __fadd_rn(c0, c1); t[{rb}] = __fsub_rn(c2, c3);")
        nxt[2 * i] = ra
        nxt[2 * i + 1] = rb
    phys = nxt
    # six in-place radix-2 stages (hca.cpp:2501-2523); index pattern is already in place in the reference
    for stage in range(6):
        blocks = 1 << stage
        size_bits = 6 - stage
        half_bits = size_bits - 1
        size = 1 << size_bits
        half = 1 << half_bits
        for blk in range(blocks):
            for i in range(half):
                fp = (blk * size + i) * 2
                bp = fp + size
                s = f(sin[half_bits, i]); c = f(cos[half_bits, i])
                f0, f1, b0, b1 = phys[fp], phys[fp + 1], phys[bp], phys[bp + 1]
                out.append(f"    a = __fsub_rn(t[{f0}], t[{b0}]); b = __fsub_rn(t[{f1}], t[{b1}]); "
                           f"t[{f0}] = __fadd_rn(t[{f0}], t[{b0}]); t[{f1}] = __fadd_rn(t[{f1}], t[{b1}]); "
                           f"c0 = __fmul_rn(a, {c}); c1 = __fmul_rn(b, {s}); c2 = __fmul_rn(a, {s}); c3 = __fmul_rn(b, {c}); "
                           f"t[{b0}] = __fadd_rn(c0, c1); t[{b1}] = __fsub_rn(c2, c3);")
    for k in range(128):
        out.append(f"    emit({k}, __fmul_rn(t[{phys[int(shuffle[k])]}], 0.125f));")
    out.append("}")
    return out


def gen_warp_tables() -> list[str]:
    """Tables for the warp-cooperative transform (hca_imdct_kernel): coefficient p of a 128-point block lives in
    lane p>>2, register p&3, and never moves. The seven sum/difference passes then pair slots that differ in bit
    0,1,..,6 of p, the seven rotation passes pair bit 6,5,..,0, and the window pairs bit 0 -- i.e. every exchange is
    a register swap or one __shfl_xor. Per slot and rotation pass the update is  v*S + partner*C  with the sign of
    the reference's subtraction folded into C (x - y == x + (-y) exactly in IEEE arithmetic)."""
    sin, cos = T.imdct_trig()
    sin = sin.reshape(7, 64)
    cos = cos.reshape(7, 64)
    win = T.window()
    phys = list(range(128))
    half = 64
    bit = 0
    while half >= 1:
        nxt = [None] * 128
        for j in range(64 // half):
            for k in range(half):
                a, b = phys[j * 2 * half + 2 * k], phys[j * 2 * half + 2 * k + 1]
                assert a ^ b == 1 << bit and not a & (1 << bit)
                nxt[j * 2 * half + k], nxt[j * 2 * half + half + k] = a, b
        phys = nxt
        half //= 2
        bit += 1
    rs = [[0] * 128 for _ in range(7)]
    rc = [[0] * 128 for _ in range(7)]
    for stage in range(7):
        half = 1 << stage
        nxt = [None] * 128
        for j in range(64 >> stage):
            for k in range(half):
                a, b = phys[j * 2 * half + k], phys[j * 2 * half + half + k]
                assert a ^ b == 1 << (6 - stage)
                s, c = int(sin[stage, j * half + k]), int(cos[stage, j * half + k])
                rs[stage][a], rc[stage][a] = s, c ^ 0x80000000      # a*sin - b*cos
                rs[stage][b], rc[stage][b] = s, c                   # b*sin + a*cos
                nxt[j * 2 * half + k], nxt[j * 2 * half + 2 * half - 1 - k] = a, b
        phys = nxt
    final = [0] * 128
    for logical, p in enumerate(phys):
        final[p] = logical
    # window: odd slots hold dct[j], j >= 64, their even neighbour holds dct[127-j] (carried to the next subframe)
    w1, w2, s1, s2 = [0] * 64, [0] * 64, [0] * 64, [0] * 64
    for p in range(1, 128, 2):
        j = final[p]
        assert j >= 64 and final[p ^ 1] == 127 - j
        w1[p >> 1], w2[p >> 1] = int(win[j - 64]), int(win[191 - j])
        s1[p >> 1], s2[p >> 1] = j - 64, 191 - j
    out = ["// per-slot tables of the warp-cooperative IMDCT (slot p = 4*lane + register)"]

    def emit(name, ctype, vals, fmt):
        out.append(f"__device__ const {ctype} {name}[{len(vals)}] = {{")
        for i in range(0, len(vals), 8):
            out.append("    " + ", ".join(fmt(v) for v in vals[i:i + 8]) + ",")
        out.append("};")
    emit("kRotS", "uint32_t", [v for row in rs for v in row], lambda v: f"0x{v:08X}u")
    emit("kRotC", "uint32_t", [v for row in rc for v in row], lambda v: f"0x{v:08X}u")
    emit("kWinA", "uint32_t", w1, lambda v: f"0x{v:08X}u")   # weight of sample kWinPosA: w[j-64]
    emit("kWinB", "uint32_t", w2, lambda v: f"0x{v:08X}u")   # weight of sample kWinPosB: w[191-j]
    emit("kWinPosA", "uint8_t", s1, str)
    emit("kWinPosB", "uint8_t", s2, str)
    return out


def gen_mdct_warp_tables() -> list[str]:
    """Tables for the warp-cooperative forward MDCT (hca_encode_kernel). The reference's DCT4 (hca.cpp:2481-2527) is a
    pre-rotation into 64 complex values z[k] = (t[2k], t[2k+1]) followed by a 64-point radix-2 decimation-in-frequency
    network (pair distance 32,16,..,1) and a bit-reversal gather. Lane l keeps z[l] and z[l+32]: the first pass is
    in-lane, the other five are one __shfl_xor each; both of a lane's values share the pass's rotation factor."""
    sin, cos = T.mdct_trig()
    sin = sin.reshape(8, 128)
    cos = cos.reshape(8, 128)
    win = T.window()
    shuffle = [int(v) for v in T.enc_shuffle()]
    inv = [0] * 128
    for i, t in enumerate(shuffle):
        inv[t] = i
    out = ["// per-lane tables of the warp-cooperative forward MDCT (lane l holds z[l] and z[l+32])"]

    def emit(name, ctype, vals, fmt):
        out.append(f"__device__ const {ctype} {name}[{len(vals)}] = {{")
        for i in range(0, len(vals), 8):
            out.append("    " + ", ".join(fmt(v) for v in vals[i:i + 8]) + ",")
        out.append("};")
    hx = lambda v: f"0x{int(v):08X}u"
    emit("kMdctPreCos", "uint32_t", [cos[7, k] for k in range(64)], hx)
    emit("kMdctPreSin", "uint32_t", [sin[7, k] for k in range(64)], hx)
    # pass s (0..5): pair distance d = 32 >> s, factor row 5 - s, index k mod d (same for k = l and k = l + 32)
    emit("kMdctCos", "uint32_t", [cos[5 - s, l % (32 >> s)] for s in range(6) for l in range(32)], hx)
    emit("kMdctSin", "uint32_t", [sin[5 - s, l % (32 >> s)] for s in range(6) for l in range(32)], hx)
    # window factors a lane needs: W[2l], W[63-2l], W[64+2l], W[127-2l]
    emit("kMdctWin", "uint32_t", [win[i] for l in range(32) for i in (2 * l, 63 - 2 * l, 64 + 2 * l, 127 - 2 * l)], hx)
    # where the lane's four results t[2l], t[2l+1], t[2l+64], t[2l+65] go in the spectrum
    emit("kMdctDest", "uint8_t", [inv[i] for l in range(32) for i in (2 * l, 2 * l + 1, 2 * l + 64, 2 * l + 65)], str)
    return out


def main():
    lines = ["// GENERATED by tools/gen_dct.py -- do not edit.", "#pragma once", ""]
    lines += gen_warp_tables()
    lines.append("")
    lines += gen_mdct_warp_tables()
    path = os.path.join(ROOT, "pycricodecs_b200", "csrc", "hca_dct_gen.inc")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", path, len(lines), "lines")
    # thread-resident decoder transform (hca_imdct_fast_kernel): only the DCT-IV of gen_imdct + the PCM window
    LOCKSTEP = 128                          # the transform calls sync(k) every so many fp32 instructions (instruction-cache convoy)
    dct = gen_imdct(LOCKSTEP, packed=True)
    end = dct.index("}")                    # first function = hca_dct4_dec
    lines = ["// GENERATED by tools/gen_dct.py -- do not edit.", "#pragma once", ""] + dct[: end + 1] + [""] + gen_thread_window(LOCKSTEP, packed=True)
    path = os.path.join(ROOT, "pycricodecs_b200", "csrc", "hca_dct_thread_gen.inc")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", path, len(lines), "lines")
    # the same transform split over a pair of warps (hca_imdct_pair_kernel)
    lines = gen_pair_file(LOCKSTEP)
    path = os.path.join(ROOT, "pycricodecs_b200", "csrc", "hca_dct_pair_gen.inc")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", path, len(lines), "lines")


if __name__ == "__main__":
    main()

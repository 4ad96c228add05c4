"""CPU: generated constant tables are pinned by digest and equal the reference's literal arrays."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_tables as G  # noqa: E402

PINNED = {
    "CRC16": "756cf0d79c4503b3", "DEC_SCALING": "80eaaab769802ab8", "DEC_RANGE": "594acbd539f7b4dd", "SCALE_CONV": "ce3889e5db50fcd8",
    "INTENSITY_RATIO": "6f53503da52ecfb0", "IMDCT_SIN": "bf110e682f8cec86", "IMDCT_COS": "72d09011bd727778", "WINDOW": "fd4b9f1973aa0a78",
    "INVERT": "596fe25e69de6eab", "MAX_BITS": "bccecd5122d1510e", "READ_BITS": "2e1cc91f0d50f0d4", "READ_VALS": "2cc4f662b2917a8c",
    "ENC_RES_CURVE": "ae26f4723b8c0df3", "ENC_Q_BITS": "4dea801cd96da815", "ENC_Q_CODE": "7c9c807b445fc2a3", "ENC_INV_STEP": "cd2e4715a8858fa3",
    "ENC_DEAD_ZONE": "46be60ec7b43774b", "ENC_RATIO_BOUNDS": "7b0403a9464abe57", "ENC_Q_SCALING": "40bb54ebaca01923",
    "ENC_COST_ROWS": "ad41c2373ac9383d", "ENC_RANK_KEYS": "93e36edbab3de372", "ENC_RANK_ROWS": "65aff724ccb0bb05",
    "MDCT_SIN": "f6e0ce2a262954bb", "MDCT_COS": "5cd990a6f62da25a", "ENC_SHUFFLE": "143d219b1e658c0b", "ADX_STATIC_COEF": "f0a57863809129eb",
    "ATH_BASE": "103a013614c0f314",
}


def test_table_digests_are_pinned():
    got = {name: G.digest(arr) for name, _, arr in G.all_tables()}
    assert got == PINNED


def test_generated_headers_are_current():
    text = G.render()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("pycricodecs_b200/csrc/cri_tables.h", "oracle/cri_tables.h"):
        assert open(os.path.join(root, rel)).read() == text, f"{rel} is stale: run tools/gen_tables.py"


def test_crc_byte_step_identity():
    """The kernels' table-free CRC step: table[v] == (v<<1) ^ (v<<2) ^ (parity(v) ? 0x8003 : 0)."""
    t = G.crc16_table()
    for v in range(256):
        assert int(t[v]) == ((v << 1) ^ (v << 2) ^ (0x8003 if bin(v).count("1") & 1 else 0)) & 0xFFFF


def test_tables_equal_reference_arrays(ref):
    T = {n: np.asarray(a).ravel() for n, _, a in G.all_tables()}
    same = [("CRC16", "crc", np.uint16), ("DEC_SCALING", "dec_scaling", np.uint32), ("DEC_RANGE", "dec_range", np.uint32),
            ("SCALE_CONV", "scale_conv", np.uint32), ("INTENSITY_RATIO", "intensity_ratio", np.uint32), ("IMDCT_SIN", "imdct_sin", np.uint32),
            ("IMDCT_COS", "imdct_cos", np.uint32), ("WINDOW", "window", np.uint32), ("INVERT", "invert", np.uint8), ("MAX_BITS", "max_bit", np.uint8),
            ("READ_BITS", "read_bit", np.uint8), ("ENC_INV_STEP", "enc_inv_step", np.uint32), ("ENC_DEAD_ZONE", "enc_dead_zone", np.uint32),
            ("ENC_RATIO_BOUNDS", "enc_ratio_bounds", np.uint32), ("ENC_Q_SCALING", "enc_q_scaling", np.uint32), ("MDCT_SIN", "mdct_sin", np.uint32),
            ("MDCT_COS", "mdct_cos", np.uint32), ("ENC_SHUFFLE", "enc_shuffle", np.uint8), ("ADX_STATIC_COEF", "adx_static", np.int16),
            ("ATH_BASE", "ath_base", np.uint8)]
    for mine, theirs, dt in same:
        assert np.array_equal(T[mine].astype(dt), ref.table(theirs, dt)), mine
    assert np.array_equal(T["READ_VALS"].astype(np.float32), ref.table("read_val", np.float32))
    assert np.array_equal(T["ENC_RES_CURVE"].astype(np.int32), ref.table("enc_res_curve", np.int32))
    assert np.array_equal(T["ENC_Q_BITS"].astype(np.int32), ref.table("enc_q_bits", np.int32))
    assert np.array_equal(T["ENC_Q_CODE"].astype(np.int8), ref.table("enc_q_value", np.int8))


def test_code_length_threshold_restatement():
    """hca_unpack_fast_kernel advances the bit position with  used = max_bits[r] - (code < short_below[r])  for both
    code families; for the prefix codebooks (r <= 7) that must restate read_bit_table (hca.cpp:1513-1526), for the
    sign-magnitude family it is the reference's 'zero gives the sign bit back' rule (hca.cpp:1554-1559)."""
    T = {n: np.asarray(a).ravel() for n, _, a in G.all_tables()}
    lo, hi = 0x26AE2620, 0x22222222
    short_below = [(lo >> (4 * r)) & 15 for r in range(8)] + [(hi >> (4 * (r - 8))) & 15 for r in range(8, 16)]
    for r in range(8):
        mb = int(T["MAX_BITS"][r])
        for code in range(1 << mb):
            assert int(T["READ_BITS"][(r << 4) | code]) == mb - (1 if code < short_below[r] else 0), (r, code)
    assert short_below[8:] == [2] * 8           # code >> 1 == 0  <=>  code < 2
    # the slice-by-4 CRC tables are the byte table iterated over zero bytes (checked bit for bit on the GPU by the
    # decode parity tests; here: the algebra)
    t0 = [int(v) for v in G.crc16_table()]

    def step(c, b):
        return ((c << 8) & 0xFFFF) ^ t0[((c >> 8) ^ b) & 0xFF]
    tk = [t0]
    for _ in range(3):
        tk.append([step(c, 0) for c in tk[-1]])
    rng = np.random.default_rng(3)
    for _ in range(200):
        c = int(rng.integers(0, 1 << 16)); bs = [int(x) for x in rng.integers(0, 256, 4)]
        want = c
        for b in bs:
            want = step(want, b)
        x = (bs[0] | bs[1] << 8 | bs[2] << 16 | bs[3] << 24) ^ (((c >> 8) & 0xFF) | ((c & 0xFF) << 8))
        assert tk[3][x & 0xFF] ^ tk[2][(x >> 8) & 0xFF] ^ tk[1][(x >> 16) & 0xFF] ^ tk[0][x >> 24] == want


def test_encoder_cost_rows_reproduce_the_table_driven_bit_count():
    """ENC_COST_ROWS (two comparisons per coefficient) against the reference's own arithmetic (quantise, look the code
    length up; hca.cpp:2772-2787) on random values, on every threshold +- a few ulps, on zero and on the clamp value."""
    import importlib.util
    import os
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(root, "tools", "gen_tables.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    rows = g.enc_cost_rows()
    rng = np.random.default_rng(3)
    clamp = np.float32(0.9999999)
    xs = [rng.uniform(-1, 1, 200000).astype(np.float32), (rng.standard_normal(100000) * 0.05).astype(np.float32),
          np.array([0.0, -0.0, clamp, -clamp], np.float32)]
    for r in range(1, 16):
        for w in rows[r][:2]:
            u = int(w) & 0x7FFFFFFF
            near = np.arange(u - 40, u + 41, dtype=np.uint32).view(np.float32)
            xs.append(near)
            xs.append(-near)
    x = np.clip(np.concatenate(xs), -clamp, clamp).astype(np.float32)
    for r in range(0, 16):
        neg_n, p = rows[r][:2].view(np.float32)
        full8, overfull = int(rows[r][2]), int(rows[r][3])
        model = full8 // 8 - ((x > neg_n) & (x < p)).astype(np.int64) - overfull * (x == clamp)
        want = g._enc_table_bits(x, r) if r else np.zeros(len(x), np.int64)
        assert np.array_equal(model, want), r


def test_encoder_rank_tables_reproduce_the_cost_rows():
    """ENC_RANK_KEYS / ENC_RANK_ROWS (one bucket look-up and one comparison per coefficient, hca_encode_kernel's counted
    bit costs) against ENC_COST_ROWS (two comparisons per coefficient and probe): for every resolution, `inside` of the
    row model == the coefficient's rank reaches past the resolution's position in the sorted order."""
    rows = G.enc_cost_rows()
    pos, keys, packed, base = G.enc_cost_ranks()
    rng = np.random.default_rng(5)
    clamp = np.float32(0.9999999)
    xs = [rng.uniform(-1, 1, 200000).astype(np.float32), (rng.standard_normal(200000) * 0.01).astype(np.float32),
          (rng.standard_normal(100000) * 1e-4).astype(np.float32), np.array([0.0, -0.0, clamp, -clamp, 1e-30, -1e-30], np.float32)]
    for r in range(1, 16):
        for w in rows[r][:2]:
            u = int(w) & 0x7FFFFFFF
            near = np.arange(u - 40, u + 41, dtype=np.uint32).view(np.float32)
            xs += [near, -near]
    x = np.clip(np.concatenate(xs), -clamp, clamp).astype(np.float32)
    bits = x.view(np.uint32)
    a = bits & 0x7FFFFFFF
    sign = (bits >> 31).astype(np.int64)
    q = np.maximum(a >> G.ENC_RANK_SHIFT, base).astype(np.int64) - base
    assert q.max() == G.ENC_RANK_BUCKETS - 1
    entry = keys[sign, q]
    rank = entry[:, 1].astype(np.int64) + (a < entry[:, 0])
    for r in range(0, 16):
        neg_n, p = rows[r][:2].view(np.float32)
        inside = (x > neg_n) & (x < p)
        assert np.array_equal(inside, rank > int(pos[r])), r
        assert int(packed[r]) == 4 * int(pos[r]) | int(rows[r][2]) << 8 | int(rows[r][3]) << 16


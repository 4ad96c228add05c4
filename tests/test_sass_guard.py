"""Build guard for fp32 parity: the decode transform kernels must round every product and every sum separately, so their
machine code may contain no fused multiply-add except the two-wide sums written as fma(q, {1.0f, 1.0f}, p) (hca_sum2:
832 per kernel = 2 transforms x 384 rotation sums + 64 window sums) -- a contraction by the assembler would show up
as a scalar FFMA, a packed multiply, or a different FFMA2 count. Reads the built library with cuobjdump (no GPU)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pycricodecs_b200", "libcricodecs_b200.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_transform_kernels_have_no_contracted_arithmetic():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    kernels = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)
    fast = {k: v for k, v in kernels.items() if "hca_imdct_fast_kernel" in k}
    assert len(fast) == 4                                   # mono / stereo x plain / joint
    for k, lines in fast.items():
        ops = [m.group(1) for m in (re.search(r"\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln) for ln in lines) if m]
        count = lambda op: sum(1 for o in ops if o == op)
        assert count("FFMA") == 0 and count("FMUL2") == 0, k
        assert count("FFMA2") == 832, (k, count("FFMA2"))
        assert count("FADD2") == 768, (k, count("FADD2"))
        assert count("FMUL") >= 3712, (k, count("FMUL"))    # 2 x 1792 rotation products + 256 window products, a few folded
    # the general-path transform is scalar code: no fused multiply-add of any width (the encoder is not checked this
    # way: its correctly rounded divisions, __fdiv_rn, expand to FFMA sequences)
    for k, lines in kernels.items():
        if "hca_imdct_kernel" in k:
            text = "\n".join(lines)
            assert " FFMA" not in text and "FFMA2" not in text, k

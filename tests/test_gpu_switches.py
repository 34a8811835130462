"""The alternative code paths behind the YSI_* switches must agree with the default path on the benchmarked
model: the implicit-GEMM neck vs the im2col GEMM, the folded LayerNorm (none / LayerNorm1 / both) and the decoder's fused
out-projection + LayerNorm4 vs the separate kernels.  The library reads the switches once per process, so every configuration runs
tests/switch_probe.py in its own process."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _probe(tmp_path, name, env):
    out = os.path.join(str(tmp_path), name + ".npz")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(HERE, "switch_probe.py"), out], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def test_switched_paths_agree_with_the_default(tmp_path):
    ref = _probe(tmp_path, "default", {})
    # implicit-GEMM 3x3 convolution vs im2col + GEMM: same tiles, same k order -> the same bits
    a = _probe(tmp_path, "neck_im2col", {"YSI_NECK_IMPLICIT": "0"})
    assert np.array_equal(a["emb"], ref["emb"]), rel_l2(a["emb"], ref["emb"])
    # LayerNorm: separate kernels / LayerNorm1 folded (default) / both folded. Different rounding points of the 16-bit operand,
    # same arithmetic otherwise: well inside the operand-rounding noise of the encoder (~4e-4 vs the fp32 oracle)
    e0 = _probe(tmp_path, "ln_separate", {"YSI_LN_FUSED": "0"})
    e3 = _probe(tmp_path, "ln_both", {"YSI_LN_FUSED": "3"})
    d0, d3 = rel_l2(e0["emb"], ref["emb"]), rel_l2(e3["emb"], ref["emb"])
    print("embeddings rel-L2 vs the default path: separate LayerNorm kernels %.2e, both LayerNorms folded %.2e" % (d0, d3))
    assert 0.0 < d0 < 1e-3 and 0.0 < d3 < 1e-3
    # decoder: fused out-projection + LayerNorm4 vs GEMM + keys_ln_kernel (fp32 arithmetic, different summation order)
    k = _probe(tmp_path, "dec_separate", {"YSI_DEC_FUSED_LN": "0"})
    assert np.array_equal(k["emb"], ref["emb"])
    dl = rel_l2(k["low"], ref["low"])
    print("low-res logits rel-L2, separate LayerNorm4 kernel vs fused epilogue: %.2e" % dl)
    assert dl < 1e-5

"""Golden vectors from the reference's OWN known-answer machinery for the hot path's data formats: the Python
`gguf.quants.dequantize` of /root/reference/tools/convert_hf_to_gguf/gguf-py, which the reference's
gguf-py/tests/test_quants.py:187-262 pins bit-exact against libggml's dequantize_row_* (SURVEY section 4).  The Python
reference cannot travel to the GPU box, so its outputs on seeded blocks are committed here.
Run in the build container:  python tests/golden/make_golden_gguf_py.py   ->  tests/golden/gguf_py.npz
Inputs are regenerated from the seeds by the test (powerserve_b200.synth.random_blocks is deterministic)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from powerserve_b200 import synth  # noqa: E402
from tests.golden.cases import GGUF_PY_CASES  # noqa: E402

sys.path.append("/root/reference/tools/convert_hf_to_gguf/gguf-py")  # after ours: it has a `tests` package of its own
import gguf  # noqa: E402  (the reference's package)
from gguf import quants  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for name, ggml_type, blk, rows, n_blocks, seed, scale in GGUF_PY_CASES:
        w = synth.random_blocks(np.random.default_rng(seed), ggml_type, rows, blk * n_blocks, scale)
        d = quants.dequantize(w, getattr(gguf.GGMLQuantizationType, name))
        assert d.dtype == np.float32 and d.shape == (rows, blk * n_blocks)
        out[f"{name}/{seed}"] = np.ascontiguousarray(d).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "gguf_py.npz"), **out)
    print("wrote", os.path.join(HERE, "gguf_py.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

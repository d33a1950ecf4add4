"""Generate the committed golden vectors from the COMPILED REFERENCE (oracle/_ref, built from /root/reference by
`make -C oracle ref`).  Run in the build container:  python tests/golden/make_golden.py

ops.npz     op-level known answers of the reference's own kernels on seeded inputs (bit patterns stored as uint32):
            quantize_row_q8_K / q8_0, powerserve_compute_forward_mul_mat for Q4_K/Q6_K/Q4_0/Q8_0 (bs 1 and 3),
            rms_norm, rope (NORM and NEOX), softmax_ext, silu_hadamard, get_embedding.
models.npz  logits + greedy ids of the reference's model stack (ps_ref_run) on the synthetic tiny models.
Inputs are regenerated from the seeds by the tests (powerserve_b200.synth is deterministic), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from powerserve_b200 import synth  # noqa: E402
from tests import _libs as L  # noqa: E402
from tests import _model as M  # noqa: E402
from tests.golden import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert L.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, fn in cases.OP_CASES.items():
        res = fn(cases.RefBackend())
        for k, v in res.items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    mout = {}
    for preset, n_prompt, batch, n_dec in cases.MODEL_CASES:
        d = M.model_dir(preset)
        prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=7 + n_prompt)
        ids, lg, _ = M.run_reference(d, prompt, n_dec, batch_size=batch, n_threads=3, dump_logits=n_dec)
        key = f"{preset}/{n_prompt}/{batch}"
        mout[key + "/ids"] = np.asarray(ids, np.int32)
        mout[key + "/logits_bits"] = L.bits(lg)
    np.savez_compressed(os.path.join(HERE, "models.npz"), **mout)
    print("wrote", len(out), "op arrays and", len(mout), "model arrays")


if __name__ == "__main__":
    main()

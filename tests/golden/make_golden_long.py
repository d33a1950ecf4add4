"""Golden vectors of the COMPILED REFERENCE (oracle/_ref/ps_ref_run) at the benchmarked context length and beyond:
prompt >= 2048 tokens (prefill in chunks of 128) followed by greedy decode steps, on the real per-layer shapes of the
BASELINE models (`slice-*-long`).  Run in the build container:  python tests/golden/make_golden_long.py
Stored: greedy ids + logits bit patterns per case (tests/golden/long_ctx.npz); inputs are regenerated from seeds."""
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


def long_prompt(preset, n_prompt):
    return synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=100 + n_prompt)


def main():
    assert L.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for preset, n_prompt, batch, n_dec in cases.LONG_CASES:
        d = M.model_dir(preset)
        ids, lg, tm = M.run_reference(d, long_prompt(preset, n_prompt), n_dec, batch_size=batch, n_threads=7, dump_logits=n_dec)
        key = f"{preset}/{n_prompt}/{batch}"
        out[key + "/ids"] = np.asarray(ids, np.int32)
        out[key + "/logits_bits"] = L.bits(lg)
        print(key, ids, tm)
    np.savez_compressed(os.path.join(HERE, "long_ctx.npz"), **out)


if __name__ == "__main__":
    main()

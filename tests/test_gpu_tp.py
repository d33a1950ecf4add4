"""Tensor parallelism on real GPUs: N processes, one per GPU, row-sharded weights, NCCL all-gathers
inside the decode graph (mode nccl) or fused into the producing kernels as NVLink peer stores + epoch flags (mode p2p).
Row sharding keeps every dot product whole, so the result must be BIT-IDENTICAL to one GPU —
checked against the oracle.  Needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M

pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        return capi.load_library().ps_cuda_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["p2p", "p2p_fence", "nccl"])
@pytest.mark.parametrize("preset,size", [("tiny-llama", 2), ("slice-1b", 2), ("slice-1b", 4)])
def test_tp_decode_bit_exact(preset, size, mode):
    if n_gpus() < size:
        pytest.skip(f"needs {size} GPUs")
    d = M.model_dir(preset)
    n_prompt, n_dec = 19, 12
    with tempfile.TemporaryDirectory() as td:
        procs = [subprocess.Popen([sys.executable, os.path.join(L.ROOT, "tools", "tp_worker.py"), str(r), str(size), os.path.join(td, "id"), d,
                                   str(n_prompt), str(n_dec), os.path.join(td, "out"), mode], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for r in range(size)]
        outs = [p.communicate(timeout=300)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        res = [np.load(os.path.join(td, f"out.rank{r}.npz")) for r in range(size)]
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=11)
    # a tensor-parallel batch (prefill chunk of 16, verify-shaped batch of 6) runs as batched row-sharded GEMMs with one
    # all-gather per exchange: it must equal the reference run with the SAME chunking (the reference's results depend on
    # it: ggml_vec_soft_max_f32 uses its SIMD exp for full 8-groups of a row and libm expf for the tail, and the row length
    # is the chunk's last position + 1)
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, 4, batch_size=16)
    ids_long, _ = om.generate(prompt, n_dec, batch_size=16)
    om.reset()
    rows_o = om.forward(prompt[:6])
    om.close()
    for r in res:
        assert list(r["ids"]) == ids_o
        L.assert_bit_equal(r["logits"], lg_o, "tensor-parallel logits vs oracle")
        assert list(r["dev_ids"]) == ids_long
        L.assert_bit_equal(r["batch_logits"], rows_o, "tensor-parallel batch with lm_head (verify shape)")
        L.assert_bit_equal(r["batch_dev"], rows_o, "ps_cuda_logits_dev after a tensor-parallel batch")
        assert int(r["tp_error"]) == 0
        assert int(r["p2p"]) == (1 if mode.startswith("p2p") else 0)
        if mode == "nccl":
            assert int(r["gathers"]) > 0

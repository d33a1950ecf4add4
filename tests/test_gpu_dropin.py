"""The drop-in, end to end: PowerServe's OWN C++ stack (LlamaModel/Qwen2Model::forward -> Graph -> Executor ->
Platform) built with the CUDABackend of powerserve_b200/host plugged in must produce the same logits
and greedy ids as the reference's CPU path — checked against the committed golden vectors of the compiled reference and,
for the fused decode path, against the oracle.  Needs powerserve_b200/host/_build/ps_cuda_run (built where
/root/reference exists; the binary travels with gpurun)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from powerserve_b200 import synth
from tests import _libs as L
from tests import _model as M
from tests.golden import cases

pytestmark = pytest.mark.gpu
EXE = os.path.join(L.ROOT, "powerserve_b200", "host", "_build", "ps_cuda_run")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="ps_cuda_run not built (needs /root/reference)")


def run_dropin(path, prompt, n_decode, batch_size, dump_logits, per_op=False, device_topk=0):
    env = dict(os.environ)
    env["POWERSERVE_CUDA_PER_OP"] = "1" if per_op else "0"
    with tempfile.TemporaryDirectory() as td:
        pf = os.path.join(td, "prompt.txt")
        open(pf, "w").write(" ".join(str(int(t)) for t in prompt))
        extra = ["--device-topk", str(device_topk)] if device_topk else ["--dump-logits", str(dump_logits)]
        r = subprocess.run([EXE, path, "2", str(batch_size), pf, str(n_decode), os.path.join(td, "out")] + extra, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        ids = [int(x) for x in open(os.path.join(td, "out.ids")).read().split()]
        if device_topk:
            return ids, None
        vocab = json.load(open(os.path.join(path, "model.json")))["llm_config"]["vocab_size"]
        return ids, np.fromfile(os.path.join(td, "out.logits"), dtype=np.float32).reshape(-1, vocab)


@needs_exe
@pytest.mark.parametrize("preset,n_prompt,batch,n_dec", cases.MODEL_CASES)
def test_powerserve_stack_on_cuda_matches_reference_golden(preset, n_prompt, batch, n_dec):
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "models.npz"))
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=7 + n_prompt)
    ids, lg = run_dropin(M.model_dir(preset), prompt, n_dec, batch, n_dec)
    key = f"{preset}/{n_prompt}/{batch}"
    assert ids == list(gold[key + "/ids"])
    assert (L.bits(lg) == gold[key + "/logits_bits"]).all()


@needs_exe
def test_powerserve_stack_on_cuda_fused_decode_matches_oracle():
    d = M.model_dir("slice-1b")
    prompt = synth.random_prompt(synth.PRESETS["slice-1b"].vocab_size, 23, seed=5)
    ids, lg = run_dropin(d, prompt, 6, 16, 6)
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, 6, batch_size=16)
    om.close()
    L.assert_bit_equal(lg, lg_o, "drop-in logits vs oracle")
    assert ids == ids_o


@needs_exe
@pytest.mark.parametrize("preset,n_prompt,batch,n_dec", cases.MODEL_CASES)
def test_powerserve_unfused_graph_op_by_op_on_cuda_matches_reference_golden(preset, n_prompt, batch, n_dec):
    """POWERSERVE_CUDA_PER_OP=1: Executor::allocate_buffers / Executor::run dispatch on the CUDA backend op by op - the
    reference's own unfused graph (VIEW / PERMUTE / TRANSPOSE / CONT / COPY into cache views / the two fp32 attention
    mat-muls over strided views / GET_MASK / SOFTMAX_EXT ...) with every intermediate in a CUDABuffer.  Same golden vectors."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "models.npz"))
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=7 + n_prompt)
    ids, lg = run_dropin(M.model_dir(preset), prompt, n_dec, batch, n_dec, per_op=True)
    key = f"{preset}/{n_prompt}/{batch}"
    assert ids == list(gold[key + "/ids"])
    assert (L.bits(lg) == gold[key + "/logits_bits"]).all()


@needs_exe
@pytest.mark.parametrize("preset,n_prompt,batch,n_dec", cases.MODEL_CASES[:3])
def test_powerserve_stack_with_device_side_topk(preset, n_prompt, batch, n_dec):
    """the stack with lazy logits - CUDA_FORWARD leaves the logits on the device, CUDABackend::topk runs TopKSampler
    there and the host picks from 40 (logit, token) pairs; the greedy ids must be the golden ones."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "models.npz"))
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=7 + n_prompt)
    ids, _ = run_dropin(M.model_dir(preset), prompt, n_dec, batch, 0, device_topk=40)
    assert ids == list(gold[f"{preset}/{n_prompt}/{batch}/ids"])

"""Pin the oracle (oracle/ps_oracle.c) against the COMMITTED golden vectors (tests/golden/*.npz), which
tests/golden/make_golden.py generated from the compiled reference.  Runs anywhere gcc is (no /root/reference needed)."""
import os

import numpy as np
import pytest

from powerserve_b200 import synth
from tests import _libs as L
from tests import _model as M
from tests.golden import cases

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", list(cases.OP_CASES))
def test_oracle_ops_match_golden(case):
    gold = np.load(os.path.join(G, "ops.npz"))
    res = cases.OP_CASES[case](cases.OracleBackend())
    assert res, case
    for k, v in res.items():
        g = gold[f"{case}/{k}"]
        assert g.shape == v.shape and (g == v).all(), f"{case}/{k}: oracle differs from the reference's golden output"


@pytest.mark.parametrize("preset,n_prompt,batch,n_dec", cases.MODEL_CASES)
def test_oracle_models_match_golden(preset, n_prompt, batch, n_dec):
    gold = np.load(os.path.join(G, "models.npz"))
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=7 + n_prompt)
    om = M.OracleModel(M.model_dir(preset))
    ids, lg = om.generate(prompt, n_dec, batch_size=batch)
    om.close()
    key = f"{preset}/{n_prompt}/{batch}"
    assert ids == list(gold[key + "/ids"])
    assert (L.bits(lg) == gold[key + "/logits_bits"]).all()

"""Debug: tcgen05 prefill GEMM vs the multi-column row-walker vs the oracle on small models."""
import sys; sys.path.insert(0, ".")
import numpy as np
from powerserve_b200 import capi, synth
from tests import _libs as L, _model as M
for preset in sys.argv[1:] or ["tiny-llama", "slice-1b"]:
    d = M.model_dir(preset)
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, 49, seed=3)
    om = M.OracleModel(d); ids_o, lg_o = om.generate(prompt, 2, batch_size=64); om.close()
    cm = capi.CudaModel(d, max_batch=64)
    print(preset, "tc_ok", cm.be.counter("tc_ok"))
    for tc in (0, 1):
        cm.be.set_option("tc", tc)
        ids, lg = cm.generate(prompt, 2, batch_size=64)
        print(f"  tc={tc}: ids_equal={ids == ids_o} logit mismatches={int((L.bits(lg) != L.bits(lg_o)).sum())} max|d|={np.abs(lg - lg_o).max():.3e} "
              f"gemm launches={cm.be.counter('tc_gemm_launches')} tc_error={cm.be.counter('tc_error')}")
    cm.close()

"""Debug: the persistent step kernel vs the per-phase fused path on one preset (run under compute-sanitizer if needed)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M
preset = sys.argv[1] if len(sys.argv) > 1 else "tiny-bigvocab"
n_prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = M.model_dir(preset)
shape = synth.PRESETS[preset]
prompt = synth.random_prompt(shape.vocab_size, n_prompt, seed=11)
cm = capi.CudaModel(d, max_batch=16)
cm.be.set_option("persist", 0)
ids0, lg0 = cm.generate(prompt, 6, batch_size=16)
cm.be.set_option("persist", 1)
ids1, lg1 = cm.generate(prompt, 6, batch_size=16)
print(preset, "ids", ids0, ids1, "step_error", cm.be.counter("step_error"))
L.assert_bit_equal(lg1, lg0, "persistent vs per-phase")
cm.reset(); cm.prefill(prompt, 16)
print("device loop", list(cm.decode_greedy(int(prompt[-1]), 6)))
cm.close()
print("ok")

import sys
import numpy as np
sys.path.insert(0, ".")
from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M
preset = sys.argv[1] if len(sys.argv) > 1 else "tiny-deep"
d = M.model_dir(preset)
shape = synth.PRESETS[preset]
prompt = synth.random_prompt(shape.vocab_size, 45, seed=21)
om = M.OracleModel(d)
om.reset(); om.forward(prompt[:33], lm_head=False)
lo_batch = om.forward(prompt[33:45])
om.reset(); om.forward(prompt[:33], lm_head=False)
lo_single = np.stack([om.forward([int(t)])[0] for t in prompt[33:45]])
om.close()
print("oracle batch vs single", np.abs(lo_batch - lo_single).max())
for fused in (0, 1):
    a, b = capi.CudaModel(d, max_batch=64), capi.CudaModel(d, max_batch=64)
    a.be.set_option("fused", fused); b.be.set_option("fused", fused)
    a.prefill(prompt[:34], 33); b.prefill(prompt[:34], 33)
    la = a.forward(prompt[33:45])
    lb = np.stack([b.forward([int(t)])[0] for t in prompt[33:45]])
    print(preset, "fused", fused, "batch vs oracle batch", np.abs(la - lo_batch).max(), "| single vs oracle single", np.abs(lb - lo_single).max())
    a.close(); b.close()

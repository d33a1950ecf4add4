"""Debug: per-CTA timeline of the fused mat-vec kernels of one decode step (globaltimer stamps)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b"
shape = synth.PRESETS[model]
if len(sys.argv) > 2:
    shape.n_layers = int(sys.argv[2])
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=16, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
m.prefill(synth.random_prompt(shape.vocab_size, 33), 16)
m.decode_greedy(1, 4)
for pdl, graph in [(0, 0), (1, 1)]:
    m.be.set_option("graph", graph); m.be.set_option("pdl", pdl)
    m.decode_greedy(1, 2)
    m.be.set_option("trace", 1)
    m.decode_greedy(1, 1)
    nl = 4 * shape.n_layers + 1
    buf = np.zeros((nl, 148, 16), np.int64)
    m.be._ck(m.be.L.ps_cuda_read_trace(m.be.h, buf.ctypes.data, nl))
    print(f"=== pdl={pdl} graph={graph}: kernel# kind | start spread | per-CTA medians (us): init, depwait, prologue, t0wait, tile0, t1wait, ... total | kernel span")
    t00 = buf[:, :, 0][buf[:, :, 0] > 0].min()
    names = ["QKV", "WO", "GATEUP", "DOWN"]
    for k in range(min(nl, 9)):
        b = buf[k]
        ok = b[:, 0] > 0
        if not ok.any():
            continue
        b = b[ok].astype(np.float64)
        st = b[:, 0]
        seg = lambda i, j: np.median((b[:, j] - b[:, i])[(b[:, j] > 0) & (b[:, i] > 0)]) / 1e3 if ((b[:, j] > 0) & (b[:, i] > 0)).any() else float("nan")
        span = (b[:, 13].max() - st.min()) / 1e3
        print(f"{k:2d} {names[k % 4] if k < nl - 1 else 'LMHEAD':6s} start+{(st.min() - t00) / 1e3:8.2f} spread {(st.max() - st.min()) / 1e3:5.2f} | init {seg(0, 1):5.2f} dep {seg(1, 2):5.2f} pro {seg(2, 3):5.2f} "
              f"w0 {seg(4, 5):5.2f} t0 {seg(5, 6):5.2f} w1 {seg(6, 7):5.2f} t1 {seg(7, 8):5.2f} w2 {seg(8, 9):5.2f} t2 {seg(9, 10):5.2f} | cta total {seg(0, 13):6.2f} | span {span:6.2f}")
    m.be.set_option("trace", 0)

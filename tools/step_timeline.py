"""Device timeline of ONE persistent decode step (ps_k_step, option "trace"): per phase, the first CTA's entry, the last
CTA's exit, when the last CTA had its inputs (the (value, epoch) words of the previous phase / the barrier) and its
quantised image, from %globaltimer stamps.

    python tools/step_timeline.py [model] [n_layers] [ctx] [--opt=name=value ...]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
model = args[0] if len(args) > 0 else "llama-3.1-8b"
shape = synth.PRESETS[model]
if len(args) > 1:
    shape.n_layers = int(args[1])
ctx_len = int(args[2]) if len(args) > 2 else 64
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
m.prefill(synth.random_prompt(shape.vocab_size, ctx_len + 1), 128)
for a in sys.argv[1:]:
    if a.startswith("--opt="):
        k, v = a[6:].split("=")
        m.be.set_option(k, int(v))
m.decode_greedy(1, 4)
m.decode_greedy(1, 32)
ms = m.be.counter("last_device_ns") / 1e6 / 32
m.be.set_option("trace", 1)
m.decode_greedy(1, 1)
n = 1 + 6 * shape.n_layers + 1
buf = np.zeros((n, 8), np.int64)
m.be._ck(m.be.L.ps_cuda_read_trace(m.be.h, buf.ctypes.data, n))
m.be.set_option("trace", 0)
names = ["EMBED"] + ["QKV", "SCORES", "PV", "WO", "GATEUP", "DOWN"] * shape.n_layers + ["LMHEAD"]
t0 = buf[0, 0]
print(f"{model} layers={shape.n_layers} ctx={ctx_len}: {ms * 1e3:.1f} us/step untraced ({1e3 / ms:.1f} tok/s); traced step below (us)")
print(f"{'#':>3s} {'phase':7s} {'first in':>9s} {'inputs':>8s} {'image':>8s} {'last out':>9s} {'span':>7s} {'excl':>7s}")
prev_end = t0
tot = {}
for k in range(n):
    s, e, _, inp, img = buf[k][:5]
    f = lambda v: (v - t0) / 1e3
    excl = (e - max(prev_end, t0)) / 1e3
    if k < 14 or k >= n - 2:
        print(f"{k:3d} {names[k]:7s} {f(s):9.2f} {(f(inp) if inp else float('nan')):8.2f} {(f(img) if img else float('nan')):8.2f} {f(e):9.2f} {(e - s) / 1e3:7.2f} {excl:7.2f}")
    tot.setdefault(names[k], []).append((excl, (inp - prev_end) / 1e3 if inp else 0.0, (img - inp) / 1e3 if inp else 0.0, (e - img) / 1e3 if img else 0.0,
                                         buf[k][5] / 1e3, buf[k][6] / 1e3, buf[k][3] / 1e3, buf[k][4] / 1e3, buf[k][7] / 1e3))
    prev_end = e
print("per phase kind: mean exclusive time (last exit - previous phase's last exit) | wait for inputs | build image | walk")
for nm, v in tot.items():
    a = np.array(v)
    if nm == "PV":
        print(f"  {nm:7s} n={len(v):3d} excl {a[:, 0].mean():7.2f} us  (slowest CTA, from entry: max known {a[:, 6].mean():6.2f}  sum known {a[:, 7].mean():6.2f}  P.V done {a[:, 4].mean():6.2f} us; exp loop 1st trip {a[:, 5].mean():.2f} kcyc, 2nd trip {a[:, 8].mean():.2f} kcyc)   sum {a[:, 0].sum():8.1f} us")
    else:
        print(f"  {nm:7s} n={len(v):3d} excl {a[:, 0].mean():7.2f} us  inputs {a[:, 1].mean():6.2f}  image {a[:, 2].mean():6.2f}  walk {a[:, 3].mean():6.2f}  [slowest warp: walk {a[:, 4].mean():6.2f} kcyc, of which waiting for weights {a[:, 5].mean():6.2f}, epilogues {a[:, 8].mean():6.2f} kcyc]   sum {a[:, 0].sum():8.1f} us")
print(f"step span {(buf[n - 1, 1] - t0) / 1e3:.1f} us")
m.close()

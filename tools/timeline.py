"""Debug: device timeline of ONE fused decode step (globaltimer stamps written by every kernel of the step, option
"trace"): per kernel start / dependency-resolved / prologue-done / end, and the gap to the previous kernel's end.

    python tools/timeline.py [model] [n_layers] [ctx] [--nopdl] [--nograph]
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
m.be.set_option("pdl", 0 if "--nopdl" in sys.argv else 1)
m.be.set_option("graph", 0 if "--nograph" in sys.argv else 1)
for a in sys.argv[1:]:
    if a.startswith("--opt="):      # e.g. --opt=ksplit=0
        k, v = a[6:].split("=")
        m.be.set_option(k, int(v))
m.decode_greedy(1, 4)          # graph capture + warm-up
m.decode_greedy(1, 16)
ms = m.be.counter("last_device_ns") / 1e6 / 16
m.be.set_option("trace", 1)
m.decode_greedy(1, 1)
n = 1 + 6 * shape.n_layers + 2
buf = np.zeros((n, 8), np.int64)
m.be._ck(m.be.L.ps_cuda_read_trace(m.be.h, buf.ctypes.data, n))
big = np.zeros((512, 8), np.int64)
m.be._ck(m.be.L.ps_cuda_read_trace(m.be.h, big.ctypes.data, 512))
m.be.set_option("trace", 0)
names = ["EMBED"] + ["QKV", "ATTN1", "ATTN2", "WO", "GATEUP", "DOWN"] * shape.n_layers + ["LMHEAD", "ARGMAX"]
t0 = buf[0, 0]
print(f"{model} layers={shape.n_layers} ctx={ctx_len}: {ms * 1e3:.1f} us/step untraced; traced step below (us)")
print(f"{'#':>3s} {'kernel':7s} {'start':>8s} {'dep_ok':>7s} {'pro_ok':>7s} {'end':>8s} {'dur':>6s} {'gap':>6s}")
prev_end = t0
tot = {}
for k in range(n):
    s, e, d, p = buf[k][:4]
    extra = " ".join(f"{v / 1e3:5.2f}" for v in buf[k][4:8]) if buf[k][4:8].any() else ""
    f = lambda v: (v - t0) / 1e3
    dep = f"{f(d):7.2f}" if d < 2**62 else "      -"
    pro = (f"{f(p):7.2f}" if p > 10**9 else f"{p / 1e3:6.2f}d") if p > 0 else "      -"
    if k < 14 or k >= n - 3:
        print(f"{k:3d} {names[k]:7s} {f(s):8.2f} {dep} {pro} {f(e):8.2f} {(e - s) / 1e3:6.2f} {(s - prev_end) / 1e3:6.2f}  {extra}")
    tot.setdefault(names[k], []).append(((e - max(s, prev_end)) / 1e3, (e - prev_end) / 1e3))
    prev_end = e
print("per-kernel-kind mean exclusive time (end - max(start, prev end)) and step share (end - prev end):")
for nm, v in tot.items():
    a = np.array(v)
    print(f"  {nm:7s} n={len(v):3d} excl {a[:, 0].mean():7.2f} us  share-of-step {a[:, 1].sum():8.1f} us")
if any(a.startswith("--cta") for a in sys.argv):   # per-CTA stream trace of the last launch of the kind chosen with --opt=cta_trace=K (default 1 = Wdown)
    c = big[256:256 + 148]
    dep0 = c[:, 0].min()
    print("per-CTA trace: dep(us, rel) begin-dep end-dep loop_kcyc wait_kcyc blocks octets")
    order = np.argsort(c[:, 2])
    for i in list(order[:6]) + list(order[-12:]):
        print(f"  cta {i:3d} dep {(c[i,0]-dep0)/1e3:6.2f} begin {(c[i,1]-c[i,0])/1e3:6.2f} end {(c[i,2]-c[i,0])/1e3:6.2f} loop {c[i,3]/1e3:7.2f} wait {c[i,4]/1e3:6.2f} blocks {c[i,5]} oct {c[i,6]}")
    print("  mean end-dep", (c[:, 2] - c[:, 0]).mean() / 1e3, "max", (c[:, 2] - c[:, 0]).max() / 1e3, "mean loop kcyc", c[:, 3].mean() / 1e3, "mean wait kcyc", c[:, 4].mean() / 1e3,
          "cyc/block (excl wait)", ((c[:, 3] - c[:, 4]) / np.maximum(c[:, 5], 1)).mean(), "GHz", (c[:, 3] / np.maximum(c[:, 2] - c[:, 1], 1)).mean())
print(f"step span {(buf[n - 1, 1] - t0) / 1e3:.1f} us")
m.close()

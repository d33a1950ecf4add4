"""A/B of decode-path options in ONE process: the same 8B-shaped slice, the same context, each option set timed over the
same decode steps (device time of ps_cuda_decode_greedy) and checked to produce the same token ids and logits bits.

    python tools/ab_decode.py [model] [n_layers] [ctx] "name=v,name=v" "name=v" ...     ("" = defaults)
    --timeline: also print the per-kernel-kind exclusive times of one traced step per option set
"""
import hashlib
import sys

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

pos_args = [a for a in sys.argv[1:] if not a.startswith("--")]
model, n_layers, ctx_len = pos_args[0], int(pos_args[1]), int(pos_args[2])
sets = pos_args[3:] or [""]
shape = synth.PRESETS[model]
shape.n_layers = n_layers
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
m.prefill(synth.random_prompt(shape.vocab_size, ctx_len + 1), 128)
base_pos = m.position
names = ["EMBED"] + ["QKV", "ATTN1", "ATTN2", "WO", "GATEUP", "DOWN"] * n_layers + ["LMHEAD", "ARGMAX"]
defaults = {}
ref = None
for spec in sets:
    opts = dict(kv.split("=") for kv in spec.split(",") if kv)
    for k in defaults:
        m.be.set_option(k, defaults[k])
    for k, v in opts.items():
        defaults.setdefault(k, {"rw_ksplit": 0, "rw_defer": 0, "rw_unroll2": 1, "pdl": 1, "graph": 1, "rw_kb": 0, "ops_graph": 1, "attn_group": 0, "mv_kpar": 1}.get(k, 0))
        m.be.set_option(k, int(v))
    m.be.kv_truncate(base_pos)
    ids = m.decode_greedy(1, 8)           # capture + warm-up; ids and logits are the parity sample
    logits = m.be.read_device(m.be.logits_dev(), shape.vocab_size)
    sig = (tuple(int(i) for i in ids), hashlib.sha256(logits.tobytes()).hexdigest()[:16])
    if ref is None:
        ref = sig
    best = 1e9
    for _ in range(3):
        m.be.kv_truncate(base_pos)
        m.decode_greedy(1, 32)
        best = min(best, m.be.counter("last_device_ns") / 1e3 / 32)
    line = f"[{spec or 'defaults':40s}] {best:8.1f} us/step  same_as_first={sig == ref} launch_err={m.be.counter('tc_error')}"
    if "--timeline" in sys.argv:
        m.be.kv_truncate(base_pos)
        m.be.set_option("trace", 1)
        m.decode_greedy(1, 1)
        n = len(names)
        buf = np.zeros((n, 8), np.int64)
        m.be._ck(m.be.L.ps_cuda_read_trace(m.be.h, buf.ctypes.data, n))
        m.be.set_option("trace", 0)
        tot, pro, prev_end = {}, {}, buf[0, 0]
        for k in range(n):
            s, e = buf[k][:2]
            tot.setdefault(names[k], []).append((e - max(s, prev_end)) / 1e3)
            # dependency resolved (first CTA) relative to the kernel's start, then the slowest CTA's prologue probes relative to ITS dependency
            pro.setdefault(names[k], []).append([(buf[k][2] - s) / 1e3 if buf[k][2] < 2**62 else 0.0] + [v / 1e3 for v in buf[k][4:7]] + [buf[k][3] / 1e3 if 0 < buf[k][3] < 10**9 else 0.0])
            prev_end = e
        line += "  | " + " ".join(f"{nm} {np.mean(v):.2f}" for nm, v in tot.items())
        line += "\n      prologue (dep-start, ss, scaled, quantised, prologue end; us): " + " ".join(
            f"{nm} " + "/".join(f"{x:.2f}" for x in np.mean(np.array(v), axis=0)) for nm, v in pro.items() if nm in ("QKV", "WO", "GATEUP", "DOWN", "LMHEAD"))
    print(line, flush=True)
m.close()

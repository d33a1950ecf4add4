"""Per-kernel device timeline of ONE tensor-parallel fused decode step (rank 0's view), N ranks on N GPUs of one box.

    python tools/tp_timeline.py [size] [n_layers] [ctx] [p2p|nccl]          # parent: spawns the ranks
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "--rank":
    from powerserve_b200 import capi, gguf, synth
    rank, size, n_layers, ctx, mode, td = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6], sys.argv[7]

    def put(name, data):
        with open(os.path.join(td, name + ".tmp"), "wb") as f:
            f.write(data)
        os.replace(os.path.join(td, name + ".tmp"), os.path.join(td, name))

    def get(name):
        t0 = time.time()
        while not os.path.exists(os.path.join(td, name)):
            if time.time() - t0 > 300:
                raise SystemExit(f"no {name} after 300 s")
            time.sleep(0.05)
        return open(os.path.join(td, name), "rb").read()

    if rank == 0:
        put("id", capi.tp_unique_id())
    nid = get("id")
    shape = synth.PRESETS["llama-3.1-8b"]
    shape.n_layers, shape.vocab_size, shape.n_ctx = n_layers, 4096, 4096
    tensors = synth.generate_tensors(shape, 0)
    tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
    desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096, tp_rank=rank, tp_size=size)
    m = capi.CudaModel(desc=desc, tensors=tmap, device=rank, nccl_id=nid)
    if mode == "p2p":
        put(f"ipc{rank}", m.tp_export())
        m.tp_import([get(f"ipc{r}") for r in range(size)])
    m.prefill(synth.random_prompt(shape.vocab_size, ctx + 1), 128)
    m.decode_greedy(1, 4)
    m.decode_greedy(1, 16)
    ms = m.be.counter("last_device_ns") / 1e6 / 16
    m.be.set_option("trace", 1)
    m.decode_greedy(1, 1)
    n = 512
    buf = np.zeros((n, 8), np.int64)
    m.be._ck(m.be.L.ps_cuda_read_trace(m.be.h, buf.ctypes.data, n))
    m.be.set_option("trace", 0)
    if rank == 0:
        used = [k for k in range(256) if buf[k, 1] > 0]
        t0 = buf[0, 0]
        print(f"tp{size} {mode} layers={n_layers} ctx={ctx}: {ms * 1e3:.1f} us/step untraced; traced launches: {len(used)}")
        print(f"{'#':>3s} {'start':>8s} {'dep_ok':>7s} {'pro_ok':>7s} {'end':>8s} {'dur':>6s} {'excl':>6s}  probes")
        prev = t0
        for k in used:
            s, e, d, p = buf[k][:4]
            f = lambda v: (v - t0) / 1e3
            dep = f"{f(d):7.2f}" if 0 < d < 2**62 else "      -"
            pro = (f"{f(p):7.2f}" if p > 10**9 else f"{p / 1e3:6.2f}d") if p > 0 else "      -"
            extra = " ".join(f"{v / 1e3:5.2f}" for v in buf[k][4:8])
            if k < 16 or k >= len(used) - 3:
                print(f"{k:3d} {f(s):8.2f} {dep} {pro} {f(e):8.2f} {(e - s) / 1e3:6.2f} {(e - max(s, prev)) / 1e3:6.2f}  {extra}")
            prev = e
        print(f"step span {(buf[used[-1], 1] - t0) / 1e3:.1f} us, tp_error {m.be.counter('tp_error')}")
    m.close()
    sys.exit(0)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_layers = sys.argv[2] if len(sys.argv) > 2 else "4"
ctx = sys.argv[3] if len(sys.argv) > 3 else "256"
mode = sys.argv[4] if len(sys.argv) > 4 else "p2p"
with tempfile.TemporaryDirectory() as td:
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--rank", str(r), str(size), n_layers, ctx, mode, td],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(size)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    print(outs[0])
    for r, p in enumerate(procs):
        if p.returncode:
            print(f"rank {r} failed:\n{outs[r][-2000:]}")

"""One rank of a tensor-parallel group (one process per GPU): loads a synthetic model directory with row-sharded
weights, prefills a prompt, decodes greedily on the device, and dumps ids (+ the logits of a few host-driven steps).

    python tools/tp_worker.py <rank> <size> <id_file> <model_dir> <n_prompt> <n_decode> <out_prefix> [nccl|p2p]

The 128-byte NCCL id travels through <id_file> (rank 0 writes it), the 64-byte CUDA-IPC handles of the exchange heaps
through <id_file>.ipc<rank>; any other transport works as well.  Mode `p2p` (default) = all-gathers fused into the
producing kernels as peer stores (`p2p_fence`: with the fence + epoch-flag protocol everywhere); `nccl` = NCCL all-gathers
between the kernels.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powerserve_b200 import capi, synth  # noqa: E402

rank, size, id_file, model_dir, n_prompt, n_decode, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
if rank == 0:
    nid = capi.tp_unique_id()
    with open(id_file + ".tmp", "wb") as f:
        f.write(nid)
    os.replace(id_file + ".tmp", id_file)
else:
    t0 = time.time()
    while not os.path.exists(id_file):
        if time.time() - t0 > 120:
            raise SystemExit("no NCCL id after 120 s")
        time.sleep(0.05)
    nid = open(id_file, "rb").read()
mode = sys.argv[8] if len(sys.argv) > 8 else "p2p"
m = capi.CudaModel(model_dir, max_batch=32, device=rank, tp_rank=rank, tp_size=size, nccl_id=nid)
if mode.startswith("p2p"):
    with open(f"{id_file}.ipc{rank}.tmp", "wb") as f:
        f.write(m.tp_export())
    os.replace(f"{id_file}.ipc{rank}.tmp", f"{id_file}.ipc{rank}")
    handles = []
    for r in range(size):
        t0 = time.time()
        while not os.path.exists(f"{id_file}.ipc{r}"):
            if time.time() - t0 > 120:
                raise SystemExit("no IPC handle after 120 s")
            time.sleep(0.05)
        handles.append(open(f"{id_file}.ipc{r}", "rb").read())
    m.tp_import(handles)
    if mode == "p2p_fence":   # the fence + epoch-flag protocol for every exchange (default: in-band flags for the per-layer ones)
        m.be.set_option("tp_ll", 0)
prompt = synth.random_prompt(m.vocab, n_prompt, seed=11)
ids, logits = m.generate(prompt, 4, batch_size=16)            # host-driven steps: logits gathered on every rank
m.reset()
m.prefill(prompt, 16)
dev_ids = m.decode_greedy(int(prompt[-1]), n_decode)           # graph-replayed device loop
m.reset()
batch_logits = m.forward(prompt[:6], lm_head=True)                # a batch WITH lm_head (the speculative-verify shape)
batch_dev = m.be.read_device(m.be.logits_dev(), 6 * m.vocab).reshape(6, m.vocab)
np.savez(f"{out}.rank{rank}.npz", ids=np.asarray(ids, np.int32), logits=logits, dev_ids=dev_ids, batch_logits=batch_logits, batch_dev=batch_dev,
         gathers=m.be.counter("tp_allgathers"), p2p=m.be.counter("tp_p2p"), tp_error=m.be.counter("tp_error"),
         ms=m.be.counter("last_device_ns") / 1e6 / n_decode)
m.close()

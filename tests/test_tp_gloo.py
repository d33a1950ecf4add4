"""Tensor-parallel sharding logic on CPU: world_size 2 over torch.distributed `gloo`.  Each rank runs the oracle's
operators on its row shards (tests/_tp_sim.py, a phase-by-phase mirror of the CUDA decode step) and the shards are
exchanged with dist.all_gather; the result must be BIT-IDENTICAL to the unsharded oracle (row sharding keeps every dot
product whole).  Also checks the sharding plan's validation and byte ranges."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from powerserve_b200 import gguf, synth
from tests import _tp_plan as tp
from tests import _libs as L
from tests import _model as M


def _worker(rank, size, port, model_dir, prompt, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from tests._tp_sim import ShardedOracle

    def all_gather(a, slot=None):
        t = torch.from_numpy(np.ascontiguousarray(a))
        outs = [torch.empty_like(t) for _ in range(size)]
        dist.all_gather(outs, t)
        return torch.cat(outs).numpy()

    m = ShardedOracle(model_dir, rank, size, all_gather)
    logits = []
    for tok in prompt[:-1]:
        m.step(int(tok), lm_head=False)
    tok = int(prompt[-1])
    for _ in range(3):
        lg = m.step(tok)
        logits.append(lg)
        tok = int(np.argmax(lg))
    np.save(os.path.join(out_dir, f"logits{rank}.npy"), np.stack(logits))
    dist.destroy_process_group()


@pytest.mark.parametrize("preset", ["tiny-llama", "tiny-qwen2"])
def test_row_sharded_forward_is_bit_identical(preset):
    d = M.model_dir(preset)
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, 7, seed=21)
    om = M.OracleModel(d)
    _, ref = om.generate(prompt, 3, batch_size=1)
    om.close()
    with tempfile.TemporaryDirectory() as td:
        port = 29600 + os.getpid() % 300
        mp.spawn(_worker, args=(2, port, d, prompt, td), nprocs=2, join=True)
        for r in range(2):
            L.assert_bit_equal(np.load(os.path.join(td, f"logits{r}.npy")), ref, f"{preset}: rank {r} sharded logits vs unsharded oracle")


def _inband_worker(rank, size, names, n_words, model_dir, prompt, out_dir):
    from multiprocessing import shared_memory

    from tests._tp_sim import InbandExchange, ShardedOracle
    shms = {s: [shared_memory.SharedMemory(name=names[s][p]) for p in range(size)] for s in InbandExchange.SLOTS}
    mirrors = {s: [np.ndarray((n_words,), np.uint64, buffer=m.buf) for m in shms[s]] for s in InbandExchange.SLOTS}
    ex = InbandExchange(rank, size, mirrors)
    m = ShardedOracle(model_dir, rank, size, ex)
    logits = []
    for tok in prompt[:-1]:
        m.step(int(tok), lm_head=False)
    tok = int(prompt[-1])
    for _ in range(3):
        lg = m.step(tok)
        logits.append(lg)
        tok = int(np.argmax(lg))
    np.save(os.path.join(out_dir, f"logits{rank}.npy"), np.stack(logits))
    del mirrors, ex, m
    for s in shms.values():
        for h in s:
            h.close()


def test_inband_flag_exchange_model():
    """The fence-free exchange of the CUDA tensor-parallel path, modelled on CPU with real concurrency: two processes,
    64-bit {value, epoch} words in shared memory, no barrier anywhere - the only synchronisation is the consumer's poll."""
    from multiprocessing import shared_memory

    from tests._tp_sim import InbandExchange
    preset = "tiny-llama"
    d = M.model_dir(preset)
    sh = synth.PRESETS[preset]
    prompt = synth.random_prompt(sh.vocab_size, 7, seed=21)
    om = M.OracleModel(d)
    _, ref = om.generate(prompt, 3, batch_size=1)
    om.close()
    n_words = max(sh.ffn_dim, sh.vocab_size, sh.dim)
    shms = {s: [shared_memory.SharedMemory(create=True, size=8 * n_words) for _ in range(2)] for s in InbandExchange.SLOTS}
    try:
        for s in shms.values():
            for h in s:
                np.ndarray((n_words,), np.uint64, buffer=h.buf)[:] = 0
        names = {s: [h.name for h in v] for s, v in shms.items()}
        with tempfile.TemporaryDirectory() as td:
            mp.spawn(_inband_worker, args=(2, names, n_words, d, prompt, td), nprocs=2, join=True)
            for r in range(2):
                L.assert_bit_equal(np.load(os.path.join(td, f"logits{r}.npy")), ref, f"rank {r}: in-band-flag exchange vs unsharded oracle")
    finally:
        for s in shms.values():
            for h in s:
                h.close()
                h.unlink()


def test_sharding_plan():
    with pytest.raises(ValueError):
        tp.validate(32, 8, 14336, 128256, 4096, 3)
    tp.validate(32, 8, 14336, 128256, 4096, 8)
    g = gguf.GGUFFile(os.path.join(M.model_dir("tiny-llama"), "ggml", "weights.gguf"))
    t = g["blk.0.ffn_down.weight"]                      # {K = ffn, rows = dim}
    row_bytes = gguf.tensor_bytes(t.ggml_type, (t.shape[0], 1))
    off, rows = tp.shard_tensor("blk.0.ffn_down.weight", t, 1, 2)
    assert rows == t.shape[1] // 2 and off == rows * row_bytes
    off, rows = tp.shard_tensor("blk.0.attn_norm.weight", g["blk.0.attn_norm.weight"], 1, 2)
    assert off == 0                                      # norm weights are replicated
    assert tp.gathers_per_token(32) == 129

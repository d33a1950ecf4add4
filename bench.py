#!/usr/bin/env python
"""bench.py — BASELINE.json metric: "Llama-3.1-8B Q4_K decode tok/s & prefill tok/s @1 GPU; HBM GB/s vs peak".

    python bench.py --gpus N --steps K --warmup W            # our CUDA backend (through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own ggml CPU path (oracle/_ref)

A "step" is ONE DECODED TOKEN of the named model with `--prompt` tokens of context already in the KV cache
(BASELINE.json configs[2]: Llama-3.1-8B Q4_K, prefill 2048 + decode).  `value` is decode tokens/s with everything
resident in HBM, timed with CUDA events on the backend's stream; `e2e` is the same metric through the reference-
facing call (`ps_cuda_forward`: HOST token ids in, HOST logits out, every step); prefill tokens/s is reported in
`prefill`.  Weights are synthetic random Q4_K blocks (no model files / network on the box); they are 4.2 GB per
token, far larger than the 126 MB L2, so no explicit L2 flush is needed between steps.

Multi-GPU (N > 1): ONE model row-sharded over the N GPUs (BASELINE.json configs[4], tensor parallel, strong scaling):
every rank owns 1/N of the rows of every matrix and its own kv heads, outputs are exchanged with NCCL all-gathers inside
the decode graph (bit-identical to one GPU).  `--replicas` runs N independent replicas instead (weak scaling, no
collective on the data path; `value` is the sum over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from powerserve_b200 import gguf, synth  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def weight_bytes_per_token(shape: synth.ModelShape) -> int:
    """Algorithmic bytes of one decoded token: every matmul weight block as stored in the GGUF, read once
    (SURVEY.md section 8(d)): 7 matrices per layer + lm_head."""
    n = 0
    for name, t, shp, _ in synth.tensor_plan(shape):
        if t != gguf.GGML_F32 and name != "token_embd.weight":
            n += gguf.tensor_bytes(t, shp)
    if shape.tied:  # lm_head re-reads the embedding table
        n += gguf.tensor_bytes(shape.embd_type or shape.wtype, (shape.dim, shape.vocab_size))
    return n


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def __enter__(self):
        def run():
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    r = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5)
                    if r.returncode == 0 and r.stdout.strip():
                        self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
                except Exception:
                    pass
                self._stop.wait(0.1)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def cpu_isa_v4() -> bool:
    try:
        flags = open("/proc/cpuinfo").read().split("flags", 1)[1].split("\n", 1)[0].split()
    except Exception:
        return False
    return all(f in flags for f in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))


def ref_binary(timing: bool):
    """The compiled reference driver: the AVX2 build (GGML_NATIVE=OFF flags, the one parity is pinned on) or, for TIMING on a
    host with AVX-512, the x86-64-v4 build (what GGML_NATIVE=ON would select there)."""
    v4 = os.path.join(ROOT, "oracle", "_ref", "v4", "ps_ref_run")
    if timing and os.path.exists(v4) and cpu_isa_v4():
        return v4, "x86-64-v4 (AVX-512) timing build"
    return os.path.join(ROOT, "oracle", "_ref", "ps_ref_run"), "AVX2+FMA+F16C build (GGML_NATIVE=OFF flags)"


class RefModelDir:
    """The synthetic model written once as a PowerServe model directory (model.json + ggml/weights.gguf) for the CPU legs."""

    def __init__(self, shape, tensors):
        self.td = tempfile.TemporaryDirectory(dir=os.environ.get("PS_BENCH_TMP", None))
        d = self.td.name
        os.makedirs(os.path.join(d, "ggml"))
        json.dump(synth.model_json(shape), open(os.path.join(d, "model.json"), "w"))
        gguf.write_gguf(os.path.join(d, "ggml", "weights.gguf"), tensors, arch=shape.arch)
        self.path = d

    def close(self):
        self.td.cleanup()


def cpu_reference_run(mdir: str, vocab: int, prompt, n_decode, n_threads, batch_size=128, timing=False, forced=None, dump_logits=0):
    """Run the reference's own CPU implementation (oracle/_ref/ps_ref_run: PowerServe's model -> graph -> executor ->
    GGMLBackend -> vendored ggml, compiled from /root/reference by oracle/Makefile) on the SAME weights."""
    exe, isa = ref_binary(timing)
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/ps_ref_run is missing: run `make -C oracle ref` where /root/reference exists")
    pf = os.path.join(mdir, "prompt.txt")
    open(pf, "w").write(" ".join(str(int(t)) for t in prompt))
    cmd = [exe, mdir, str(n_threads), str(batch_size), pf, str(n_decode), os.path.join(mdir, "out")]
    if forced is not None:
        ff = os.path.join(mdir, "forced.txt")
        open(ff, "w").write(" ".join(str(int(t)) for t in forced))
        cmd += ["--force", ff]
    if dump_logits:
        cmd += ["--dump-logits", str(dump_logits)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
    if r.returncode != 0:
        raise RuntimeError("ps_ref_run failed: " + r.stderr[-1000:])
    tm = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    ids = [int(x) for x in open(os.path.join(mdir, "out.ids")).read().split()]
    logits = np.fromfile(os.path.join(mdir, "out.logits"), dtype=np.float32).reshape(-1, vocab) if dump_logits else None
    return {"decode_tok_s": tm["decode_tok_s"], "prefill_tok_s": tm["prefill_tok_s"], "ids": ids, "logits": logits, "isa": isa}


PARITY_PROMPT, PARITY_STEPS = 17, 9


def parity_inputs(shape):
    return synth.random_prompt(shape.vocab_size, PARITY_PROMPT, seed=1234), synth.random_prompt(shape.vocab_size, PARITY_STEPS, seed=4321)


def gpu_parity_leg(model, shape):
    """Teacher-forced parity sample on the benchmarked weights: 17-token prompt fed with batch_size 1 (the chunking every
    parallelism mode shares), then 9 steps whose inputs are FORCED random ids (so the ids are not a fixed point).
    Returns the greedy ids and a hash of the logits bits; 1 GPU, tensor-parallel and the CPU reference must agree."""
    import hashlib
    cp, forced = parity_inputs(shape)
    ids, lg = model.generate(cp, PARITY_STEPS, batch_size=1, forced=forced)
    return [int(x) for x in ids], hashlib.sha256(np.ascontiguousarray(lg).tobytes()).hexdigest()[:16]


def cpu_parity_leg(mdir, shape, cpu_threads):
    import hashlib
    cp, forced = parity_inputs(shape)
    res = cpu_reference_run(mdir, shape.vocab_size, cp, PARITY_STEPS, cpu_threads, batch_size=1, forced=forced, dump_logits=PARITY_STEPS)
    return res, hashlib.sha256(np.ascontiguousarray(res["logits"]).tobytes()).hexdigest()[:16]


def decode_leg(model, tok, steps, warmup):
    """device-resident greedy decode: (tok/s, ms/step) from the backend's CUDA events"""
    model.decode_greedy(tok, warmup)
    model.decode_greedy(tok, steps)
    ms = model.be.counter("last_device_ns") / 1e6 / steps
    return 1e3 / ms, ms


def extra_config1_1b(args, hbm_peak):
    """BASELINE configs[1]: Llama-3.2-1B Q4_K, 32-token prompt, decode (dequant-matvec HBM roofline)."""
    from powerserve_b200 import capi
    shape = synth.PRESETS["llama-3.2-1b"]
    shape.n_ctx = 4096
    tensors = synth.generate_tensors(shape, args.seed)
    tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
    desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=32, n_ctx=shape.n_ctx)
    m = capi.CudaModel(desc=desc, tensors=tmap)
    prompt = synth.random_prompt(shape.vocab_size, 33, seed=1234)
    m.prefill(prompt, 32)
    tps, ms = decode_leg(m, int(prompt[-1]), 256, 8)
    wb = weight_bytes_per_token(shape)
    t0 = time.perf_counter()
    t = int(prompt[-1])
    for _ in range(64):
        t = int(np.argmax(m.forward([t])[0]))
    e2e = 64 / (time.perf_counter() - t0)
    out = {"workload": "llama-3.2-1b Q4_K synthetic: 32-token prompt, 256 greedy decode tokens (BASELINE configs[1])", "value": tps, "unit": "tok/s", "ms_per_step": ms,
           "e2e": {"value": e2e, "unit": "tok/s", "what": "host token -> ps_cuda_forward -> host logits, 64 steps"}, "weight_bytes_per_token": wb,
           "roofline_step": {"achieved": wb / (ms * 1e-3) / 1e9, "peak": hbm_peak, "frac": wb / (ms * 1e-3) / 1e9 / hbm_peak, "unit": "GB/s"}}
    return out, m, tensors, shape


def extra_config0_qwen2(args, hbm_peak):
    """BASELINE configs[0]: Qwen2-0.5B Q4_0 (the reference's own CPU-runnable case), 32-token prompt, greedy decode - the fused
    32-block path (ps_mv32.cuh: Q4_0 weights x Q8_0 activations, NEOX rope, q|k|v biases, 7 query heads per kv head)."""
    from powerserve_b200 import capi
    shape = synth.PRESETS["qwen2-0.5b"]
    shape.n_ctx = 4096
    tensors = synth.generate_tensors(shape, args.seed)
    tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
    desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=32, n_ctx=shape.n_ctx)
    m = capi.CudaModel(desc=desc, tensors=tmap)
    prompt = synth.random_prompt(shape.vocab_size, 33, seed=1234)
    m.prefill(prompt, 32)
    tps, ms = decode_leg(m, int(prompt[-1]), 64, 8)
    wb = weight_bytes_per_token(shape)
    out = {"workload": "qwen2-0.5b Q4_0 synthetic: 32-token prompt, 64 greedy decode tokens (BASELINE configs[0])", "value": tps, "unit": "tok/s", "ms_per_step": ms,
           "fused_32_block_path": bool(m.be.counter("mv32_ok")), "weight_bytes_per_token": wb,
           "roofline_step": {"achieved": wb / (ms * 1e-3) / 1e9, "peak": hbm_peak, "frac": wb / (ms * 1e-3) / 1e9 / hbm_peak, "unit": "GB/s"}}
    m.close()
    return out


def extra_sessions(model, shape, hbm_peak, n_sess=64, n_steps=24):
    """Server-side batching on the benchmarked model: n_sess independent sessions (own KV sets, 32-token
    prompts) advanced one token each per forward pass through ps_cuda_forward_sessions - one weight stream per pass, so the
    aggregate rate may exceed the one-sequence HBM roofline.  Device time of the passes; ids come back, logits stay on the device."""
    sids = [0] + [model.session_create() for _ in range(n_sess - 1)]
    toks = []
    for k, sid in enumerate(sids):
        model.session_select(sid)
        model.reset()
        p = synth.random_prompt(shape.vocab_size, 33, seed=4321 + k)
        model.prefill(p, 32)
        toks.append(int(p[-1]))
    model.session_select(0)
    ns = 0.0
    for step in range(4 + n_steps):
        _, ids = model.forward_sessions(sids, toks, want_logits=False)
        toks = [int(t) for t in ids]
        if step >= 4:
            ns += model.be.counter("last_device_ns")
    for sid in sids[1:]:
        model.session_destroy(sid)
    model.reset()
    agg = n_sess * n_steps / (ns * 1e-9)
    wb = weight_bytes_per_token(shape)
    return {"workload": f"{n_sess} concurrent sessions x 1 token per forward pass, 32-token prompts (ps_cuda_forward_sessions)", "value": agg, "unit": "tok/s (aggregate)",
            "ms_per_pass": ns * 1e-6 / n_steps, "sessions": n_sess, "one_sequence_hbm_roofline_tok_s": hbm_peak * 1e9 / wb,
            "note": "every session's logits are bit-identical to decoding it alone (tests/test_gpu_sessions.py)"}


def extra_config3_spec(args, target, target_shape, draft, n_tokens=96):
    """BASELINE configs[3]: Llama-3.1-8B target + Llama-3.2-1B draft, token-tree speculative decoding (draft_batch_size 12, tree
    defaults of speculative_config.hpp:21-36), 32-token prompt.  Synthetic weights: the two models are uncorrelated, so the
    acceptance rate is that of random drafts - the leg measures the machinery (tree verify width 12, KV slot operations), not
    a speed-up; the plain greedy rate of the same target on the same prompt stands beside it."""
    from powerserve_b200 import capi
    prompt = synth.random_prompt(target_shape.vocab_size, 33, seed=1234)
    sd = capi.SpecDecoder(target, draft)
    sd.generate(prompt, 8, prefill_batch=32)                       # warm-up
    ids, st = sd.generate(prompt, n_tokens, prefill_batch=32)
    sd.close()
    dec_s = st["draft_s"] + st["verify_s"]
    target.reset(); target.prefill(prompt, 32)
    plain_ids = [int(x) for x in target.decode_greedy(int(prompt[-1]), n_tokens)]
    plain_tps = n_tokens / (target.be.counter("last_device_ns") / 1e9)
    agree = next((k for k in range(n_tokens) if int(ids[k]) != plain_ids[k]), n_tokens)
    return {"workload": "llama-3.1-8b target + llama-3.2-1b draft, token-tree speculative decode, 32-token prompt (BASELINE configs[3])",
            "value": st["n_generated_tokens"] / dec_s, "unit": "tok/s", "tokens": int(st["n_generated_tokens"]), "iterations": int(st["n_iterations"]),
            "tokens_per_iteration": st["n_generated_tokens"] / max(st["n_iterations"], 1), "draft_forwards_per_iteration": st["n_draft_times"] / max(st["n_iterations"], 1),
            "accepted_draft_tokens": int(st["n_accepted_tokens"]), "draft_s": st["draft_s"], "verify_s": st["verify_s"],
            "ms_per_iteration": 1e3 * dec_s / max(st["n_iterations"], 1), "plain_greedy_tok_s_same_prompt": plain_tps,
            "ids_equal_plain_greedy_prefix": int(agree), "note": "synthetic (uncorrelated) weights: random-draft acceptance; losslessness is tested in tests/test_gpu_spec.py"}


def extra_powerserve_stack(mdir, shape, prompt_len, n_decode, cpu_threads, device_topk=0):
    """e2e through PowerServe's OWN stack: Model::forward -> graph -> executor -> CUDA_FORWARD op -> libps_cuda.so, logits into the
    executor's CPUBuffer, host arg-max per token (powerserve_b200/host/_build/ps_cuda_run, the drop-in demonstration of INTEGRATION.md)."""
    exe = os.path.join(ROOT, "powerserve_b200", "host", "_build", "ps_cuda_run")
    if not os.path.exists(exe):
        return {"value": None, "why": "powerserve_b200/host/_build/ps_cuda_run not built (needs /root/reference at build time)"}
    pf = os.path.join(mdir, "prompt_stack.txt")
    open(pf, "w").write(" ".join(str(int(t)) for t in synth.random_prompt(shape.vocab_size, prompt_len + 1, seed=1234)))
    extra = ["--device-topk", str(device_topk)] if device_topk else []
    r = subprocess.run([exe, mdir, str(cpu_threads), "128", pf, str(n_decode), os.path.join(mdir, "stack_out")] + extra, capture_output=True, text=True, timeout=1200)
    if r.returncode != 0:
        return {"value": None, "why": ("ps_cuda_run failed: " + r.stderr[-300:])}
    tm = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    return {"value": tm["decode_tok_s"], "unit": "tok/s", "prefill_tok_s": tm["prefill_tok_s"], "context": prompt_len, "steps": n_decode,
            "what": "PowerServe's Model::forward / executor / Platform on the CUDA backend (CUDA_FORWARD graph op), host arg-max over the CPUBuffer logits; wall clock like app/run/run.cpp:96-154"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-3.1-8b")
    ap.add_argument("--prompt", type=int, default=2048, help="tokens of context before the timed decode steps")
    ap.add_argument("--prefill-batch", type=int, default=128)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tp-nccl", action="store_true", help="tensor parallel with NCCL all-gathers instead of the fused peer-store kernels")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas instead of tensor parallelism")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (1B config, speculative config, PowerServe-stack e2e)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    shape = synth.PRESETS[args.model]
    # context capacity: the default run fits 4096 (the attention kernel then keeps its V^T rows in shared memory); a longer
    # --steps / --prompt grows it (every leg below rolls the KV position back to prompt + warmup before it starts)
    shape.n_ctx = max(shape.n_ctx, 4096, -(-(args.prompt + 1 + args.warmup + args.steps + 32) // 256) * 256)
    hbm_peak, peak_src = peaks()
    wbytes = weight_bytes_per_token(shape)
    n_cpu = os.cpu_count() or 2
    cpu_threads = args.cpu_threads or max(1, min(n_cpu - 1, 32))
    which = "configs[1]: 1B decode" if args.model == "llama-3.2-1b" else f"configs[2]: prefill {args.prompt} + decode"
    cfg = {"workload": f"{args.model} Q4_K synthetic: decode 1 token/step at context {args.prompt}+ (BASELINE {which})",
           "context": args.prompt, "prefill_batch": args.prefill_batch, "weight_bytes_per_token": wbytes,
           "l2": "weights (>= 0.7 GB/token) exceed the 126 MB L2; no explicit flush",
           "parallelism": "1 GPU" if world == 1 else (f"replicas x{world}" if args.replicas else f"tp{world} (row-sharded, " + ("NCCL all-gather)" if args.tp_nccl else "all-gather fused into the kernels as NVLink peer stores)"))}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # like for like: the SAME prompt length and decode steps as our arm (a 2048-token CPU prefill is about a minute);
        # only the step count is bounded (<= 64 timed tokens) and the real numbers are what the line states
        n_timed = max(2, min(args.steps, 64))
        n_warm = max(1, min(args.warmup, 4))   # ps_ref_run's decode clock already excludes the first token (run.cpp:96-154)
        tensors = synth.generate_tensors(shape, args.seed)
        prompt = synth.random_prompt(shape.vocab_size, args.prompt + 1, seed=1234)
        t0 = time.time()
        md = RefModelDir(shape, tensors)
        try:
            res = cpu_reference_run(md.path, shape.vocab_size, prompt, n_warm + n_timed, cpu_threads, batch_size=args.prefill_batch, timing=True)
        finally:
            md.close()
        cfg["reference_isa"] = res["isa"]
        out = {"impl": "reference", "metric": "decode_tok_s", "value": res["decode_tok_s"], "unit": "tok/s", "n_gpus": args.gpus,
               "steps": n_timed + n_warm - 1, "warmup": 1, "steps_requested": args.steps, "ms_per_step": 1000.0 / max(res["decode_tok_s"], 1e-9),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8xint4 dot, fp32 accumulate", "data": "synthetic",
               "config": cfg, "prefill": {"value": res["prefill_tok_s"], "unit": "tok/s", "tokens": len(prompt) - 1},
               "cpu_baseline": {"value": res["decode_tok_s"], "unit": "tok/s", "cores": cpu_threads, "kind": "reference",
                                "sample": f"{len(prompt) - 1}-token prefill (batch {args.prefill_batch}) + {n_warm + n_timed} greedy decode tokens at context "
                                          f"{args.prompt}+ of the full model, wall clock like app/run/run.cpp:96-154 (first decode token excluded); {res['isa']}"},
               "e2e": {"value": res["decode_tok_s"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "wall_s": time.time() - t0}
        print(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA, through the C ABI)
    import torch
    import torch.distributed as dist

    from powerserve_b200 import build, capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    build.build()
    tp = world if (world > 1 and not args.replicas) else 1
    tensors = synth.generate_tensors(shape, args.seed + (1000 * rank if tp == 1 else 0))   # tensor parallel: the SAME model on every rank
    tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
    nccl_id = None
    if tp > 1:  # rank 0's NCCL id reaches the other ranks through torch.distributed (plumbing only)
        idt = torch.frombuffer(bytearray(capi.tp_unique_id() if rank == 0 else bytes(128)), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=args.prefill_batch, n_ctx=shape.n_ctx, qkv_bias=shape.qkv_bias,
                                     tp_rank=rank if tp > 1 else 0, tp_size=tp)
    model = capi.CudaModel(desc=desc, tensors=tmap, device=local_rank, nccl_id=nccl_id)
    if tp > 1 and not args.tp_nccl:  # fused compute + all-gather over NVLink peer memory: exchange the CUDA-IPC handles
        mine = torch.frombuffer(bytearray(model.tp_export()), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        model.tp_import([bytes(h.cpu().numpy().tobytes()) for h in allh])
    prompt = synth.random_prompt(shape.vocab_size, args.prompt + 1, seed=1234)

    # prefill (ModelTokenIterator loop: prompt[:-1] in chunks, lm_head=false), wall clock through the C ABI; one untimed chunk
    # first (lazy allocations, NCCL channel set-up of a tensor-parallel group)
    model.prefill(prompt[:args.prefill_batch + 1], args.prefill_batch)
    model.reset()
    barrier_fn = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize
    barrier_fn()
    t0 = time.perf_counter()
    model.prefill(prompt, args.prefill_batch)
    prefill_s = time.perf_counter() - t0
    tok = int(prompt[-1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- value: device-resident decode (token id fed back on the device), CUDA events on the backend stream
    model.decode_greedy(tok, args.warmup)
    barrier()
    l0 = model.be.counter("kernel_launches")
    with ClockSampler(local_rank) as clk:
        ids = model.decode_greedy(tok, args.steps)
        dev_ms = model.be.counter("last_device_ns") / 1e6
    launches = model.be.counter("kernel_launches") - l0
    barrier()
    # --- roofline leg: the dominant kernel (the row-walker Q4_K mat-vec, 5 launch sites per step: q|k|v, o, gate|up, down,
    # lm_head) timed launch by launch with CUDA events on the backend stream (un-graphed, no PDL overlap, same step sequence)
    model.be.kv_rollback(args.steps)
    k_steps = min(args.steps, 16)
    model.be.set_option("ktime", 1)
    model.decode_greedy(tok, 2)
    model.decode_greedy(tok, k_steps)
    mv_ns, mv_n = model.be.counter("matvec_kernel_ns"), model.be.counter("matvec_kernel_launches")
    model.be.set_option("ktime", 0)
    model.be.kv_rollback(2 + k_steps)
    # the same launches INSIDE the graph-replayed, PDL-chained step: every kernel stamps %globaltimer at start / end (option
    # "trace"); a launch's exclusive time is end - max(start, end of the previous launch).  Explains the gap between the
    # event-timed figure above (serialised launches: + launch latency, no overlap of the weight prefetch) and the step time.
    in_graph = None
    if tp == 1:
        n_l = 1 + 6 * shape.n_layers + 2   # embed, (qkv, attn1, attn2, wo, gate|up, down) x layers, lm_head, arg-max
        if n_l <= 255:
            model.be.set_option("trace", 1)
            model.decode_greedy(tok, 1)
            tb = np.zeros((n_l, 8), np.int64)
            model.be._ck(model.be.L.ps_cuda_read_trace(model.be.h, tb.ctypes.data, n_l))
            model.be.set_option("trace", 0)
            model.be.kv_rollback(1)
            excl = np.maximum(tb[1:, 1] - np.maximum(tb[1:, 0], tb[:-1, 1]), 0)   # launch k + 1 vs launch k
            mv_idx = [6 * L + o for L in range(shape.n_layers) for o in (0, 3, 4, 5)] + [6 * shape.n_layers]  # into excl (launch - 1)
            att_idx = [6 * L + o for L in range(shape.n_layers) for o in (1, 2)]
            in_graph = {"matvec_us_per_step": float(excl[mv_idx].sum() / 1e3), "attention_us_per_step": float(excl[att_idx].sum() / 1e3),
                        "step_span_us": float((tb[-1, 1] - tb[0, 0]) / 1e3), "launches": len(mv_idx)}
    barrier()
    # --- e2e: host token -> ps_cuda_forward -> host logits, every step
    h0, d0 = model.be.counter("h2d_bytes"), model.be.counter("d2h_bytes")
    t0 = time.perf_counter()
    t = tok
    for _ in range(args.steps):
        lg = model.forward([t])[0]
        t = int(np.argmax(lg))
    e2e_s = time.perf_counter() - t0
    h2d = (model.be.counter("h2d_bytes") - h0) / args.steps
    d2h = (model.be.counter("d2h_bytes") - d0) / args.steps
    barrier()

    ms_dev, e2e_ms = dev_ms, e2e_s * 1e3
    if world > 1:
        tt = torch.tensor([ms_dev, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, e2e_ms = tt.tolist()
    n_models = world if tp == 1 else 1                      # replicas decode `world` streams, a tensor-parallel group one
    value = n_models * args.steps / (ms_dev / 1e3)
    e2e_value = n_models * args.steps / (e2e_ms / 1e3)
    wbytes_gpu = wbytes // tp
    step_gbs = wbytes_gpu * args.steps / (ms_dev / 1e3) / 1e9  # per GPU, whole step
    mv_bytes_per_launch = wbytes_gpu * k_steps / max(mv_n, 1)      # algorithmic bytes (of this GPU's shard) of an average mat-vec launch
    mv_avg_s = mv_ns / 1e9 / max(mv_n, 1)
    achieved = mv_bytes_per_launch / max(mv_avg_s, 1e-12) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_matvec_traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath) and tp == 1:
        traffic = json.load(open(tpath)).get(args.model, {}).get("dram_bytes_per_launch")
    out = {"metric": "decode_tok_s", "value": value, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak" if tp == 1 else "strong", "vs_baseline": None,
           "dtype": "int8xint4 dot, fp32 accumulate (bit-exact with the ggml CPU reference)", "data": "synthetic", "config": cfg,
           "prefill": {"value": n_models * (len(prompt) - 1) / prefill_s, "unit": "tok/s", "tokens": len(prompt) - 1, "timing": "wall clock through ps_cuda_forward"},
           "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": traffic, "peak_source": peak_src + ", sustained copy",
                        "kernel": "ps_k_rw_matvec (row-walker Q4_K mat-vec, all launch sites of the decode step)",
                        "bytes_per_launch": mv_bytes_per_launch, "avg_launch_us": mv_avg_s * 1e6, "launches_timed": int(mv_n),
                        "what": "algorithmic GGUF weight bytes per mat-vec launch / average launch duration (CUDA events on the backend stream)",
                        "step": {"achieved": step_gbs, "frac": step_gbs / hbm_peak,
                                 "what": "whole decode step incl. attention, prologues and launch gaps: weight bytes per token / step time"}},
           "clocks": clk.summary(), "greedy_ids_head": [int(x) for x in ids[:8]]}
    if in_graph:  # the same kernel inside the graph: algorithmic bytes of the step / summed exclusive time of its mat-vec launches
        ig_gbs = wbytes_gpu / (in_graph["matvec_us_per_step"] * 1e-6) / 1e9
        out["roofline"]["in_graph"] = dict(in_graph, achieved=ig_gbs, frac=ig_gbs / hbm_peak,
                                           what="graph + PDL replay, %globaltimer stamps inside the kernels (option trace): weight bytes per token / "
                                                "summed exclusive time of the mat-vec launches; the traced step runs ~3 % slower than the timed ones")
    if tp > 1:
        out["tp"] = {"size": tp, "p2p": bool(model.be.counter("tp_p2p")), "peer_wait_error": model.be.counter("tp_error"), "nccl_allgathers_per_step": (model.be.counter("tp_allgathers")) // max(1, model.be.counter("graph_replays") + 1),
                     "note": "row sharding keeps every dot product whole: results are bit-identical to one GPU (tests/test_gpu_tp.py)"}
    # --- secondary legs on the same box (N = 1 only): BASELINE configs[1] and configs[3]; they explain, they are not the headline
    kv_bytes = 2 * shape.n_layers * shape.kv_dim * 4 * (args.prompt + args.warmup + args.steps // 2)
    out["roofline"]["step_incl_kv"] = {"achieved": (wbytes_gpu + kv_bytes // tp) / (ms_dev / args.steps * 1e-3) / 1e9, "kv_bytes_per_token": kv_bytes,
                                       "frac": (wbytes_gpu + kv_bytes // tp) / (ms_dev / args.steps * 1e-3) / 1e9 / hbm_peak,
                                       "what": "the same step counting the fp32 K/V cache rows the attention reads as well (weights + KV bytes per token / step time)"}
    if world == 1 and not args.no_extras and args.model == "llama-3.1-8b":
        extras = {}
        try:
            c1, m1b, t1b, s1b = extra_config1_1b(args, hbm_peak)
            extras["configs[1]"] = c1
            try:
                extras["configs[3]"] = extra_config3_spec(args, model, shape, m1b)
            except Exception as e:
                extras["configs[3]"] = {"value": None, "why": str(e)[:300]}
            m1b.close()
        except Exception as e:
            extras["configs[1]"] = {"value": None, "why": str(e)[:300]}
        try:
            extras["sessions"] = extra_sessions(model, shape, hbm_peak)
        except Exception as e:
            extras["sessions"] = {"value": None, "why": str(e)[:300]}
        try:
            extras["configs[0]"] = extra_config0_qwen2(args, hbm_peak)
        except Exception as e:
            extras["configs[0]"] = {"value": None, "why": str(e)[:300]}
        out["extras"] = extras
    # --- parity on the benchmarked weights (every N): teacher-forced sample, ids + logits-bits hash; rank 0 checks them against
    # the compiled CPU reference (AVX2 build) run on the same weights and inputs
    closed = False
    gp_ids, gp_hash = gpu_parity_leg(model, shape)
    out["parity"] = {"sample": f"{PARITY_PROMPT}-token prompt (batch 1) + {PARITY_STEPS} teacher-forced steps", "greedy_ids": gp_ids, "logits_sha256_16": gp_hash}
    if world > 1:
        hs = [None] * world
        dist.all_gather_object(hs, gp_hash)
        out["parity"]["all_ranks_equal"] = len(set(hs)) == 1
    if rank == 0 and not args.no_cpu_baseline:
        try:
            md = RefModelDir(shape, tensors)
            try:
                pres, chash = cpu_parity_leg(md.path, shape, cpu_threads)
                out["parity"].update({"cpu_logits_sha256_16": chash, "logits_bit_exact": chash == gp_hash, "greedy_ids_match": pres["ids"] == gp_ids})
                if world == 1 and not args.no_extras:
                    model.close()       # the stack binary binds its own context: free this one's HBM first
                    closed = True
                    out["e2e_powerserve_stack"] = extra_powerserve_stack(md.path, shape, args.prompt, min(args.steps, 64), cpu_threads)
                    lazy = extra_powerserve_stack(md.path, shape, args.prompt, min(args.steps, 64), cpu_threads, device_topk=40)
                    lazy["what"] = "the same stack with device-side sampling (SURVEY 8 f3): logits stay on the device, TopKSampler(40) runs there, 80 words per token cross PCIe"
                    out["e2e_powerserve_stack_device_topk"] = lazy
                if world == 1:
                    cp = synth.random_prompt(shape.vocab_size, 17, seed=1234)
                    res = cpu_reference_run(md.path, shape.vocab_size, cp, 9, cpu_threads)
                    out["cpu_baseline"] = {"value": res["decode_tok_s"], "unit": "tok/s", "cores": cpu_threads, "kind": "reference",
                                           "sample": "17-token prompt + 9 greedy decode tokens of the full model (same weights), " + res["isa"] +
                                                     "; the like-for-like CPU number (same context and steps) is `bench.py --impl reference`",
                                           "prefill_tok_s": res["prefill_tok_s"], "greedy_ids_match": out["parity"]["greedy_ids_match"]}
            finally:
                md.close()
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            out["cpu_baseline"] = {"value": None, "unit": "tok/s", "cores": cpu_threads, "kind": "unavailable", "sample": str(e)[:200]}
    if not closed:
        model.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

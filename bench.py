#!/usr/bin/env python
"""encode_batch throughput of the B200 encode path (GB/s of input UTF-8 bytes, bit-exact ids).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N ...            # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[1] -- cl100k_base, 100 000 synthetic ~1 KB
English-like documents (tools/synth.py cfg2, seed 102 + 1000*rank).  Weak scaling: every rank
encodes its own 100 000-document shard; the only collective is the all-gather of the
per-rank id counts (NCCL).  One JSON line is printed by rank 0.

value      device-resident: packed bytes + offsets already in HBM, ids + offsets left in HBM;
           every step timed with CUDA events on the launching stream, L2 flushed between steps.
e2e        the same batch through the C-ABI host call spl_encode_batch from pinned host
           buffers: H2D copy, kernels, D2H of ids + offsets inside the timed region.
python_api the Python methods on a 10 000-document sample: Tokenizer.encode_batch (list[str] ->
           list[list[int]], the reference's signature) and encode_batch_packed (numpy arrays out).
roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration, against the
           measured HBM copy bandwidth (MEASURED_PEAKS.json, else the profiling guide's fallback).
cpu_baseline  oracle/c_oracle.c (C restatement of the reference algorithm, PCRE2-JIT regex,
           OpenMP over documents = Rayon par_iter) on a bounded sample, rank 0 at N=1.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

METRIC = "encode_batch GB/s input bytes (bit-exact ids)"
UNIT = "GB/s"
WORKLOAD = "cfg2: cl100k_base, 100k synthetic ~1 KB English docs (BASELINE.json configs[1])"
HBM_FALLBACK_GBS = 6650.0


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(rank: int, n_docs: int):
    import synth
    from splintr_b200 import presets as P
    vb = P.load_vocab_bytes("cl100k_base.tiktoken")
    data, offsets = synth.cfg2(vb, n_docs, seed_offset=1000 * rank)
    return vb, np.ascontiguousarray(data), np.ascontiguousarray(offsets, dtype=np.uint64)


def cpu_sample_docs(offsets, target_bytes):
    d = int(np.searchsorted(offsets, np.uint64(target_bytes), side="left"))
    return max(1, min(d, len(offsets) - 1))


def run_cpu(vb, data, offsets, budget_s: float, steps: int = 1, warmup: int = 0):
    """C oracle with all host threads on a bounded sample; returns (GB/s, cores, sample text, ids, out_off, n_docs)."""
    from oracle.c_oracle import COracle, max_threads
    from splintr_b200 import presets as P
    p = P.PRESETS["cl100k_base"]
    orc = COracle(vb, p.pattern, p.special_tokens, False)
    # every host core this process may run on (torchrun exports OMP_NUM_THREADS=1: do not inherit that for the baseline)
    cores = max(len(os.sched_getaffinity(0)), 1) if hasattr(os, "sched_getaffinity") else max(os.cpu_count() or 1, 1)
    nd0 = cpu_sample_docs(offsets, 4 << 20)
    t0 = time.perf_counter()
    orc.encode_packed(data[:int(offsets[nd0])], offsets[:nd0 + 1], n_threads=cores)
    rate = int(offsets[nd0]) / max(time.perf_counter() - t0, 1e-6)
    nd = cpu_sample_docs(offsets, min(rate * budget_s, float(offsets[-1])))
    sb, so = data[:int(offsets[nd])], offsets[:nd + 1]
    for _ in range(warmup):
        orc.encode_packed(sb, so, n_threads=cores)
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        ids, off = orc.encode_packed(sb, so, n_threads=cores)
        times.append(time.perf_counter() - t0)
    gbs = len(sb) / (sum(times) / len(times)) / 1e9
    sample = f"first {nd} docs ({len(sb) / 1e6:.1f} MB) of the workload, {cores} OpenMP threads, PCRE2-JIT, LRU omitted"
    return gbs, cores, sample, ids, off, nd, sum(times) / len(times)


def reference_arm(args, rank):
    if rank != 0:
        return
    vb, data, offsets = make_workload(0, args.docs)
    budget = max(2.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    gbs, cores, sample, _, _, nd, sec = run_cpu(vb, data, offsets, budget, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
            "config": {"workload": WORKLOAD, "docs_per_step": nd, "note": "reference algorithm (C restatement; the Rust crate "
                       "cannot be built in this image) on the host cores"},
            "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=100_000, help="documents per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the encode path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / debug lines: not on stdout (one JSON line there)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from splintr_b200 import _lib, Tokenizer
    if _lib.needs_build():
        if local_rank == 0:
            _lib.build()
        barrier()
    lib = _lib.load()
    vb, data, offsets = make_workload(rank, args.docs)
    n_bytes, n_docs = int(len(data)), len(offsets) - 1
    tok = Tokenizer.from_pretrained("cl100k_base", devices=[local_rank])

    # ---- device-resident inputs -----------------------------------------------------------
    pad = (-n_bytes) % 16
    d_bytes_full = torch.zeros(n_bytes + pad, dtype=torch.uint8, device=dev)
    d_bytes_full[:n_bytes].copy_(torch.from_numpy(data))
    d_bytes = d_bytes_full[:n_bytes]
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_ids = torch.empty(n_bytes, dtype=torch.int32, device=dev)
    d_out = torch.empty(n_docs + 1, dtype=torch.int64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    counts = torch.zeros(world, dtype=torch.int64, device=dev)

    def step_device():
        tok.encode_device(d_bytes, d_off, ids_out=d_ids, out_offsets=d_out, sync=False)
        if world > 1:                                                       # the path's only exchange step
            dist.all_gather_into_tensor(counts, d_out[n_docs:n_docs + 1])

    no_flush = os.environ.get("SPL_BENCH_NO_FLUSH") == "1"     # diagnostics only: the reported runs always flush
    for _ in range(args.warmup):
        if not no_flush:
            flush.zero_()
        step_device()
    barrier()
    tok.set_profiling(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ktimes = {}
    barrier()
    for i in range(args.steps):
        if not no_flush:
            flush.zero_()
        ev[i][0].record()
        step_device()
        ev[i][1].record()
        if i + 1 == args.steps or i % 4 == 3:
            torch.cuda.synchronize()
            for k, v in tok.last_kernel_times().items():
                ktimes.setdefault(k, []).append(v)
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    tot_bytes = torch.tensor([float(n_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_bytes, op=dist.ReduceOp.SUM)
    tok.set_profiling(False)
    n_tok = int(d_out[n_docs].item())
    dev_ms_per_step = float(total_ms.item()) / args.steps
    value = float(tot_bytes.item()) / (dev_ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host C-ABI call (pinned host buffers) ----------------------
    h_ptr = lib.spl_alloc_pinned(n_bytes + 64)
    h_off_ptr = lib.spl_alloc_pinned((n_docs + 1) * 8)
    ctypes.memmove(h_ptr, data.ctypes.data, n_bytes)
    ctypes.memmove(h_off_ptr, offsets.ctypes.data, (n_docs + 1) * 8)
    h_off = np.ctypeslib.as_array(ctypes.cast(h_off_ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(n_docs + 1,))

    e2e_stats = {}

    def step_e2e():
        res = ctypes.c_void_p()
        rc = lib.spl_encode_batch(tok._handle, ctypes.c_void_p(h_ptr), ctypes.c_void_p(h_off_ptr), n_docs, 0, ctypes.byref(res))
        if rc != 0:
            raise RuntimeError(_lib.last_error(tok._handle))
        st = _lib.SplStats()
        lib.spl_result_stats(res, ctypes.byref(st))
        e2e_stats.update(h2d=int(st.h2d_bytes), d2h=int(st.d2h_bytes), tokens=int(st.n_tokens), dev_ms=float(st.total_ms),
                         launches=int(st.n_launches))
        first = int(ctypes.cast(lib.spl_result_ids(res), ctypes.POINTER(ctypes.c_uint32))[0]) if st.n_tokens else 0   # host reads the result
        lib.spl_result_free(res)
        if world > 1:
            dist.all_gather_into_tensor(counts, torch.tensor([st.n_tokens], dtype=torch.int64, device=dev))
        return first

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = float(tot_bytes.item()) * args.steps / float(e2e_s.item()) / 1e9
    assert e2e_stats["tokens"] == n_tok, "host and device entry points disagree on the id count"

    # ---- the Python surface (SURVEY 8d "T3"): list[str] in, list[list[int]] / packed arrays out; rank 0, N=1 ----
    py_api = None
    if rank == 0 and world == 1:
        import synth
        nd_s = min(n_docs, 10_000)
        texts = synth.unpack_texts(data[:int(offsets[nd_s])], offsets[:nd_s + 1])
        nb_s = int(offsets[nd_s])
        best = {"list": 1e9, "packed": 1e9}
        for _ in range(3):
            t1 = time.perf_counter(); tok.encode_batch(texts); best["list"] = min(best["list"], time.perf_counter() - t1)
            t1 = time.perf_counter(); tok.encode_batch_packed(texts); best["packed"] = min(best["packed"], time.perf_counter() - t1)
        py_api = {"encode_batch_list_of_lists": nb_s / best["list"] / 1e9, "encode_batch_packed": nb_s / best["packed"] / 1e9,
                  "unit": UNIT, "sample": f"first {nd_s} docs ({nb_s / 1e6:.1f} MB), best of 3, str packing and result objects included"}

    # ---- CPU baseline + parity spot check (rank 0, N=1) ------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, cores, sample, c_ids, c_off, nd, _ = run_cpu(vb, data, offsets, 12.0)
        g_off = d_out[:nd + 1].cpu().numpy().astype(np.uint64)
        g_ids = d_ids[:int(g_off[-1])].cpu().numpy().astype(np.uint32)
        parity = bool(np.array_equal(g_off, c_off) and np.array_equal(g_ids, c_ids))
        cpu = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        peak, peak_src = hbm_peak()
        kmean = {k: sum(v) / len(v) for k, v in ktimes.items()}
        dom = max(kmean, key=kmean.get) if kmean else None
        # algorithmic bytes of the whole path per launch of the dominant kernel:
        # every input byte read once, every u32 id written once, u64 doc offsets in and out
        b_alg = n_bytes + 4 * n_tok + 16 * (n_docs + 1)
        roof = None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom, {}).get("bytes") if args.docs == 100_000 else None
        except Exception:
            traffic = None
        if dom:
            ach = b_alg / (kmean[dom] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg,
                    "kernel_ms": kmean, "kernel_share_of_step": kmean[dom] / max(sum(kmean.values()), 1e-9)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/u32 integer", "data": "synthetic",
                "config": {"workload": WORKLOAD, "docs_per_gpu": n_docs, "bytes_per_gpu": n_bytes, "tokens_per_gpu": n_tok,
                           "parallelism": f"doc-sharded dp{world}", "l2": "NOT flushed (diagnostic run)" if no_flush else "flushed between steps (512 MiB memset)",
                           "timing": "per-step CUDA events on the launching stream, max over ranks"},
                "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_stats["h2d"], "d2h_bytes_per_step": e2e_stats["d2h"],
                        "timing": "host wall clock around spl_encode_batch, barrier + synchronize both sides, max over ranks",
                        "device_ms_per_step": e2e_stats["dev_ms"]},
                "python_api": py_api,
                "gpu_launches": args.steps * tok.launches_per_call(False), "clocks": clocks,
                "ids_match_cpu_baseline": parity}
        print(json.dumps(line), flush=True)
    lib.spl_free_pinned(h_ptr)
    lib.spl_free_pinned(h_off_ptr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

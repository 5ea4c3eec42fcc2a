#!/usr/bin/env python
"""encode_batch throughput of the B200 encode path (GB/s of input UTF-8 bytes, bit-exact ids).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N ...            # the reference algorithm on the host cores

Headline workload (config.workload): BASELINE.json configs[1] -- cl100k_base, 100 000 synthetic ~1 KB
English-like documents (tools/synth.py cfg2, seed 102 + 1000*rank).  Weak scaling: every rank encodes its own
shard; the path's only collective is the all-gather of the per-rank id counts (NCCL).  Rank 0 prints ONE JSON line.

value        device-resident: packed bytes + offsets already in HBM, ids + offsets left in HBM; every step timed with
             CUDA events on the launching stream, L2 flushed between steps, max over ranks.
e2e          the same batch through the C-ABI host call spl_encode_batch from pinned host buffers: H2D copy, kernels,
             D2H of ids + offsets inside the timed region.  e2e.pcie_ceiling is what plain duplex cudaMemcpyAsync of
             the same bytes reaches with all ranks copying at once (the bus, not the kernels, bounds e2e).
roofline     whole step and dominant kernel: algorithmic bytes per launch / CUDA-event duration, against the measured
             HBM copy bandwidth (MEASURED_PEAKS.json, else the profiling guide's fallback).
cpu_baseline oracle/c_oracle.c (C restatement of the reference algorithm, PCRE2-JIT regex, OpenMP over documents =
             Rayon par_iter) on a bounded sample, rank 0 at N=1; at N>1 every rank checks a sample of its own ids.
configs      the other BASELINE.json configs (cfg1, cfg3, cfg4, cfg5) at the per-GPU sizes named there, each with
             value / e2e / roofline / cpu_baseline / ids_match_cpu_baseline.
strong       ONE host batch (cfg3 shard tiled to >= 2 GB) through ONE handle that spans all N GPUs
             (spl_create(devices=[0..N-1])): the reference's Tokenizer::encode_batch call shape, strong scaling.
python_api   the Python methods on a 10 000-document sample and single-text / small-batch latency (cfg1).
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
HBM_FALLBACK_GBS = 6650.0
HEADLINE = "cfg2"

# per-GPU shard of every BASELINE.json config (seed offset 1000 * rank)
CONFIGS = {
    "cfg1": dict(vocab="cl100k_base", docs=1_000,
                 workload="cfg1: cl100k_base, 1 000 short English texts (~100 B each) (BASELINE.json configs[0])"),
    "cfg2": dict(vocab="cl100k_base", docs=100_000,
                 workload="cfg2: cl100k_base, 100k synthetic ~1 KB English docs (BASELINE.json configs[1])"),
    "cfg3": dict(vocab="o200k_base", docs=125_000,
                 workload="cfg3: o200k_base, mixed code/JSON/prose ~2 KB docs, 125 000 docs per GPU = the 8-GPU shard of the "
                          "1M-doc batch (BASELINE.json configs[2])"),
    "cfg4": dict(vocab="llama3", docs=100,
                 workload="cfg4: llama3, long ~1 MB docs with deep per-piece merge chains, 100 docs per GPU = the 1 % "
                          "sample of the 10k-doc batch BASELINE.md allows (BASELINE.json configs[3])"),
    "cfg5": dict(vocab="deepseek_v3", docs=100_000,
                 workload="cfg5: deepseek_v3, 100k UTF-8-heavy Chinese docs (~1.5 KB) (BASELINE.json configs[4])"),
}


def generate(cfg: str, rank: int, docs=None):
    import synth
    from splintr_b200 import presets as P
    c = CONFIGS[cfg]
    vb = P.load_vocab_bytes(P.PRESETS[c["vocab"]].vocab_file)
    n = docs if docs is not None else c["docs"]
    so = 1000 * rank
    if cfg == "cfg1":
        d, o = synth.cfg1(vb, n, seed_offset=so)
    elif cfg == "cfg2":
        d, o = synth.cfg2(vb, n, seed_offset=so)
    elif cfg == "cfg3":
        d, o = synth.cfg3(vb, n, seed_offset=so)
    elif cfg == "cfg4":
        d, o = synth.cfg4(vb, n, 1_000_000.0, seed_offset=so)
    else:
        d, o = synth.cfg5(vb, n, seed_offset=so)
    return vb, np.ascontiguousarray(d), np.ascontiguousarray(o, dtype=np.uint64)


def config_dict(cfg: str, n_docs: int, n_bytes: int, world: int):
    """identical in both arms (the driver compares the dicts)"""
    return {"workload": CONFIGS[cfg]["workload"], "docs_per_gpu": n_docs, "bytes_per_gpu": n_bytes,
            "parallelism": f"doc-sharded dp{world}"}


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


def host_cores() -> int:
    return max(len(os.sched_getaffinity(0)), 1) if hasattr(os, "sched_getaffinity") else max(os.cpu_count() or 1, 1)


def bind_to_gpu_numa(local_rank: int):
    """One process per GPU: run on (and first-touch the pinned buffers from) the NUMA node the GPU hangs off, so that
    host<->device copies do not cross the socket link.  Returns a description for the bench line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        if bus is None:
            return "unbound (no PCI id)"
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return "unbound (single node)"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"unbound (node {node} outside the affinity mask)"
        os.sched_setaffinity(0, cpus)
        return f"node {node} ({len(cpus)} cpus)"
    except Exception as e:                                   # noqa: BLE001
        return f"unbound ({type(e).__name__})"


_ORACLES = {}


def c_oracle_for(vocab: str):
    from oracle.c_oracle import COracle
    from splintr_b200 import presets as P
    if vocab not in _ORACLES:
        p = P.PRESETS[vocab]
        _ORACLES[vocab] = COracle(P.load_vocab_bytes(p.vocab_file), p.pattern, p.special_tokens, p.byte_level)
    return _ORACLES[vocab]


def sample_docs(offsets, target_bytes, min_docs=1):
    d = int(np.searchsorted(offsets, np.uint64(max(int(target_bytes), 1)), side="left"))
    return max(min_docs, min(max(d, 1), len(offsets) - 1))


def run_cpu(vocab, data, offsets, budget_s: float, steps: int = 1, warmup: int = 0, threads=None):
    """C oracle with all host threads on a bounded sample of the workload.
    Returns (GB/s, cores, sample text, ids, out_off, n_docs, seconds per step)."""
    orc = c_oracle_for(vocab)
    # every host core this process may run on (torchrun exports OMP_NUM_THREADS=1: do not inherit that for the baseline)
    cores = threads or host_cores()
    n_docs = len(offsets) - 1
    nd0 = sample_docs(offsets, 4 << 20, min(cores, n_docs))           # OpenMP runs over documents: at least one per thread
    t0 = time.perf_counter()
    orc.encode_packed(data[:int(offsets[nd0])], offsets[:nd0 + 1], n_threads=cores)
    rate = int(offsets[nd0]) / max(time.perf_counter() - t0, 1e-6)
    nd = sample_docs(offsets, min(rate * budget_s, float(offsets[-1])), min(cores, n_docs))
    sb, so = data[:int(offsets[nd])], offsets[:nd + 1]
    for _ in range(warmup):
        orc.encode_packed(sb, so, n_threads=cores)
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        ids, off = orc.encode_packed(sb, so, n_threads=cores)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    gbs = len(sb) / sec / 1e9
    sample = f"first {nd} docs ({len(sb) / 1e6:.1f} MB) of the workload, {cores} OpenMP threads, PCRE2-JIT, LRU omitted"
    return gbs, cores, sample, ids, off, nd, sec


def reference_arm(args, rank):
    if rank != 0:
        return
    cfg = args.config
    _, data, offsets = generate(cfg, 0, args.docs)
    budget = max(2.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    gbs, cores, sample, _, _, nd, sec = run_cpu(CONFIGS[cfg]["vocab"], data, offsets, budget, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
            "config": config_dict(cfg, len(offsets) - 1, int(len(data)), args.gpus),
            "note": "reference algorithm (C restatement, oracle/c_oracle.c; the Rust crate cannot be built in this image) on the "
                    f"host cores; each step encodes a bounded sample: {sample}",
            "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Ctx:
    pass


def measure(cx, cfg: str, steps: int, warmup: int, cpu_budget: float, traffic_table):
    """Device-resident value, e2e, roofline, CPU baseline and parity of one config on this rank's shard."""
    import torch
    import torch.distributed as dist
    from splintr_b200 import Tokenizer, _lib
    lib, dev, world, rank = cx.lib, cx.dev, cx.world, cx.rank
    c = CONFIGS[cfg]
    docs = cx.docs_override if (cfg == HEADLINE and cx.docs_override) else None
    t_gen = time.perf_counter()
    _, data, offsets = generate(cfg, rank, docs)
    t_gen = time.perf_counter() - t_gen
    n_bytes, n_docs = int(len(data)), len(offsets) - 1
    if c["vocab"] not in cx.toks:
        cx.toks[c["vocab"]] = Tokenizer.from_pretrained(c["vocab"], devices=[cx.local_rank])
    tok = cx.toks[c["vocab"]]

    # ---- device-resident inputs -----------------------------------------------------------
    pad = (-n_bytes) % 16
    d_bytes_full = torch.zeros(n_bytes + pad, dtype=torch.uint8, device=dev)
    d_bytes_full[:n_bytes].copy_(torch.from_numpy(data))
    d_bytes = d_bytes_full[:n_bytes]
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_ids = torch.empty(max(n_bytes, 16), dtype=torch.int32, device=dev)
    d_out = torch.empty(n_docs + 1, dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)

    def step_device():
        tok.encode_device(d_bytes, d_off, ids_out=d_ids, out_offsets=d_out, sync=False)
        if world > 1:                                                       # the path's only exchange step
            dist.all_gather_into_tensor(counts, d_out[n_docs:n_docs + 1])

    for _ in range(warmup):
        if not cx.no_flush:
            cx.flush.zero_()
        step_device()
    cx.barrier()
    # ---- the timed region: K steps, each ONE graph launch of the pass (8 kernel nodes) on the current stream ----------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    cx.barrier()
    for i in range(steps):
        if not cx.no_flush:
            cx.flush.zero_()
        ev[i][0].record()
        step_device()
        ev[i][1].record()
    torch.cuda.synchronize()
    cx.barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    tot_bytes = torch.tensor([float(n_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_bytes, op=dist.ReduceOp.SUM)
    # ---- K further steps launched kernel by kernel with CUDA events around every kernel (roofline.kernel_ms): the
    # events cost the step ~10 %, so they are kept out of the timed region; the step time under them is reported too
    tok.set_profiling(True)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ktimes = {}
    for i in range(steps):
        if not cx.no_flush:
            cx.flush.zero_()
        ev2[i][0].record()
        step_device()
        ev2[i][1].record()
        if i + 1 == steps or i % 4 == 3:
            torch.cuda.synchronize()
            for k, v in tok.last_kernel_times().items():
                ktimes.setdefault(k, []).append(v)
    tok.set_profiling(False)
    prof_step_ms = sum(a.elapsed_time(b) for a, b in ev2) / steps
    cx.barrier()
    n_tok = int(d_out[n_docs].item())
    dev_ms_per_step = float(total_ms.item()) / steps
    value = float(tot_bytes.item()) / (dev_ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host C-ABI call (pinned host buffers) ----------------------
    h_ptr = lib.spl_alloc_pinned(n_bytes + 64)
    h_off_ptr = lib.spl_alloc_pinned((n_docs + 1) * 8)
    ctypes.memmove(h_ptr, data.ctypes.data, n_bytes)
    ctypes.memmove(h_off_ptr, offsets.ctypes.data, (n_docs + 1) * 8)
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
        return first

    for _ in range(warmup):
        step_e2e()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    torch.cuda.synchronize()
    t_mine = time.perf_counter() - t0
    cx.barrier()
    e2e_s = torch.tensor([t_mine], dtype=torch.float64, device=dev)
    tok_all = torch.tensor([float(e2e_stats["tokens"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(tok_all, op=dist.ReduceOp.SUM)                      # the exchange step of the host path: id counts
    e2e_value = float(tot_bytes.item()) * steps / float(e2e_s.item()) / 1e9
    assert e2e_stats["tokens"] == n_tok, "host and device entry points disagree on the id count"
    lib.spl_free_pinned(h_ptr)
    lib.spl_free_pinned(h_off_ptr)

    # ---- CPU baseline + parity ----------------------------------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not cx.no_cpu:
        gbs, cores, sample, c_ids, c_off, nd, _ = run_cpu(c["vocab"], data, offsets, cpu_budget)
        g_off = d_out[:nd + 1].cpu().numpy().astype(np.uint64)
        g_ids = d_ids[:int(g_off[-1])].cpu().numpy().astype(np.uint32)
        parity = bool(np.array_equal(g_off, c_off) and np.array_equal(g_ids, c_ids))
        cpu = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    elif world > 1 and not cx.no_cpu:
        # every rank holds its own ids against the oracle on a sample of its shard (few threads: N ranks share the
        # host), and the global output offsets (splintr_b200/distributed.py: the all-gathered counts) must tile
        from splintr_b200.distributed import exchange_counts
        nd = sample_docs(offsets, 2 << 20, 1)
        thr = max(1, host_cores() // world)
        c_ids, c_off = c_oracle_for(c["vocab"]).encode_packed(data[:int(offsets[nd])], offsets[:nd + 1], n_threads=thr)
        g_off = d_out[:nd + 1].cpu().numpy().astype(np.uint64)
        g_ids = d_ids[:int(g_off[-1])].cpu().numpy().astype(np.uint32)
        ok = bool(np.array_equal(g_off, c_off) and np.array_equal(g_ids, c_ids))
        cnt = exchange_counts(n_docs, n_tok, dev)
        ok = ok and int(cnt[rank, 1]) == n_tok and int(cnt[:, 1].sum()) == int(tok_all.item())
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = bool(flag.item())

    # ---- roofline -------------------------------------------------------------------------------------
    peak, peak_src = hbm_peak()
    kmean = {k: sum(v) / len(v) for k, v in ktimes.items()}
    dom = max(kmean, key=kmean.get) if kmean else None
    # algorithmic bytes of the whole path per launch: every input byte read once, every u32 id
    # written once, u64 document offsets in and out
    b_alg = n_bytes + 4 * n_tok + 16 * (n_docs + 1)
    roof = None
    if dom:
        step_local_ms = sum(step_ms) / len(step_ms)
        ach_dom = b_alg / (kmean[dom] * 1e-3) / 1e9
        ach_step = b_alg / (step_local_ms * 1e-3) / 1e9
        tr = (traffic_table or {}).get(cfg) or {}
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach_dom, "peak": peak, "unit": "GB/s", "frac": ach_dom / peak,
                "traffic": (tr.get(dom) or {}).get("bytes"),
                "whole_step_achieved": ach_step, "whole_step_frac": ach_step / peak,
                "traffic_whole_step": sum(v.get("bytes", 0) for v in tr.values()) if tr else None,
                "traffic_source": tr and "profiles/ncu_traffic.json (ncu --set full of this workload, per launch)" or None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg,
                "kernel_ms": kmean, "kernel_share_of_step": kmean[dom] / max(sum(kmean.values()), 1e-9),
                "kernel_ms_from": f"{steps} further steps launched kernel by kernel with CUDA events around every kernel "
                                  f"({prof_step_ms:.4f} ms per step that way); the timed steps are one graph launch each",
                "ms_per_step_kernel_by_kernel": prof_step_ms}
    out = {"value": value, "unit": UNIT, "ms_per_step": dev_ms_per_step,
           "config": config_dict(cfg, n_docs, n_bytes, world), "tokens_per_gpu": n_tok,
           "roofline": roof, "cpu_baseline": cpu,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_stats["h2d"], "d2h_bytes_per_step": e2e_stats["d2h"],
                   "ms_per_step": float(e2e_s.item()) / steps * 1e3, "device_ms_per_step": e2e_stats["dev_ms"],
                   "timing": "host wall clock around spl_encode_batch, barrier + synchronize both sides, max over ranks"},
           "ids_match_cpu_baseline": parity, "launches_per_step": tok.launches_per_call(False),
           "launches_per_call": e2e_stats["launches"], "generate_s": round(t_gen, 2)}
    keep = Ctx()
    keep.data, keep.offsets, keep.tok, keep.d_ids, keep.d_out, keep.n_tok = data, offsets, tok, d_ids, d_out, n_tok
    return out, keep


def pcie_ceiling(cx, h2d_bytes: int, d2h_bytes: int, reps: int = 5):
    """Plain duplex cudaMemcpyAsync of one step's bytes from / to pinned host memory, all ranks at once."""
    import torch
    import torch.distributed as dist
    dev = cx.dev
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_o = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = {}
    for mode in ("h2d", "d2h", "duplex"):
        ts = []
        for _ in range(reps + 1):
            cx.barrier()
            t0 = time.perf_counter()
            if mode in ("h2d", "duplex"):
                with torch.cuda.stream(s1):
                    for a, b in zip(d_in.chunk(8), h_in.chunk(8)):
                        a.copy_(b, non_blocking=True)
            if mode in ("d2h", "duplex"):
                with torch.cuda.stream(s2):
                    for a, b in zip(h_out.chunk(8), d_o.chunk(8)):
                        a.copy_(b, non_blocking=True)
            s1.synchronize(); s2.synchronize()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor([min(ts[1:])], dtype=torch.float64, device=dev)
        if cx.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best[mode] = float(t.item())
    return best


def strong_scaling(cx, keep3):
    """Rank 0: ONE batch through ONE handle over all N GPUs (the reference's single encode_batch call)."""
    import torch
    from splintr_b200 import Tokenizer, _lib
    lib = cx.lib
    data, offsets = keep3.data, keep3.offsets
    n1, nd1 = int(len(data)), len(offsets) - 1
    reps = max(1, -(-(2 << 30) // n1))
    n, nd = n1 * reps, nd1 * reps
    h_ptr = lib.spl_alloc_pinned(n + 64)
    h_off_ptr = lib.spl_alloc_pinned((nd + 1) * 8)
    h_off = np.ctypeslib.as_array(ctypes.cast(h_off_ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(nd + 1,))
    for r in range(reps):
        ctypes.memmove(h_ptr + r * n1, data.ctypes.data, n1)
        h_off[r * nd1:(r + 1) * nd1] = offsets[:-1] + np.uint64(r * n1)
    h_off[nd] = n
    n_dev = getattr(cx, "strong_devices", cx.world)
    tok = Tokenizer.from_pretrained(CONFIGS["cfg3"]["vocab"], devices=list(range(n_dev)))
    times, ok, n_tok, stats = [], True, 0, None
    for it in range(3):
        res = ctypes.c_void_p()
        t0 = time.perf_counter()
        rc = lib.spl_encode_batch(tok._handle, ctypes.c_void_p(h_ptr), ctypes.c_void_p(h_off_ptr), nd, 0, ctypes.byref(res))
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(_lib.last_error(tok._handle))
        st = _lib.SplStats()
        lib.spl_result_stats(res, ctypes.byref(st))
        n_tok = int(st.n_tokens)
        if it:
            times.append(dt)
        if it == 2:
            # every tile of the batch is the shard this rank encoded alone (device-resident, checked against the oracle)
            ids = np.ctypeslib.as_array(ctypes.cast(lib.spl_result_ids(res), ctypes.POINTER(ctypes.c_uint32)), shape=(n_tok,))
            off = np.ctypeslib.as_array(ctypes.cast(lib.spl_result_offsets(res), ctypes.POINTER(ctypes.c_uint64)), shape=(nd + 1,))
            one = keep3.d_ids[:keep3.n_tok].cpu().numpy().astype(np.uint32)
            one_off = keep3.d_out.cpu().numpy().astype(np.uint64)
            ok = n_tok == keep3.n_tok * reps
            for r in (0, reps // 2, reps - 1):
                ok = ok and np.array_equal(ids[r * keep3.n_tok:(r + 1) * keep3.n_tok], one)
                ok = ok and np.array_equal(off[r * nd1:(r + 1) * nd1 + 1], one_off + np.uint64(r * keep3.n_tok))
            stats = {"h2d_bytes": int(st.h2d_bytes), "d2h_bytes": int(st.d2h_bytes), "device_ms": float(st.total_ms)}
        lib.spl_result_free(res)
    del tok
    lib.spl_free_pinned(h_ptr)
    lib.spl_free_pinned(h_off_ptr)
    best = min(times)
    return {"value": n / best / 1e9, "unit": UNIT, "scaling": "strong", "n_gpus": n_dev, "batch_bytes": n, "batch_docs": nd,
            "ms_per_call": best * 1e3, "ids_match_single_device": bool(ok), **(stats or {}),
            "what": f"one spl_encode_batch call, one handle over devices 0..{n_dev - 1} (single process), pinned host in, ids + "
                    f"offsets in host memory out; batch = the cfg3 shard tiled {reps}x; best of 2 after 1 warm-up, host wall clock"}


def python_api(keep, cx):
    """T3: the Python methods (list[str] in), rank 0 at N=1."""
    import synth
    tok, data, offsets = keep.tok, keep.data, keep.offsets
    nd_s = min(len(offsets) - 1, 10_000)
    texts = synth.unpack_texts(data[:int(offsets[nd_s])], offsets[:nd_s + 1])
    nb_s = int(offsets[nd_s])
    best = {"list": 1e9, "packed": 1e9}
    for _ in range(3):
        t1 = time.perf_counter(); tok.encode_batch(texts); best["list"] = min(best["list"], time.perf_counter() - t1)
        t1 = time.perf_counter(); tok.encode_batch_packed(texts); best["packed"] = min(best["packed"], time.perf_counter() - t1)
    return {"encode_batch_list_of_lists": nb_s / best["list"] / 1e9, "encode_batch_packed": nb_s / best["packed"] / 1e9,
            "unit": UNIT, "sample": f"first {nd_s} docs ({nb_s / 1e6:.1f} MB), best of 3, str packing and result objects included"}


def small_batch_latency(keep1):
    """cfg1 through the Python API: one ~100-byte text per call and the 1 000-text batch, next to the C oracle."""
    import synth
    tok, data, offsets = keep1.tok, keep1.data, keep1.offsets
    texts = synth.unpack_texts(data, offsets)
    orc = c_oracle_for(CONFIGS["cfg1"]["vocab"])
    for t in texts[:20]:
        tok.encode(t)
    t0 = time.perf_counter()
    for t in texts[:200]:
        tok.encode(t)
    enc_us = (time.perf_counter() - t0) / 200 * 1e6
    t0 = time.perf_counter()
    for t in texts[:200]:
        orc.encode(t)
    cpu_enc_us = (time.perf_counter() - t0) / 200 * 1e6
    best = {"list": 1e9, "packed": 1e9, "cpu": 1e9}
    for _ in range(5):
        t1 = time.perf_counter(); tok.encode_batch(texts); best["list"] = min(best["list"], time.perf_counter() - t1)
        t1 = time.perf_counter(); tok.encode_batch_packed(texts); best["packed"] = min(best["packed"], time.perf_counter() - t1)
        t1 = time.perf_counter(); orc.encode_packed(data, offsets, n_threads=host_cores()); best["cpu"] = min(best["cpu"], time.perf_counter() - t1)
    nb = int(len(data))
    return {"encode_us_per_text": enc_us, "cpu_encode_us_per_text": cpu_enc_us,
            "encode_batch_1000_texts_us": best["list"] * 1e6, "encode_batch_packed_1000_texts_us": best["packed"] * 1e6,
            "cpu_encode_batch_1000_texts_us": best["cpu"] * 1e6,
            "encode_batch_mb_s": nb / best["list"] / 1e6, "cpu_encode_batch_mb_s": nb / best["cpu"] / 1e6,
            "note": "Python Tokenizer.encode(text) / encode_batch(texts) incl. str packing and list building; cpu = oracle/c_oracle.c "
                    "through ctypes at the packed boundary (all host threads for the batch)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=HEADLINE, choices=list(CONFIGS), help="reference arm: which config (default: the headline)")
    ap.add_argument("--docs", type=int, default=None, help="documents per rank of the headline config (default: the named size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the configs / strong / python_api blocks (profiling runs)")
    ap.add_argument("--only-strong", action="store_true", help="single process: just the strong-scaling arm over --gpus devices")
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
    numa = bind_to_gpu_numa(local_rank)
    gloo = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / debug lines: not on stdout (one JSON line there)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        gloo = dist.new_group(backend="gloo")                     # host-side barrier (an NCCL barrier spins on the GPU)

    cx = Ctx()
    cx.dev, cx.world, cx.rank, cx.local_rank = dev, world, rank, local_rank
    cx.no_cpu = args.no_cpu_baseline
    cx.no_flush = os.environ.get("SPL_BENCH_NO_FLUSH") == "1"     # diagnostics only: the reported runs always flush
    cx.docs_override = args.docs
    cx.toks = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    cx.barrier = barrier

    from splintr_b200 import _lib
    # build (if needed) by local rank 0 only; everybody waits -- the decision is broadcast so that the barrier matches
    need = torch.tensor([1 if (local_rank == 0 and _lib.needs_build()) else 0], device=dev)
    if world > 1:
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
    if int(need.item()):
        if local_rank == 0:
            _lib.build()
        barrier()
    cx.lib = _lib.load()
    cx.flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    if args.only_strong:                                     # development aid: python bench.py --gpus N --only-strong (no torchrun)
        cx.strong_devices = args.gpus
        cx.no_cpu = True
        _, keep3 = measure(cx, "cfg3", 2, 3, 0.0, None)
        print(json.dumps({"strong": strong_scaling(cx, keep3)}), flush=True)
        return

    traffic_table = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic_table = json.load(f)
    except Exception:
        traffic_table = None

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    head, keep2 = measure(cx, HEADLINE, args.steps, args.warmup, 12.0, traffic_table)
    clocks = sampler.stop() if rank == 0 else None

    configs, strong, py_api, small, ceiling = {}, None, None, None, None
    if not args.only_headline:
        ceiling = pcie_ceiling(cx, head["e2e"]["h2d_bytes_per_step"], head["e2e"]["d2h_bytes_per_step"])
        keep = {}
        for cfg in ("cfg1", "cfg3", "cfg4", "cfg5"):
            configs[cfg], keep[cfg] = measure(cx, cfg, min(args.steps, 10), min(args.warmup, 3), 4.0, traffic_table)
            if cfg not in ("cfg1", "cfg3"):
                keep[cfg] = None
            torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            py_api = python_api(keep2, cx)
            if not cx.no_cpu:
                small = small_batch_latency(keep["cfg1"])
        # strong scaling: rank 0 drives one handle over all GPUs, the other ranks wait on the host
        barrier()
        if rank == 0:
            try:
                strong = strong_scaling(cx, keep["cfg3"])
            except Exception as e:                                   # noqa: BLE001
                strong = {"error": f"{type(e).__name__}: {e}"}
        if gloo is not None:
            dist.barrier(group=gloo)

    if rank == 0:
        e2e = dict(head["e2e"])
        if ceiling:
            nb = head["config"]["bytes_per_gpu"] * world
            e2e["pcie_ceiling"] = {"value": nb / ceiling["duplex"] / 1e9, "unit": UNIT,
                                   "h2d_gb_s_per_gpu": head["e2e"]["h2d_bytes_per_step"] / ceiling["h2d"] / 1e9,
                                   "d2h_gb_s_per_gpu": head["e2e"]["d2h_bytes_per_step"] / ceiling["d2h"] / 1e9,
                                   "what": "input GB/s if one step's H2D and D2H bytes move as plain duplex cudaMemcpyAsync (8 chunks "
                                           "each way, pinned), all ranks at once, max over ranks, best of 5"}
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/u32 integer", "data": "synthetic",
                "config": head["config"],
                "method": {"l2": "NOT flushed (diagnostic run)" if cx.no_flush else "flushed between steps (512 MiB memset)",
                           "timing": "per-step CUDA events on the launching stream around the step (one graph launch = the pass's kernel nodes), max over ranks", "host_numa": numa,
                           "tokens_per_gpu": head["tokens_per_gpu"]},
                "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"],
                "e2e": e2e,
                "gpu_launches": args.steps * head["launches_per_step"], "gpu_launches_e2e": args.steps * head["launches_per_call"],
                "clocks": clocks,
                "ids_match_cpu_baseline": head["ids_match_cpu_baseline"],
                "configs": configs or None, "strong": strong, "python_api": py_api, "small_batch": small}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

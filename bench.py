#!/usr/bin/env python3
"""bench.py -- the openclaw-governance hot path on B200: messages scanned / s.

One *step* = one pass of the scan hot path (prefilter scan + exact verification + result words) over
one batch of synthetic messages.  Workload at N=1: BASELINE.json configs[1] -- 1 M synthetic
256-byte messages x 500-rule firewall scan; for N>1 every rank scans its own 1 M-message shard
(weak scaling; no data-path collective for the scan).  The JSON line also carries
  roofline     : achieved algorithmic HBM GB/s of the scan kernel (n*(L+12) bytes per launch,
                 SURVEY 8(d)) against MEASURED_PEAKS.json, measured live with CUDA events
  cpu_baseline : the CPU oracle port (oracle/, kind "port") on the box's host cores, bounded sample
  e2e          : the same metric through the C ABI with HOST (pinned) buffers, copies inside
  extra.merkle : Proof-of-Guardrails Merkle leaves/s (C3: 16 Mi x 256 B leaves per rank), block roots
                 all-gathered over NCCL and folded (the one collective on this path)
`--impl reference` times the reference path's CPU restatement (Node.js is not in this image or on
the GPU box, and the reference has no native sources to compile, so oracle/ is the only runnable
form of it) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MSGS = 1 << 20
MSG_LEN = 256
N_RULES = 500
P_HIT = 0.01
MERKLE_LEAVES = 1 << 24
MERKLE_LEAF = 256
MERKLE_BLOCK_LOG2 = 16


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML poll every ~0.5 ms between
    start (construction) and stop() -- the region is only a few milliseconds long, an `nvidia-smi -lms` child would
    mostly report the idle GPU after it.  Falls back to one nvidia-smi query if NVML is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80))

    def __init__(self, index: int):
        import threading
        self.samples, self.reasons, self.max_mhz, self._stop, self.t = [], set(), None, False, None
        self.index = index
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            try:
                pr = torch.cuda.get_device_properties(index)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)   # first calls are slow: not inside the region

            def loop():
                k = 0
                while not self._stop:
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        if k % 3 == 0:
                            r = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                            for name, bit in self.REASONS:
                                if r & bit:
                                    self.reasons.add(name)
                    except Exception:
                        pass
                    k += 1
                    time.sleep(0.0003)
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception:
            self.t = None

    def stop(self):
        if self.t is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi after the region (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        self._stop = True
        self.t.join(timeout=2)
        sm = self.samples
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "NVML polled inside the timed region"}


def reference_arm(args, rank, world, emit=None):
    """The reference's own CPU implementation of the path, restated (oracle/): policy-semantics scan
    (matchesAny: RegExp.test per rule, gov/src/conditions/context.ts:9-25) on all host threads."""
    if rank != 0:
        return
    from oracle import oracle as O
    from vainplex_openclaw_b200 import workload as W
    rl = W.make_rules(N_RULES)
    regs = [O.Regex(r["source"], "i" if r["flags"] else "") for r in rl]
    threads = os.cpu_count() or 1
    # bounded sample per step, sized from a short calibration so that the whole --steps K --warmup W run stays near two
    # minutes whatever K and W the caller picks (CG_REF_SAMPLE pins it)
    cal_n = 4096
    data_t, off_t, _ = W.make_messages(65536, MSG_LEN, rl, p_hit=P_HIT)
    data, off = data_t.numpy(), off_t.numpy().astype(np.uint64)
    O.scan_policy(regs, data[: cal_n * MSG_LEN + 64], off[:cal_n + 1], threads=threads, want_bits=False)       # untimed: page in, spawn threads
    t0 = time.perf_counter()
    O.scan_policy(regs, data[: cal_n * MSG_LEN + 64], off[:cal_n + 1], threads=threads, want_bits=False)
    rate = cal_n / max(time.perf_counter() - t0, 1e-6)
    budget_s = float(os.environ.get("CG_REF_SECONDS", 120.0))
    sample = int(os.environ.get("CG_REF_SAMPLE", 0)) or int(rate * budget_s / max(args.steps + args.warmup, 1))
    sample = max(1024, min(65536, sample))
    data, off = data[: sample * MSG_LEN + 64], off[: sample + 1]
    for _ in range(args.warmup):
        O.scan_policy(regs, data, off, threads=threads, want_bits=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.scan_policy(regs, data, off, threads=threads, want_bits=False)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"metric": "messages_scanned_per_s", "value": v, "unit": "msgs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "1M x 256B msgs x 500 rules (C2), policy semantics; bounded sample of %d msgs per step" % sample,
                       "rules": N_RULES, "msg_len": MSG_LEN},
            "cpu_baseline": {"value": v, "unit": "msgs/s", "cores": threads, "kind": "port",
                             "sample": "%d messages x %d rules per step, oracle/jsre.c backtracking matcher, %d threads (Node.js absent: restated oracle, not Node)" % (sample, N_RULES, threads)},
            "e2e": {"value": v, "unit": "msgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    (emit or (lambda o: print(json.dumps(o))))(line)


def main():
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner, ...)
    # goes to stderr instead; the real stdout is restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--msgs", type=int, default=N_MSGS)
    ap.add_argument("--len", type=int, default=MSG_LEN, dest="msg_len")
    ap.add_argument("--rules", type=int, default=N_RULES)
    ap.add_argument("--stride", type=int, default=0, help="gram filter stride: 0 = the compiler's choice, 2 or 4")
    ap.add_argument("--p-hit", type=float, default=None, help="fraction of messages with an injected rule-matching token (default 0.01; SURVEY 8d also names 0 and 0.10)")
    ap.add_argument("--seed-offset", type=int, default=0, help="shift the synthetic data seed (rank r of an N-GPU run uses offset r)")
    ap.add_argument("--no-merkle", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--merkle-leaves", type=int, default=MERKLE_LEAVES)
    args = ap.parse_args()
    global P_HIT
    if args.p_hit is not None:
        P_HIT = args.p_hit
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return reference_arm(args, rank, world, emit)

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    from vainplex_openclaw_b200 import _native as N, workload as W
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.init(local)
    n, L, R = args.msgs, args.msg_len, args.rules
    rl = W.make_rules(R)
    rules = W.rules_as_tuples(rl)
    t0 = time.perf_counter()
    rs = N.Ruleset(rules, options=args.stride, strict=True)
    compile_s = time.perf_counter() - t0
    info = rs.info()
    data, off64, inj = W.make_messages(n, L, rl, p_hit=P_HIT, seed=W.SEED_MSG + rank + args.seed_offset, device=dev,
                                        vocab_seed=W.SEED_MSG + args.seed_offset)
    off = off64.to(torch.int32)              # uint32 offsets (bit pattern) as the C ABI expects; n*L < 2^31 here
    # two result buffers: the library keeps two batches in flight (batch k's verify overlaps batch k+1's scan), so a
    # buffer may only be reused two calls later
    words2 = [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(2)]
    words = words2[0]
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream: handle 0 would mean "library stream" to the C ABI
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)

    step_no = [0]

    def step():
        rs.scan_batch_device(data.data_ptr(), off.data_ptr(), n, words2[step_no[0] & 1].data_ptr(), stream.cuda_stream)
        step_no[0] += 1

    for _ in range(args.warmup):
        step()
    rs.scan_join(stream.cuda_stream)
    torch.cuda.synchronize()
    # kernel breakdown (profiling events inside the library, one sync per step; not the headline timing)
    N.set_profiling(True)
    kms = []
    for _ in range(min(args.steps, 20)):
        step(); torch.cuda.synchronize(); kms.append(N.last_kernel_ms())
    N.set_profiling(False)
    kms = np.array(kms)
    rs.scan_join(stream.cuda_stream)
    counters = rs.work_counters()
    words = words2[(step_no[0] - 1) & 1]             # results of the last (sequential, profiled) step: checked against the e2e path below
    words_seq = words.clone()

    # ---- headline: K steps, device-resident inputs (256 MiB per step > 126 MB L2), barrier + sync both sides
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    rs.scan_join(stream.cuda_stream)                 # waits for the stream; raises if a batch overflowed a queue
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    pipelined_equal = bool(torch.equal(words2[0], words_seq) and torch.equal(words2[1], words_seq))
    launches = N.launch_count() - l0
    per_rank_ms = [ms / args.steps]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        allt = torch.zeros(world, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allt, t)
        per_rank_ms = [float(x) / args.steps for x in allt.tolist()]
        ms = float(allt.max().item())
        dist.barrier()
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- e2e through the host-buffer C ABI call: pinned host input -> H2D -> kernels -> D2H words
    h_data = torch.empty(data.numel(), dtype=torch.uint8, pin_memory=True); h_data.copy_(data)
    h_off = torch.empty(n + 1, dtype=torch.int32, pin_memory=True); h_off.copy_(off)
    h_words = torch.empty(n, dtype=torch.int64, pin_memory=True)
    import ctypes as C
    lib = N.load()
    nh = C.c_uint32(0)

    def e2e_step():
        N.check(lib.cg_scan_batch(rs.handle, h_data.data_ptr(), h_off.data_ptr(), n, h_words.data_ptr(), None, 0, C.byref(nh)))

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * e2e_steps / e2e_s
    same = bool(torch.equal(h_words.to(dev), words))
    # C1: one message through the blocking single-message entry point (the synchronous hooks' path), 17 built-in rules
    one = None
    if world == 1:
        rs17 = N.Ruleset(W.rules_as_tuples(W.make_rules(17)), strict=True)
        m1 = bytes(h_data.numpy()[:L])
        tok = b"please use key sk-" + b"a" * 24 + b" for bob@example.com"
        for _ in range(20):
            rs17.scan_one(m1)
        t0 = time.perf_counter()
        for _ in range(200):
            rs17.scan_one(m1)
        lat = (time.perf_counter() - t0) / 200
        w1, r1 = rs17.scan_one(tok)
        one = {"latency_us": lat * 1e6, "rules": 17, "msg_bytes": L, "hit_rules_of_probe": r1}
        rs17.close()
    # (f1) redacted output through cg_redact_batch (findMatches semantics + digests + splice on the device), host buffers
    redact = None
    if world == 1:
        nr = min(n, 1 << 18)
        hd, ho = h_data.numpy()[: nr * L + 64], h_off.numpy()[: nr + 1].view(np.uint32)
        rs.redact_batch(hd, ho)
        t0 = time.perf_counter()
        r_out, r_off, r_spans, r_dig = rs.redact_batch(hd, ho)
        dt = time.perf_counter() - t0
        redact = {"msgs": nr, "msgs_per_s_from_host": nr / dt, "spans": int(len(r_spans)), "out_bytes": int(len(r_out)), "in_bytes": int(nr * L)}

    # ---- Merkle (C3): leaves/s; block roots all-gathered over NCCL, folded on every rank
    merkle = None
    if not args.no_merkle:
        nl = args.merkle_leaves
        leaves = W.make_leaves(nl, MERKLE_LEAF, seed=W.SEED_LEAVES + rank, device=dev)
        nblk = (nl + (1 << MERKLE_BLOCK_LOG2) - 1) >> MERKLE_BLOCK_LOG2
        roots = torch.zeros(nblk * 32, dtype=torch.uint8, device=dev)
        allroots = torch.zeros(world * nblk * 32, dtype=torch.uint8, device=dev)
        root = torch.zeros(32, dtype=torch.uint8, device=dev)

        def merkle_step():
            N.check(lib.cg_merkle_block_roots_device(leaves.data_ptr(), MERKLE_LEAF, nl, MERKLE_BLOCK_LOG2, roots.data_ptr(), stream.cuda_stream))
            if world > 1:
                dist.all_gather_into_tensor(allroots, roots)
            else:
                allroots.copy_(roots)
            N.check(lib.cg_merkle_fold_device(allroots.data_ptr(), world * nblk, root.data_ptr(), stream.cuda_stream))

        for _ in range(2):
            merkle_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        msteps = 5
        m0.record(stream)
        for _ in range(msteps):
            merkle_step()
        m1.record(stream)
        torch.cuda.synchronize()
        mms = m0.elapsed_time(m1) / msteps
        if world > 1:
            t = torch.tensor([mms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mms = float(t.item())
        peak, _ = measured_peaks()
        merkle = {"leaves_per_s": world * nl / (mms * 1e-3), "ms_per_tree": mms, "leaves_per_rank": nl, "leaf_bytes": MERKLE_LEAF,
                  "root_hex": bytes(root.cpu().numpy()).hex(), "collective": "all_gather of %d block roots/rank (NCCL)" % nblk if world > 1 else "none (1 rank)",
                  "hbm_frac_of_measured": (nl * MERKLE_LEAF / (mms * 1e-3)) / (peak * 1e9),
                  "sha256_compressions_per_s": world * (nl * ((MERKLE_LEAF + 1 + 9 + 63) // 64) + 2 * (nl - 1)) / (mms * 1e-3)}
        # (f2) the same leaves through the append-only log in 16 appends from host memory (H2D inside), rank 0, N = 1 only
        if world == 1 and rank == 0:
            nlog = min(nl, 1 << 20)
            h_leaves = leaves[: nlog * MERKLE_LEAF].cpu().numpy()
            offs = (np.arange(nlog // 16 + 1, dtype=np.uint64) * MERKLE_LEAF)
            lg = N.MerkleLog(keep_leaf_digests=True)
            t0 = time.perf_counter()
            for c in range(16):
                part = h_leaves[c * (nlog // 16) * MERKLE_LEAF:]
                N.check(lib.cg_merkle_log_append(lg.handle, part.ctypes.data, offs.ctypes.data, nlog // 16))
            log_root = lg.root()
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            path = lg.proof(nlog // 3)
            proof_ms = (time.perf_counter() - t1) * 1e3
            one_shot = N.merkle_root_fixed(h_leaves, MERKLE_LEAF, nlog)
            merkle["log"] = {"leaves": nlog, "appends": 16, "leaves_per_s_from_host": nlog / dt, "root_equals_one_shot_tree": log_root == one_shot,
                             "proof_len": len(path), "proof_ms": proof_ms,
                             "proof_verifies": N.merkle_verify_proof(bytes(h_leaves[(nlog // 3) * MERKLE_LEAF:(nlog // 3 + 1) * MERKLE_LEAF]), nlog // 3, nlog, path, log_root)}
            lg.close()
        del leaves

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port) on a bounded sample of the same workload, rank 0, N=1 only
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import oracle as O
        regs = [O.Regex(r["source"], "i" if r["flags"] else "") for r in rl]
        threads = os.cpu_count() or 1
        sample = int(os.environ.get("CG_CPU_SAMPLE", 65536))
        sample = min(sample, n)
        hb = h_data.numpy()[: sample * L + 64]
        ho = np.arange(sample + 1, dtype=np.uint64) * L
        O.scan_policy(regs, hb[: 2048 * L + 64], ho[:2049], threads=threads, want_bits=False)
        t0 = time.perf_counter()
        _, cw = O.scan_policy(regs, hb, ho, threads=threads, want_bits=False)
        dt = time.perf_counter() - t0
        agree = bool(np.array_equal(cw, h_words.numpy()[:sample].view(np.uint64)))
        cpu = {"value": sample / dt, "unit": "msgs/s", "cores": threads, "kind": "port",
               "sample": "first %d messages of the batch x %d rules, oracle/jsre.c (restated oracle, not Node: Node.js absent), %d threads, %.2f s; result words equal to GPU: %s" % (sample, R, threads, dt, agree)}

    peak, peak_src = measured_peaks()
    scan_ms = float(np.median(kms[:, 0])) if len(kms) else None
    alg_bytes = n * (L + 12)
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes/launch from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("scan_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": "messages_scanned_per_s", "value": value, "unit": "msgs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%d x %dB synthetic msgs x %d-rule firewall scan per GPU (BASELINE configs[1]), policy semantics (RegExp.test per rule), p_hit=%.2f" % (n, L, R, P_HIT),
                   "rules": R, "msg_len": L, "msgs_per_gpu": n, "l2": "inputs (%.0f MB/step) larger than the 126 MB L2" % (n * L / 1e6),
                   "prefilter": {"stride": int(info.stride), "gram_keys": int(info.gram_keys), "level1b_entries": int(info.gram_entries),
                                 "factor_len_min": int(info.factor_len) & 0xff, "factor_len_max": int(info.factor_len) >> 8, "factors": int(info.n_factors),
                                 "bitmap_bytes": int(info.bitmap_bytes), "smem_image_bytes": int(info.image_bytes), "tables_resident": bool(info.tables_resident),
                                 "trigger_bytes": int(info.n_triggers), "always_candidate_rules": int(info.n_always_candidate)},
                   "step": "in-order step replayed as one CUDA graph; results of the timed steps equal the profiled step: %s" % pipelined_equal,
                   "compile_s": compile_s},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "kernel": "scan_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": scan_ms},
        "kernel_ms_note": "kernel_ms: CUDA events inside the library around each kernel of one step (profiling mode, no graph); ms_per_step: the graph-replayed steady state",
        "kernel_ms": {"scan": scan_ms, "resolve": float(np.median(kms[:, 1])), "verify": float(np.median(kms[:, 2])), "finalize": float(np.median(kms[:, 3]))},
        "candidates": {"confirmed_factor_occurrences": counters[4], **({"vm_cycle_hist_2^11..": list(counters[8:16]), "vm_cycle_max": counters[7]} if os.environ.get("CG_SCAN_DEBUG") == "2" else {}), "flagged_grams": counters[6], "messages_with_candidates": counters[0], "vm_pairs": counters[1], "flags": counters[3],
                       "hit_messages": int((words != 0).sum().item()), "injected": len(inj)},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "msgs/s", "h2d_bytes_per_step": int(n * L + 4 * (n + 1)), "d2h_bytes_per_step": int(8 * n + 64),
                "steps": e2e_steps, "words_equal_device_path": same},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "extra": {"merkle": merkle, "redact": redact, "scan_one": one, "per_rank_ms_per_step": per_rank_ms},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

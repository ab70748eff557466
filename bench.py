#!/usr/bin/env python
"""bench.py — text-predict BPTT chars/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config text|rnnca|default|classify|multi]

One "step" is one character position of recur's synchronic text-predict loop
(reference charmodel-predict.c:293-311) over all streams of this rank:
advance, one-hot forward, softmax error, truncated BPTT, one weight update.
The workload is BASELINE.json configs[1]: hidden 1023, 512 synchronic streams
per GPU, BPTT depth 30, 42-symbol alphabet, synthetic order-1 Markov text,
random-initialised weights (rnn_randomise_weights_auto, seed 1).

  value  whole-job chars/sec with the text resident in HBM
         (rnn_batch_text_train), CUDA-event timed on the library's stream,
         max over ranks.
  e2e    the same metric through rnn_batch_char_step with HOST symbol arrays:
         every step copies this step's 2*n symbol bytes host->device and reads
         the step's report sums (32 bytes) back.
  roofline / cpu_baseline: see DESIGN.md "Measurement".

N > 1: one process per GPU (torchrun), streams sharded 512 per rank (weak
scaling: `value`), [ih_delta | ho_delta] summed over the ranks each step by
the library's own exchange kernel over NVLink peer memory (NCCL with
--no-p2p).  The line also carries a `strong` sub-record - the SAME 512
streams split 512/N per GPU, as BASELINE.json words configs[1] - and a
`trained` sub-record (the job after 2000 more positions, when the adaptive
BPTT walk has deepened).  The learn rate is LEARN_RATE / N: the reference sums
the deltas over all streams, so N times the streams at the same rate is N
times the step (1e-5 at 512 streams already diverges).  The only
thing torch.distributed carries is the bootstrap (NCCL id, IPC handles), the
barriers and the max-over-ranks of the timings.

--impl reference times the reference's own CPU implementation of the same
loop (oracle/_ref, the unmodified reference compiled in place) on the host
cores: one independent replica per core, bounded sample.

--config selects one of the other BASELINE.json configurations instead (each
with its own --impl reference arm):
  rnnca     configs[4]: the cell automaton at 1920 x 1080, a step = a frame
            (rnn_cells_rnnca_run / _frame); N > 1: ONE automaton, its rows
            shared by the GPUs (strong scaling)
  default   configs[0]: one default net through the per-net API
  classify  configs[2]: 256-channel classify training through the batch calls
  multi     configs[3]: multi-head charmodel forward, 64 texts
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

HIDDEN = 1023
STREAMS = 512
DEPTH = 30
ALPHABET = 42
TEXT_LEN = 2_000_000
LEARN_RATE = 1e-6     # deltas are summed over streams: 1e-5 diverges
MOMENTUM = 0.95
SOFT_START = 2000.0
METRIC = "text-predict BPTT chars/sec"
UNIT = "chars/s"


def markov_text(n, n_symbols, seed):
    """Order-1 Markov chain, vectorised enough for 2M symbols."""
    rng = np.random.RandomState(seed)
    trans = rng.dirichlet(np.ones(n_symbols) * 0.3, size=n_symbols)
    cum = np.cumsum(trans, axis=1)
    cum[:, -1] = 1.0
    u = rng.random_sample(n)
    out = np.empty(n, dtype=np.uint8)
    s = 0
    for i in range(n):
        s = int(np.searchsorted(cum[s], u[i]))
        out[i] = s if s < n_symbols else n_symbols - 1
    return out


def synthetic_text(n=TEXT_LEN, seed=2):
    cache = os.path.join(ROOT, "gpurun_out", "bench_text_%d_%d.npy" % (n, seed))
    try:
        t = np.load(cache)
        if len(t) == n:
            return t
    except Exception:
        pass
    # generate a 200k chunk and tile it with different offsets: the loop only
    # needs a learnable symbol stream, and 2M python iterations take too long
    base = markov_text(200_000, ALPHABET, seed)
    reps = (n + len(base) - 1) // len(base)
    t = np.concatenate([np.roll(base, 7919 * r) for r in range(reps)])[:n].copy()
    try:
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.save(cache, t)
    except Exception:
        pass
    return t


# ---------------------------------------------------------------------------
# clocks sampler

class ClockSampler(threading.Thread):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for k, name in enumerate(names):
                    if s[5 + k].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# the reference arm / CPU baseline: replicas of the reference's own loop

def _ref_replica(args):
    seed, hidden, n_streams, depth, warm, steps, text = args
    import oracle
    from recur_b200 import abi
    from helpers import make_net, u8ptr
    ref = oracle.load_ref(strict=False)   # the reference's own flags
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)
    net = make_net(ref, input_size=ALPHABET, hidden=hidden, output=ALPHABET, depth=depth,
                   seed=seed, lr=LEARN_RATE)
    nets = ref.rnn_new_training_set(net, n_streams)
    if warm:
        ref.ref_multi_tap_train(nets, n_streams, u8ptr(text), len(text), 0, warm,
                                abi.RNN_MOMENTUM_WEIGHTED, MOMENTUM, SOFT_START, None, None, None)
    secs = ref.ref_multi_tap_train(nets, n_streams, u8ptr(text), len(text), warm, steps,
                                   abi.RNN_MOMENTUM_WEIGHTED, MOMENTUM, SOFT_START,
                                   None, None, None)
    return secs, steps * n_streams


def reference_cpu_throughput(cores, n_streams, warm, steps, hidden=HIDDEN, depth=DEPTH):
    """Aggregate chars/s of `cores` independent replicas of the reference's
    multi-tap loop (it is single-threaded by construction: streams alias
    shared delta arrays)."""
    import oracle
    if not oracle.have_ref():
        raise RuntimeError("oracle/_ref is not built")
    text = synthetic_text(200_000)
    jobs = [(100 + r, hidden, n_streams, depth, warm, steps, text) for r in range(cores)]
    t0 = time.time()
    if cores == 1:
        res = [_ref_replica(jobs[0])]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            res = pool.map(_ref_replica, jobs)
    wall = time.time() - t0
    # every replica ran concurrently: aggregate = sum of per-replica rates
    rate = sum(chars / secs for secs, chars in res)
    return rate, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    # bounded sample: 8 streams per replica; 30 warm-up positions fill the
    # BPTT ring (steady-state cost), then a few timed positions per "step"
    n_streams, warm = 8, DEPTH
    per_step = 4
    t_all = []
    total_chars = 0
    t0 = time.time()
    steps_done = 0
    timed = max(4, min(per_step * args.steps, 96))   # bounded: ~10-20 s of CPU work
    rate, wall = reference_cpu_throughput(cores, n_streams, warm, timed)
    ms_per_step = 1e3 * (per_step * n_streams * cores) / rate
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d independent replicas (one per core) of the reference "
                                   "multi-tap loop, H%d D%d, %d streams each, %d warm-up + %d timed "
                                   "positions, %.1f s wall" % (cores, HIDDEN, DEPTH, n_streams,
                                                  warm, timed, wall)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus, exchange=None):
    if n_gpus > 1:
        par = ("streams sharded over %d GPUs, one process each; [ih_delta | ho_delta] summed "
               "over ranks every step by: %s" % (n_gpus, exchange or "see gradient_exchange"))
    else:
        par = "1 GPU"
    return {
        "workload": "text-predict BPTT training, hidden %d, %d synchronic streams per GPU, "
                    "BPTT depth %d, %d-symbol alphabet (BASELINE.json configs[1])"
                    % (HIDDEN, STREAMS, DEPTH, ALPHABET),
        "hidden": HIDDEN, "streams_per_gpu": STREAMS, "global_streams": STREAMS * n_gpus,
        "bptt_depth": DEPTH, "alphabet": ALPHABET,
        "learn_rate": LEARN_RATE / n_gpus,
        "learn_rate_note": "%g / n_gpus: deltas are summed over all streams of all ranks"
                           % LEARN_RATE,
        "learning_style": "weighted momentum 0.95", "text": "order-1 Markov chain, seed 2",
        "parallelism": par,
        "l2": "working set per step (history ring 66 MB + error chain 66 MB + their FP16 "
              "operand planes + weights, momentum, deltas 18 MB) exceeds the 126 MB L2; "
              "no explicit flush",
    }


# ---------------------------------------------------------------------------
# our arm

def run_ours(args):
    import torch
    from recur_b200 import api, abi
    from helpers import make_net, u8ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    L = api.load_library()
    if L.rnn_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if L.rnn_b200_set_device(local_rank) != 0:
        raise SystemExit("cannot select GPU %d" % local_rank)
    from recur_b200 import dist as rdist
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        rdist.join_comm(L, dist, rank, world, device="cuda")

    def barrier():
        L.rnn_b200_synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        return rdist.max_over_ranks(dist, x, device="cuda")

    if args.engine is not None:
        L.rnn_b200_set_engine(args.engine)
    n = args.streams
    text = synthetic_text()
    # each rank reads its own stretch of the text
    lo, hi = rdist.shard_bounds(len(text), rank, world)
    my_text = np.ascontiguousarray(text[lo:hi])

    def make_job(n_streams=None, lr=None):
        """A freshly initialised net (same seed every time), its training set
        and batch, the peer exchange attached, the text in HBM."""
        n_streams = n_streams or n
        devnull_fd = os.dup(2)
        os.dup2(os.open(os.devnull, os.O_WRONLY), 2)   # the reference-style init chatter
        net = make_net(L, input_size=ALPHABET, hidden=args.hidden, output=ALPHABET, depth=DEPTH,
                       seed=1, lr=lr or LEARN_RATE / world)
        os.dup2(devnull_fd, 2)
        os.close(devnull_fd)
        nets = L.rnn_new_training_set(net, n_streams)
        batch = L.rnn_batch_new(nets, n_streams)
        exchange = attach_exchange(batch)
        L.rnn_batch_text_upload(batch, u8ptr(my_text), len(my_text))
        return net, nets, batch, exchange, n_streams

    def drop_job(job):
        net, nets, batch, _, n_streams = job
        L.rnn_batch_delete(batch)
        L.rnn_delete_training_set(nets, n_streams, 0)   # nets[0] is the prototype: it goes too

    def time_resident(batch, pos, steps, stats=None):
        """`steps` positions of rnn_batch_text_train between CUDA events on the
        library's stream: (ms, max over ranks; next position)."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        pos = L.rnn_batch_text_train(batch, pos, steps, style, MOMENTUM, SOFT_START,
                                     C.byref(stats) if stats is not None else None)
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)), pos

    def executed_depths(batch, n_streams):
        d = (C.c_int32 * n_streams)()
        L.rnn_batch_bptt_depths(batch, d)
        return np.array(list(d), dtype=np.float64)

    def attach_exchange(batch):
      exchange = "none"
      if world > 1:
          exchange = "nccl all-reduce"
          if not args.no_p2p:
              # fused split-K reduction + all-reduce over NVLink peer memory
              hb = (C.c_uint8 * 192)()
              ok = L.rnn_batch_p2p_export(batch, hb) == 0
              mine = torch.tensor(list(hb), dtype=torch.uint8, device="cuda")
              allh = [torch.zeros_like(mine) for _ in range(world)]
              dist.all_gather(allh, mine)
              flag = torch.tensor([1 if ok else 0], device="cuda")
              dist.all_reduce(flag, op=dist.ReduceOp.MIN)
              if int(flag.item()):
                  raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
                  buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
                  ok = L.rnn_batch_p2p_attach(batch, buf, rank, world) == 0
                  flag = torch.tensor([1 if ok else 0], device="cuda")
                  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                  if int(flag.item()):
                      exchange = "fused reduce + all-reduce kernel over NVLink peer memory"
                  else:
                      raise SystemExit("peer attach succeeded on some ranks only")
      return exchange

    stream = torch.cuda.ExternalStream(L.rnn_b200_stream())
    style = abi.RNN_MOMENTUM_WEIGHTED

    def warm_up(batch):
        pos = L.rnn_batch_text_train(batch, 0, max(args.warmup, 3), style, MOMENTUM, SOFT_START,
                                     None)
        # fill the BPTT ring so that every timed step walks the full depth
        if args.warmup < DEPTH and not args.cold:
            pos = L.rnn_batch_text_train(batch, pos, DEPTH - args.warmup, style, MOMENTUM,
                                         SOFT_START, None)
        return pos

    # ---- device-resident arm (value) ---------------------------------------
    job = make_job()
    net, nets, batch, exchange, _ = job
    pos = warm_up(batch)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.samples and time.time() - t_wait < 5.0:
        time.sleep(0.05)          # nvidia-smi takes a moment to print its first line
    sampler.samples.clear()
    launches0 = L.rnn_b200_kernel_launches()
    stats = api.RnnBatchCharStats()
    ms, pos = time_resident(batch, pos, args.steps, stats)
    launches = L.rnn_b200_kernel_launches() - launches0
    value = (args.steps * n * world) / (ms * 1e-3)

    # ---- the same job once training has gone on: the adaptive walk deepens ----
    trained = None
    if args.trained_after > 0:
        pos = L.rnn_batch_text_train(batch, pos, args.trained_after, style, MOMENTUM, SOFT_START,
                                     None)
        t_steps = min(args.steps, 300)
        tstats = api.RnnBatchCharStats()
        tms, pos = time_resident(batch, pos, t_steps, tstats)
        td = executed_depths(batch, n)
        trained = {"positions_before": DEPTH + args.steps + args.trained_after,
                   "steps": t_steps, "value": (t_steps * n * world) / (tms * 1e-3), "unit": UNIT,
                   "ms_per_step": tms / t_steps, "mean_executed_depth": float(td.mean()),
                   "max_executed_depth": float(td.max()),
                   "t_entropy": -tstats.entropy / max(tstats.count, 1)}

    # ---- end-to-end arm (host symbols in, report sums out, every step) -----
    # the SAME job again from the same initial weights, so that both arms do
    # the same work (the BPTT walk deepens as training goes on)
    barrier()
    drop_job(job)
    job = make_job()
    net, nets, batch, exchange, _ = job
    pos = warm_up(batch)
    spacing = (len(my_text) - 1) // n
    offs = (np.arange(n, dtype=np.int64) * spacing)
    est = api.RnnBatchCharStats()
    e2e_steps = args.steps
    # the caller's symbol arrays live in host memory before the clock starts;
    # every timed step hands one row pair to the library (H2D inside the call)
    idx = (pos + np.arange(e2e_steps, dtype=np.int64)[:, None] + offs[None, :]) % (len(my_text) - 1)
    cur_all = np.ascontiguousarray(my_text[idx])
    nxt_all = np.ascontiguousarray(my_text[idx + 1])
    cur_base, nxt_base = cur_all.ctypes.data, nxt_all.ctypes.data
    step_fn = L.rnn_batch_char_step
    est_ref = C.byref(est)
    # argument marshalling (ctypes pointer objects for each step's rows) is the
    # harness's cost, not the call's: done before the clock starts
    cur_ptrs = [C.cast(cur_base + k * n, abi.u8_p) for k in range(e2e_steps)]
    nxt_ptrs = [C.cast(nxt_base + k * n, abi.u8_p) for k in range(e2e_steps)]
    soft_start = L.rnn_calculate_momentum_soft_start
    gen0 = int(net.contents.generation)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        m = soft_start(float(gen0 + k), MOMENTUM, SOFT_START)
        step_fn(batch, cur_ptrs[k], nxt_ptrs[k], style, m, est_ref)
    L.rnn_b200_synchronize()
    t1 = time.perf_counter()
    barrier()
    e2e_ms = max_over_ranks((t1 - t0) * 1e3)
    e2e_value = (e2e_steps * n * world) / (e2e_ms * 1e-3)
    sampler.stop()

    # ---- forward only: rnn_opinion stream-steps/s (the metric's second half) --
    fwd_steps = min(args.steps, 300)
    L.rnn_batch_text_forward(batch, pos, 10)
    barrier()
    fev0, fev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fev0.record(stream)
    L.rnn_batch_text_forward(batch, pos, fwd_steps)
    fev1.record(stream)
    barrier()
    fwd_ms = max_over_ranks(fev0.elapsed_time(fev1))
    opinion_rate = (fwd_steps * n * world) / (fwd_ms * 1e-3)

    # ---- per-kernel timing for the roofline (separate, instrumented pass) --
    prof_steps = min(args.steps, 50)
    L.rnn_b200_profile_enable(1)
    pev0, pev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pev0.record(stream)
    L.rnn_batch_text_train(batch, pos, prof_steps, style, MOMENTUM, SOFT_START, None)
    pev1.record(stream)
    pms = (C.c_double * 8)()
    pln = (C.c_uint64 * 8)()
    ncls = L.rnn_b200_profile_read(pms, pln, 8)
    L.rnn_b200_profile_enable(0)
    prof_total = pev0.elapsed_time(pev1)
    depths = (C.c_int32 * n)()
    L.rnn_batch_bptt_depths(batch, depths)
    depths = np.array(list(depths), dtype=np.float64)
    mean_depth = float(depths.mean())
    max_depth = float(depths.max())
    # executed BPTT depth per stream (n_steps of the last step), from the log scalars
    kernels = {}
    for c in range(ncls):
        name = L.rnn_b200_profile_class_name(c).decode()
        if pln[c]:
            kernels[name] = {"ms_total": pms[c], "launches": int(pln[c]),
                             "ms_per_launch": pms[c] / pln[c],
                             "share_of_step": pms[c] / prof_total}
    hs1 = args.hidden + 1
    i_alg = hs1 + 1               # bias + hidden rows + the one hot input row
    flops_pair = 2.0 * i_alg * hs1  # one stream, one ring row: one of {error back, outer product}
    # algorithmic work counts the BPTT steps each stream really executed (the
    # adaptive early exit of recur-nn.c:387 is part of the algorithm, and the
    # CPU reference takes the same exit), not the nominal depth
    alg = {
        "forward": (2.0 * i_alg * hs1) * n,                  # per launch (one step)
        "bptt_chain": flops_pair * n * mean_depth,           # per launch (the whole walk)
        "weight_grad": flops_pair * n * mean_depth,          # per launch (all executed steps)
    }
    chain_launches_per_step = kernels.get("bptt_chain", {}).get("launches", prof_steps) / prof_steps
    if chain_launches_per_step > 1.5:   # one launch per BPTT step (per-step kernels)
        alg["bptt_chain"] = flops_pair * n
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_f16 = float(peaks["bf16_tflops_sustained"])
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (dense 16-bit MMA rate, sustained)"
    except Exception:
        peak_f16 = 1400.0
        peak_src = "fallback 1.4 PFLOP/s sustained dense 16-bit MMA"
    dominant = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_total"],
                   default=None)
    traffic = None
    try:   # DRAM bytes per launch of that kernel from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")))
        kname = {"bptt_chain": "k_tc_chain_persistent", "weight_grad": "k_tc_dw_pair",
                 "forward": "k_tc_nt"}.get(dominant)
        for key in tr:   # ncu prints template arguments after the name
            if kname and key.startswith(kname):
                traffic = tr[key]["dram_bytes_per_launch"]
                break
    except Exception:
        pass
    roofline = None
    if dominant:
        per_launch_s = kernels[dominant]["ms_per_launch"] * 1e-3
        achieved = alg[dominant] / per_launch_s / 1e12
        roofline = {"bound": "tensor", "kernel": dominant, "achieved": achieved, "peak": peak_f16,
                    "unit": "TFLOP/s", "frac": achieved / peak_f16, "traffic": traffic,
                    "peak_source": peak_src,
                    "frac_vs_tf32_peak": achieved / (peak_f16 / 2.0),
                    "note": "achieved = algorithmic FP32 FLOPs (unpadded, 2 per multiply-add, "
                            "only the BPTT steps each stream executed) per launch / CUDA-event "
                            "time per launch.  The kernels issue kind::f16 MMAs on FP16 hi/lo "
                            "operand planes, three per logical product, so the ceiling against "
                            "`peak` is 1/3.  Round 1 issued kind::tf32 MMAs and reported against "
                            "peak / 2 (the TF32 rate the north star names): frac_vs_tf32_peak "
                            "continues that series",
                    "kernels": kernels}

    # ---- the exchange alone (no skew from unequal BPTT depths) ---------------
    exchange_us = None
    if world > 1 and exchange and exchange.startswith("fused"):
        barrier()
        us = float(L.rnn_batch_p2p_probe(batch, 200))
        exchange_us = max_over_ranks(us) if us > 0 else None
        barrier()

    # ---- strong scaling: the SAME 512 streams split over the GPUs ------------
    strong = None
    if world > 1 and n == STREAMS and (STREAMS // world) >= 64 and not args.no_strong:
        barrier()
        drop_job(job)
        n_strong = STREAMS // world
        job = make_job(n_strong, LEARN_RATE)   # 512 streams in all: the N = 1 job's rate
        net, nets, batch, exchange, _ = job
        spos = warm_up(batch)
        s_steps = min(args.steps, 500)
        sms, spos = time_resident(batch, spos, s_steps)
        sd = executed_depths(batch, n_strong)
        strong = {"scaling": "strong", "global_streams": STREAMS, "streams_per_gpu": n_strong,
                  "steps": s_steps, "value": (s_steps * STREAMS) / (sms * 1e-3), "unit": UNIT,
                  "ms_per_step": sms / s_steps, "mean_executed_depth": float(sd.mean()),
                  "learn_rate": LEARN_RATE}

    line = None
    if rank == 0:
        step_flops = (alg["forward"] + 2 * flops_pair * n * mean_depth) * world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, exchange),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n,
                    "d2h_bytes_per_step": C.sizeof(api.RnnBatchCharStats),
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "algorithmic_tflops": step_flops / (ms / args.steps * 1e-3) / 1e12,
            "bptt": {"nominal_depth": DEPTH, "mean_executed_depth": mean_depth,
                     "max_executed_depth": max_depth,
                     "mflop_per_stream_char": (alg["forward"] / n + 2 * flops_pair * mean_depth) / 1e6},
            "train": {"t_entropy": -stats.entropy / max(stats.count, 1),
                      "accuracy": stats.correct / max(stats.count, 1)},
            "engine": {0: "auto", 1: "fma", 2: "tensor"}[L.rnn_b200_set_engine(-1)],
            "gradient_exchange": exchange,
            "gradient_exchange_us": exchange_us,
            "trained": trained,
            "strong": strong,
            "opinion": {"value": opinion_rate, "unit": "stream-steps/s (rnn_opinion only, "
                        "one-hot input, %d steps)" % fwd_steps,
                        "ms_per_step": fwd_ms / fwd_steps},
        }
        if args.hidden != HIDDEN or n != STREAMS:
            line["config"]["workload"] += " [OVERRIDDEN: hidden %d streams %d]" % (args.hidden, n)
    drop_job(job)
    if dist:
        barrier()
    L.rnn_b200_comm_leave()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = host_cores()
                one, wall1 = reference_cpu_throughput(1, 8, DEPTH, 48)
                rate, wall = reference_cpu_throughput(cores, 8, DEPTH, 96)
                line["cpu_baseline"] = {
                    "value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                    "one_core": one,
                    "sample": "%d independent replicas (one per core) of the reference multi-tap "
                              "loop, H%d D%d, 8 streams each, %d warm-up + 96 timed positions, "
                              "%.1f s wall; one_core: a single replica alone on the host, "
                              "%d + 48 positions, %.1f s (the reference is single-threaded: "
                              "streams alias shared delta arrays)"
                              % (cores, HIDDEN, DEPTH, DEPTH, wall, DEPTH, wall1)}
            except Exception as e:  # the oracle .so did not travel
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0,
                                        "kind": "reference", "sample": "unavailable: %s" % e}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------
# --config rnnca: BASELINE.json configs[4], the cellular automaton at 1080p

RNNCA_W, RNNCA_H = 1920, 1080
RNNCA_HIDDEN, RNNCA_LEN_POS, RNNCA_EDGES = 51, 2, 0
RNNCA_METRIC = "rnnca_frames_per_second"
RNNCA_UNIT = "frames/s"


def rnnca_pattern():
    """17 luma + 8 chroma neighbours, the size of RNNCA_DEFAULT_PATTERN
    (gstrnnca.h:49-51)."""
    off_y = np.array([(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3)
                      if abs(dx) + abs(dy) <= 2 or (abs(dx), abs(dy)) == (2, 2)][:17],
                     dtype=np.int32)
    off_c = np.array([(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)],
                     dtype=np.int32)
    return off_y, off_c


def rnnca_config(n_gpus):
    return {"workload": "rnnca fill_frame (gstrnnca.c:805-830), %dx%d cells, one I35/H%d/O3 ReLU "
                        "net per pixel sharing the trainers' weights; a step = one frame of the "
                        "automaton (every cell: gather 17 luma + 2x8 chroma neighbours + 2 position "
                        "terms, rnn_opinion, fast_sigmoid, bytes); BASELINE.json configs[4]"
                        % (RNNCA_W, RNNCA_H, RNNCA_HIDDEN),
            "cells": RNNCA_W * RNNCA_H, "n_gpus": n_gpus,
            "parallelism": "one automaton, rows sharded over %d GPUs, the frame's bands "
                           "all-gathered (NCCL) after every frame" % n_gpus if n_gpus > 1
                           else "single GPU",
            "cache": "per-frame working set 1.1 GB of hidden state > 126 MB L2: no flush needed"}


def _rnnca_ref_band(args):
    seed, first, n_cells, frames, frame0 = args
    import oracle
    from recur_b200 import abi
    from helpers import make_net
    ref = oracle.load_ref(strict=False)
    port = oracle.load_port()
    os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
    off_y, off_c = rnnca_pattern()
    n_in = len(off_y) + 2 * len(off_c) + RNNCA_LEN_POS
    net = make_net(ref, input_size=n_in, hidden=RNNCA_HIDDEN, output=3, depth=10, seed=seed, lr=3e-3)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    clones = (abi.RecurNN_p * n_cells)(*[ref.rnn_clone(net, fwd, abi.RECUR_RNG_SUBSEED, None)
                                         for _ in range(n_cells)])
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    frame = frame0.copy()
    out = frame0.copy()
    fill = C.cast(port.oracle_rnnca_fill_inputs, C.c_void_p)
    secs = 0.0
    for f in range(frames):
        secs += ref.ref_rnnca_cells(clones, first, n_cells, frame.ctypes.data_as(u8p),
                                    out.ctypes.data_as(u8p), RNNCA_W, RNNCA_H, fill,
                                    off_y.ctypes.data_as(ip), len(off_y), off_c.ctypes.data_as(ip),
                                    len(off_c), RNNCA_LEN_POS, RNNCA_EDGES)
        frame, out = out, frame
    return secs, n_cells * frames


def rnnca_reference_throughput(cores, rows_per_core, frames):
    """frames/s of the reference's fill_frame with its cells split into bands
    over `cores` processes (the element itself is single-threaded; the cells
    of one frame are independent, so this is the most the host could do)."""
    import oracle
    if not oracle.have_ref():
        raise RuntimeError("oracle/_ref is not built")
    rs = np.random.RandomState(3)
    frame0 = rs.randint(0, 256, size=3 * RNNCA_W * RNNCA_H).astype(np.uint8)
    n_cells = rows_per_core * RNNCA_W
    jobs = [(11, (r * rows_per_core % (RNNCA_H - rows_per_core)) * RNNCA_W, n_cells, frames, frame0)
            for r in range(cores)]
    t0 = time.time()
    if cores == 1:
        res = [_rnnca_ref_band(jobs[0])]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            res = pool.map(_rnnca_ref_band, jobs)
    wall = time.time() - t0
    cells_per_s = sum(c / s for s, c in res)
    return cells_per_s / (RNNCA_W * RNNCA_H), wall


def run_rnnca_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cores = host_cores()
    rows, frames = 64, 12
    rate, wall = rnnca_reference_throughput(cores, rows, frames)
    line = {"impl": "reference", "metric": RNNCA_METRIC, "value": rate, "unit": RNNCA_UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": rnnca_config(args.gpus),
            "cpu_baseline": {"value": rate, "unit": RNNCA_UNIT, "cores": cores, "kind": "reference",
                             "sample": "%d processes, each a band of %d rows x %d cells of the "
                                       "1080p frame for %d frames: reference rnn_opinion + "
                                       "fast_sigmoid_array around the oracle's fill_net_inputs "
                                       "(gstrnnca.c itself needs GStreamer); %.1f s wall"
                                       % (cores, rows, RNNCA_W, frames, wall)},
            "e2e": {"value": rate, "unit": RNNCA_UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_rnnca(args):
    import torch
    from recur_b200 import api
    from helpers import make_net

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    L = api.load_library()
    if L.rnn_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if L.rnn_b200_set_device(local_rank) != 0:
        raise SystemExit("cannot select GPU %d" % local_rank)
    from recur_b200 import dist as rdist
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        rdist.join_comm(L, dist, rank, world, device="cuda")

    def barrier():
        L.rnn_b200_synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    off_y, off_c = rnnca_pattern()
    len_y, len_c = len(off_y), len(off_c)
    n_in = len_y + 2 * len_c + RNNCA_LEN_POS
    n = RNNCA_W * RNNCA_H
    net = make_net(L, input_size=n_in, hidden=RNNCA_HIDDEN, output=3, depth=10, seed=11, lr=3e-3)
    # N > 1: ONE automaton, its rows shared between the GPUs (strong scaling)
    cells = (L.rnn_cells_new_sharded if world > 1 else L.rnn_cells_new)(net, RNNCA_W, RNNCA_H)
    if not cells:
        raise SystemExit("rnn_cells_new failed")
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    oy, oc = off_y.ctypes.data_as(ip), off_c.ctypes.data_as(ip)
    rs = np.random.RandomState(3)
    host_in = torch.from_numpy(rs.randint(0, 256, size=3 * n).astype(np.uint8)).pin_memory()
    host_out = torch.empty(3 * n, dtype=torch.uint8).pin_memory()
    pin = C.cast(host_in.data_ptr(), u8p)
    pout = C.cast(host_out.data_ptr(), u8p)
    stream = torch.cuda.ExternalStream(L.rnn_b200_stream())
    warm = max(args.warmup, 3)
    steps = min(args.steps, 400)

    def run(k, first=None, last=None):
        L.rnn_cells_rnnca_run(cells, first, k, last, oy, len_y, oc, len_c, RNNCA_LEN_POS, RNNCA_EDGES)

    run(warm, pin, None)
    sampler = ClockSampler(local_rank)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record(stream)
    run(steps)
    b.record(stream)
    barrier()
    ms = rdist.max_over_ranks(dist, a.elapsed_time(b), device="cuda")
    # end to end: a host frame in, a host frame out, every step
    e2e_steps = min(steps, 100)
    for _ in range(3):
        L.rnn_cells_rnnca_frame(cells, pin, pout, oy, len_y, oc, len_c, RNNCA_LEN_POS, RNNCA_EDGES)
    barrier()
    a.record(stream)
    for _ in range(e2e_steps):
        L.rnn_cells_rnnca_frame(cells, pin, pout, oy, len_y, oc, len_c, RNNCA_LEN_POS, RNNCA_EDGES)
        host_in, host_out = host_out, host_in
        pin, pout = pout, pin
    b.record(stream)
    barrier()
    e2e_ms = rdist.max_over_ranks(dist, a.elapsed_time(b), device="cuda")
    sampler.stop()
    L.rnn_cells_delete(cells)
    if dist:
        barrier()
        L.rnn_b200_comm_leave()

    line = None
    if rank == 0:
        c = net.contents
        # per cell and frame: hidden state read once and written once, 3 bytes in, 3 out
        # (neighbour bytes and the second pass over the state come from L1/L2)
        alg_bytes = n // world * (2 * c.h_size * 4 + 6)     # per GPU: its band of cells
        alg_flops = 2.0 * (n // world) * (c.i_size * c.h_size + c.h_size * 3)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak = float(peaks["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            peak, peak_src = 6550.0, "fallback (B200_PROFILING.md)"
        per_launch = ms / steps * 1e-3
        achieved = alg_bytes / per_launch / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")))
            traffic = next((v["dram_bytes_per_launch"] for k, v in tr.items()
                            if k.startswith("k_cells_frame_tc")), None)
        except Exception:
            pass
        line = {"metric": RNNCA_METRIC, "value": steps / (ms * 1e-3), "unit": RNNCA_UNIT,
                "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": rnnca_config(world),
                "e2e": {"value": e2e_steps / (e2e_ms * 1e-3), "unit": RNNCA_UNIT,
                        "h2d_bytes_per_step": 3 * n, "d2h_bytes_per_step": 3 * n,
                        "ms_per_step": e2e_ms / e2e_steps},
                "gpu_launches": 2 * steps, "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "kernel": "k_cells_frame_tc", "achieved": achieved,
                             "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "fp32_tflops": alg_flops / per_launch / 1e12,
                             "bytes_moved_per_launch": n // world * (2 * 256 + 2 * 8 + 6),
                             "note": "algorithmic bytes = cells x (hidden state once in and once "
                                     "out as FP32, 2 x h_size x 4, + 3 frame bytes in + 3 out).  "
                                     "The kernel keeps the state as FP16 hi/lo operand planes "
                                     "padded to 64 units (256 B per cell each way) plus 8 B of "
                                     "per-cell sums: bytes_moved_per_launch is what it must move "
                                     "in that layout, `traffic` what ncu saw.  The %.1f GFLOP of "
                                     "multiply-adds per frame run on the tensor cores "
                                     "(fp32_tflops = algorithmic FLOPs / time)" % (alg_flops / 1e9)}}
        if not args.no_cpu_baseline and world == 1:
            try:
                cores = host_cores()
                rate, wall = rnnca_reference_throughput(cores, 64, 12)
                one, wall1 = rnnca_reference_throughput(1, 64, 12)
                line["cpu_baseline"] = {
                    "value": rate, "unit": RNNCA_UNIT, "cores": cores, "kind": "reference",
                    "one_core": one,
                    "sample": "%d processes, each a band of 64 rows x %d cells for 12 frames: "
                              "reference rnn_opinion + fast_sigmoid_array around the oracle's "
                              "fill_net_inputs; %.1f s wall (one_core: one band alone, %.1f s; the "
                              "element itself runs fill_frame on one thread)"
                              % (cores, RNNCA_W, wall, wall1)}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": RNNCA_UNIT, "cores": 0,
                                        "kind": "reference", "sample": "unavailable: %s" % e}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------
# --config default | classify | multi: BASELINE.json configs[0], [2], [3] through
# the calls their reference callers make (scripts/configs_speed.py holds the
# loops; tests/test_gpu_configs.py checks the same loops for parity)

SMALL_CONFIGS = {
    "default": ("config1", "chars_per_second_default_net", "BASELINE.json configs[0]"),
    "classify": ("config3", "channel_frames_per_second", "BASELINE.json configs[2]"),
    "multi": ("config4", "text_chars_per_second", "BASELINE.json configs[3]"),
}


def run_small_config(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import configs_speed
    fn_name, metric, where = SMALL_CONFIGS[args.config]
    reference = args.impl == "reference"
    if not reference:
        from recur_b200 import api
        L = api.load_library()
        if L.rnn_b200_device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        n0 = L.rnn_b200_kernel_launches()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)     # the reference's chatter at net creation
    try:
        name, unit, ours_v, ref_v = getattr(configs_speed, fn_name)(
            ours=not reference, theirs=reference or not args.no_cpu_baseline)
    finally:
        os.dup2(saved, 2)
    v = ref_v if reference else ours_v
    line = {"metric": metric, "value": v, "unit": unit, "n_gpus": 1, "steps": None,
            "warmup": None, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name + " [" + where + "]",
                       "timing": "wall clock around the host-driven loop, inputs from and results "
                                 "to host memory every step (these are latency-bound API paths: "
                                 "the value IS the end-to-end number)"},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": None,
                    "d2h_bytes_per_step": None},
            "roofline": None,
            "cpu_baseline": {"value": ref_v, "unit": unit, "cores": 1, "kind": "reference",
                             "sample": "the same loop over the compiled reference on one host "
                                       "core (the reference is single-threaded)"}
            if ref_v is not None else None}
    if reference:
        line["impl"] = "reference"
        line["gpu_launches"] = 0
    else:
        line["gpu_launches"] = int(L.rnn_b200_kernel_launches() - n0)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="text",
                    choices=["text", "rnnca", "default", "classify", "multi"],
                    help="text: the headline (BASELINE configs[1]); rnnca: configs[4]; "
                         "default / classify / multi: configs[0] / [2] / [3]")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", type=int, default=None, help="0 auto, 1 FMA, 2 tensor")
    ap.add_argument("--hidden", type=int, default=HIDDEN)
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--cold", action="store_true", help="do not pre-fill the BPTT ring")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="exchange deltas with NCCL only")
    ap.add_argument("--trained-after", type=int, default=2000,
                    help="positions to train before the `trained` sub-record (0: skip it)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record")
    args = ap.parse_args()
    if args.config in SMALL_CONFIGS:
        return run_small_config(args)
    if args.config == "rnnca":
        return run_rnnca_reference(args) if args.impl == "reference" else run_rnnca(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

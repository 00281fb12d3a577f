#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config: RTFx (audio s / wall s) at 1024 concurrent streams.

One "step" = every stream advances by one cache-aware chunk (24 new 10-ms frames = 0.24 s of audio per stream):
batched push of 3840 samples per stream -> GPU log-mel frontend -> FastConformer chunk (57-frame slice, 256-step cache)
-> TDT greedy decode -> cache carry-over.  Streams are sharded over the ranks (stream i -> rank i mod N), no collective
on the data path; the only cross-rank traffic is the barrier + max-reduce of the timing.

  value : whole-job RTFx with the audio already resident in HBM (pkb_engine_push_audio_batch_device + pkb_engine_step),
          device-timed with CUDA events on the engine's stream, max over ranks.
  e2e   : the same through the host-facing C ABI call (pkb_engine_push_audio_batch from pinned host memory + step; the
          step ends with the D2H read of every stream's decode trace), wall-clock between device syncs, max over ranks.
  roofline    : the dominant kernel (tcgen05 GEMM): algorithmic 2*M*N*K per launch / its CUDA-event duration, vs the
                measured sustained bf16 peak of MEASURED_PEAKS.json.
  cpu_baseline: the CPU oracle (C restatement of rust/features + PyTorch restatement of the NeMo modules the reference's
                ORT-CPU runner executes) on the box's host cores, on a bounded sample of the same workload.
`--impl reference` times that CPU path alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "trt-asr-engine_b200")
for p in (PKG, os.path.join(PKG, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

SAMPLES_PER_STEP = 3840          # shift of 24 feature frames (contract.json:263-266)
AUDIO_S_PER_STEP = 0.24
N_CLIPS = 32
METRIC = "RTFx (audio s / wall s) @1024 streams"


def shard(n_streams: int, rank: int, world: int):
    """stream i -> rank i mod world; no data-path collective."""
    return [i for i in range(n_streams) if i % world == rank]


def reduce_timings(vals, world: int, device="cuda"):
    """max over ranks of each timing (the job finishes when the slowest rank does)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def load_traffic():
    """per-launch DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), float(d.get("hbm_gbs", 6650.0)), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.windows = index, None, [], []

    def start(self):
        """Start the sampling process (it takes a few hundred ms to deliver its first line, so it is started during the warm-up
        steps); only samples received inside a window opened by begin() / end() -- the timed regions -- are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for ts, ln in self.lines if any(a <= ts <= (b if b is not None else ts) for a, b in self.windows)]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU path (oracle)
def cpu_path_rtfx(model_dir: str, n_streams: int, n_chunks: int, warm_chunks: int = 1):
    """The reference's CPU path restated: C log-mel (all host threads) + PyTorch-CPU streaming encoder + greedy TDT.
    Returns (rtfx, cores, wall_s, audio_s).  Only the checker is executed here -- never the product."""
    import ctypes

    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from model_ref import DecodeState, ModelRef, prime, streaming_schedule, tdt_greedy_chunk
    from synth_audio import synth_clip
    so = os.path.join(ROOT, "oracle", "_build", "libfeatures_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libfeatures_ref.so"], stdout=subprocess.DEVNULL)
    fr = ctypes.CDLL(so)
    fr.fr_num_frames.restype = ctypes.c_size_t
    fr.fr_num_frames.argtypes = [ctypes.c_size_t]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = ModelRef(model_dir)
    total = warm_chunks + n_chunks
    sched = streaming_schedule(total)
    n_samp = (sched[-1][1] - 1) * 160 + 400
    clips = [synth_clip(n_samp / 16000.0 + 0.1, 1000 + i)[:n_samp] for i in range(n_streams)]
    states = []
    for _ in range(n_streams):
        st = DecodeState(m)
        prime(m, st)
        states.append(st)
    cc, ct, cl = m.initial_cache(n_streams)
    feats = np.zeros((n_streams, sched[-1][1], 128), np.float32)
    done_frames = 0
    t0 = None
    for k, (b, e) in enumerate(sched):
        if k == warm_chunks:
            t0 = time.perf_counter()
        # frontend for the frames this chunk needs that are not computed yet (streaming, like the GPU path)
        if e > done_frames:
            for i in range(n_streams):
                seg = np.ascontiguousarray(clips[i][done_frames * 160:(e - 1) * 160 + 400])
                fr.fr_logmel(seg.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(seg.size),
                             feats[i, done_frames:e].ctypes.data_as(ctypes.c_void_p), ctypes.c_int(cores))
            done_frames = e
        x = torch.from_numpy(np.ascontiguousarray(feats[:, b:e].transpose(0, 2, 1)))
        enc, el, cc, ct, cl = m.stream_step(x, torch.full((n_streams,), e - b, dtype=torch.int64), cc, ct, cl)
        for i in range(n_streams):
            tdt_greedy_chunk(m, states[i], enc[i:i + 1], int(el[i]))
    wall = time.perf_counter() - t0
    audio_s = n_streams * n_chunks * AUDIO_S_PER_STEP
    return audio_s / wall, cores, wall, audio_s


def cpu_config1_offline(model_dir: str, seconds: float = 10.0):
    """BASELINE config 1 on the CPU checker: one synthetic clip, batch 1, log-mel + per-feature normalisation + full-context encoder +
    greedy TDT (the reference's ORT-CPU offline path restated).  Returns (rtfx, wall_s)."""
    import ctypes

    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from model_ref import DecodeState, ModelRef, prime, tdt_greedy_chunk
    from synth_audio import synth_clip
    so = os.path.join(ROOT, "oracle", "_build", "libfeatures_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libfeatures_ref.so"], stdout=subprocess.DEVNULL)
    fr = ctypes.CDLL(so)
    fr.fr_num_frames.restype = ctypes.c_size_t
    fr.fr_num_frames.argtypes = [ctypes.c_size_t]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = ModelRef(model_dir)
    pcm = synth_clip(seconds, 1234)
    best = None
    for _ in range(2):                       # second pass warm
        t0 = time.perf_counter()
        T = fr.fr_num_frames(pcm.size)
        feat = np.zeros((T, 128), np.float32)
        fr.fr_logmel(pcm.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(pcm.size), feat.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(cores))
        feat = (feat - feat.mean(0)) / (feat.std(0, ddof=1) + 1e-5)
        enc, el = m.offline(torch.from_numpy(np.ascontiguousarray(feat.T)[None]), torch.tensor([T]))
        st = DecodeState(m)
        prime(m, st)
        tdt_greedy_chunk(m, st, enc, int(el))
        wall = time.perf_counter() - t0
        best = wall if best is None else min(best, wall)
    return seconds / best, best


def run_reference(args, rank):
    if rank != 0:
        return
    from make_synthetic_model import ensure_model
    model = ensure_model(os.path.join(ROOT, "models", f"synth{args.layers}"), n_layers=args.layers, seed=0)
    n_streams = args.ref_streams
    rtfx, cores, wall, audio_s = cpu_path_rtfx(model, n_streams, args.steps, max(args.warmup, 1))
    sample = f"{n_streams} streams x {args.steps} chunks ({audio_s:.2f} s audio) after {max(args.warmup,1)} warm-up chunk(s)"
    line = {"impl": "reference", "metric": METRIC, "value": rtfx, "unit": "x real time", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "streaming chunk step (57-frame slice, cache 256) + TDT greedy, CPU restatement of rust/features + "
                                   "NeMo modules behind the ORT-CPU runner", "model": f"parakeet-tdt-0.6b-v3 synthetic weights, {args.layers} layers",
                       "streams": n_streams},
            "cpu_baseline": {"value": rtfx, "unit": "x real time", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rtfx, "unit": "x real time", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    c1_rtfx, c1_wall = cpu_config1_offline(model)
    line["config1_offline_10s"] = {"rtfx": c1_rtfx, "wall_s": c1_wall, "cores": cores,
                                   "what": "BASELINE config 1: one 10 s clip, batch 1, log-mel + per-feature norm + full-context encoder + greedy TDT on the CPU port"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU path
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--streams", type=int, default=1024, help="total concurrent streams over all ranks")
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--precision", type=int, default=int(os.environ.get("PKB_BENCH_PRECISION", "0")),
                    help="0 = bf16 operands, 1 = split bf16 hi+lo (fp32-grade)")
    ap.add_argument("--ref-streams", type=int, default=64, help="streams batched together in the bounded CPU sample (cpu_baseline / --impl reference)")
    ap.add_argument("--ref-chunks", type=int, default=5, help="chunks per stream of the cpu_baseline sample (~10-20 s of CPU work at 64 streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--engines-per-gpu", type=int, default=0, help="0 = auto (2 when this rank holds <= --two-engine-max-streams streams)")
    ap.add_argument("--two-engine-max-streams", type=int, default=0)
    ap.add_argument("--longform", type=int, default=4, help="BASELINE config 5's shape: this many clips through the whole-utterance offline path "
                                                            "(4 x 1 h by default so that the default run stays within minutes; 32 = the full config; 0 = skip)")
    ap.add_argument("--longform-seconds", type=float, default=3600.0)
    ap.add_argument("--longform-blank-penalty", type=float, default=None,
                    help="PARAKEET_BLANK_PENALTY for the long-form run (default: calibrated so that the random-weight model emits speech-like token rates)")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--prefill-chunks", type=int, default=88,
                    help="untimed chunks per stream before the timed region; 86 saturate the 256-step attention cache (1 + 3 per chunk)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import binding
    from make_synthetic_model import ensure_model
    from synth_audio import synth_clip

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    model = os.path.join(ROOT, "models", f"synth{args.layers}")
    if rank == 0:
        ensure_model(model, n_layers=args.layers, seed=0)
    if world > 1:
        dist.barrier()
    if not os.path.exists(binding.LIB_PATH):
        import __graft_entry__ as g
        g.build()

    mine = shard(args.streams, rank, world)
    n = len(mine)
    # Small per-GPU batches are bound by the latency of ~410 dependent launches per step, not by throughput: two engines per GPU
    # (each with half of this rank's streams, its own CUDA stream and buffers, driven by its own host thread) interleave two such
    # chains on the same GPU.  Large batches are throughput-bound and use one engine.
    n_eng = args.engines_per_gpu if args.engines_per_gpu > 0 else (2 if 2 <= n <= args.two_engine_max_streams else 1)
    bounds = [(j * n // n_eng, (j + 1) * n // n_eng) for j in range(n_eng)]
    engs = [binding.Engine(model, device_id=local_rank, max_streams=hi - lo, precision=args.precision, max_rows=8 * (hi - lo))
            for lo, hi in bounds]
    sid_lists = [np.array([e.open() for _ in range(hi - lo)], np.int32) for e, (lo, hi) in zip(engs, bounds)]
    eng = engs[0]
    sids = sid_lists[0]
    pool = None
    if n_eng > 1:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(n_eng)

    def each(fn):
        """fn(engine index) on every engine: concurrently (ctypes calls release the GIL) when there is more than one"""
        return [fn(0)] if pool is None else list(pool.map(fn, range(n_eng)))

    prof_steps = 2
    total_pushes = 16            # audio ring: the pushes cycle through 16 x 0.24 s of synthetic audio per stream
    clip_len = total_pushes * SAMPLES_PER_STEP + 16000
    clips = np.stack([synth_clip(clip_len / 16000.0 + 0.01, 1000 + c)[:clip_len] for c in range(N_CLIPS)])
    # the j-th stream of EVERY rank carries clip j mod 32 at phase 977 j: each rank sees the same clip mix (and decode load) at every N
    local = np.array(mine) // world
    phase = (local * 977) % 12000
    clip_of = local % N_CLIPS

    def host_step_audio(k: int) -> np.ndarray:
        idx = phase[:, None] + k * SAMPLES_PER_STEP + np.arange(SAMPLES_PER_STEP)[None, :]
        return clips[clip_of[:, None], idx]

    host = torch.empty((total_pushes, n, SAMPLES_PER_STEP), dtype=torch.float32).pin_memory()
    for k in range(total_pushes):
        host[k] = torch.from_numpy(host_step_audio(k))
    dev = host.cuda(non_blocking=False)
    torch.cuda.synchronize()
    step_bytes = n * SAMPLES_PER_STEP * 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cursor = [0]

    def do_step(resident: bool):
        k = cursor[0] % total_pushes
        cursor[0] += 1

        def work(j):
            off = bounds[j][0] * SAMPLES_PER_STEP * 4
            if resident:
                engs[j].push_audio_batch_device(sid_lists[j], dev[k].data_ptr() + off, SAMPLES_PER_STEP, SAMPLES_PER_STEP)
            else:
                engs[j].push_audio_batch(sid_lists[j], host[k].data_ptr() + off, SAMPLES_PER_STEP, SAMPLES_PER_STEP)
            return engs[j].step()
        return sum(each(work))

    # prefill: 2 pushes = 46 frames >= the 41 frames of chunk 0; afterwards every push yields exactly one chunk per stream
    for j in range(n_eng):
        engs[j].push_audio_batch_device(sid_lists[j], dev[0].data_ptr() + bounds[j][0] * SAMPLES_PER_STEP * 4, SAMPLES_PER_STEP, SAMPLES_PER_STEP)
    cursor[0] = 1
    assert do_step(True) == n, "prefill did not produce chunk 0 for every stream"
    sampler = ClockSampler(local_rank)
    for i in range(max(args.prefill_chunks - 1, 0) + args.warmup):
        if i == max(args.prefill_chunks - 1, 0):
            sampler.start()          # running (and past its start-up latency) before the timed regions begin
        assert do_step(True) == n
    cache_len0 = eng.cache_len(int(sids[0]))

    # ---- timed region 1: audio resident in HBM, CUDA events on the engine stream
    barrier()
    sampler.begin()
    launches0 = sum(e.kernel_launches() for e in engs)
    ev0 = [e.event_record() for e in engs]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        do_step(True)
    ev1 = [e.event_record() for e in engs]
    # device time of the region: every engine's own stream, first event to last event; the job ends with the slowest engine
    dev_ms = max(e.event_elapsed_ms(a, b) for e, a, b in zip(engs, ev0, ev1))
    barrier()
    wall_resident = time.perf_counter() - t0
    launches = sum(e.kernel_launches() for e in engs) - launches0
    sampler.end()

    # ---- timed region 2: end to end from pinned host memory through the C ABI (H2D + step + D2H of the decode traces)
    for _ in range(2):
        do_step(False)
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        do_step(False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.end()
    barrier()
    clocks = sampler.stop()      # samples taken during the two timed regions (device-resident and end-to-end)
    graphs = sum(e.graphs_built() for e in engs)
    # decode loops of the steps so far (graph mode): device time between the CUDA events the step graph records around its WHILE node
    loop = [e.decode_loop_stats(reset=True) for e in engs]
    loop_ms, loop_bytes, loop_passes, loop_n = (sum(x[i] for x in loop) for i in range(4))

    # ---- per-launch timing of the dominant kernel (tcgen05 GEMM) for the roofline
    for e in engs:
        e.profile_enable(True)
    for _ in range(prof_steps):
        do_step(True)

    def prof(cls):
        r = [e.profile_read_class(cls) for e in engs]
        return sum(x[0] for x in r), sum(x[1] for x in r), sum(x[2] for x in r)
    gemm_ms, gemm_flops, gemm_launches = prof(0)
    attn_ms, attn_bytes, attn_launches = prof(1)
    fe_ms, fe_bytes, fe_launches = prof(2)
    dec_ms, dec_bytes, dec_loops = prof(3)
    for e in engs:
        e.profile_enable(False)
    # the frontend kernel at a size where it is not launch-bound: log-mel of one 1 h clip (BASELINE config 5's frontend)
    fe1h = None
    if world == 1 and not args.no_latency:
        eng.profile_enable(True)
        hour = np.tile(synth_clip(10.0, 1234), 360)          # 57.6 M samples
        eng.logmel(hour)
        fe1h = eng.profile_read_class(2)
        eng.profile_enable(False)

    n_tokens = sum(len(eng.tokens(int(s))) for s in sids[: min(len(sids), 64)])
    dev_s, e2e_max, wall_max = reduce_timings([dev_ms / 1e3, e2e_s, wall_resident], world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    audio_s = args.streams * args.steps * AUDIO_S_PER_STEP
    peak_tf, peak_hbm, peak_src = load_peaks()
    traffic = load_traffic()
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": audio_s / dev_s, "unit": "x real time", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision == 0 else "bf16x2 (split hi+lo operands, fp32-grade)", "data": "synthetic",
        "config": {"workload": f"{args.streams} concurrent streams sharded over {world} GPU(s) (stream i -> rank i mod N), cache-aware "
                               "streaming chunk step (57-frame slice, cache 256, 24-frame shift) + TDT greedy decode, GPU log-mel frontend",
                   "model": f"parakeet-tdt-0.6b-v3 architecture, seeded random weights, {args.layers} layers",
                   "streams_per_gpu": n, "engines_per_gpu": n_eng, "audio_s_per_step": args.streams * AUDIO_S_PER_STEP, "precision": args.precision,
                   "l2_policy": "working set per step (1.2 GB weights + 29 MB K/V per stream) exceeds the 126 MB L2; no flush needed",
                   "wall_ms_per_step_resident": 1e3 * wall_max / args.steps, "tokens_emitted_first_64_streams": n_tokens,
                   "cache_last_channel_len_at_timing": cache_len0, "prefill_chunks": args.prefill_chunks,
                   "step_graphs": graphs,
                   "execution": ("one CUDA graph per step shape (encoder chunk + device-side WHILE node around the decode iteration)" if graphs
                                 else "launch by launch"),
                   "cpu_arm_streams": args.ref_streams},
        "clocks": clocks,
        "e2e": {"value": audio_s / e2e_max, "unit": "x real time", "h2d_bytes_per_step": step_bytes * world,
                "d2h_bytes_per_step": n * 97 * 4 * world},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                     "algorithmic_flops_per_launch": gemm_flops / max(gemm_launches, 1),
                     "kernel": "gemm_tc_kernel / gemm_tc2_kernel (tcgen05 + TMA + TMEM; single-CTA and CTA-pair variants)",
                     "launches_timed": int(gemm_launches),
                     "peak_source": f"{peak_src} sustained bf16 (MEASURED_PEAKS.json)"},
    }
    # the HBM-bound pieces, same method (CUDA events around each launch, algorithmic bytes / duration) vs the measured copy peak
    line["roofline_hbm"] = [
        {"kernel": "attention_mma_kernel (K/V ring read: valid 8-key groups only, TMA + mma.sync)", "bound": "hbm", "achieved": attn_bytes / max(attn_ms, 1e-9) / 1e6,
         "peak": peak_hbm, "unit": "GB/s", "frac": attn_bytes / max(attn_ms, 1e-9) / 1e6 / peak_hbm, "launches_timed": int(attn_launches),
         "algorithmic_bytes_per_launch": attn_bytes / max(attn_launches, 1)},
        {"kernel": "logmel_reg_kernel (frontend, register-resident FFT)", "bound": "hbm", "achieved": fe_bytes / max(fe_ms, 1e-9) / 1e6, "peak": peak_hbm, "unit": "GB/s",
         "frac": fe_bytes / max(fe_ms, 1e-9) / 1e6 / peak_hbm, "launches_timed": int(fe_launches),
         "algorithmic_bytes_per_launch": fe_bytes / max(fe_launches, 1),
         "note": "latency-bound at this size (24 new frames per stream per step); 0.3 % of the step"},
    ]
    if loop_n > 0:
        dec_ms, dec_bytes, dec_loops, dec_how = loop_ms, loop_bytes, loop_n, (
            f"inside the step graph (device-side WHILE node): CUDA events recorded by the graph around the loop, {loop_passes / loop_n:.1f} passes "
            f"per loop, {1e3 * loop_ms / max(loop_passes, 1):.1f} us per pass")
    else:
        dec_how = "launch-by-launch path, host-polled loop"
    line["roofline_hbm"].append(
        {"kernel": "TDT decode loop (joint_hidden -> joint output GEMM with fused argmax -> tdt_select -> predictor pass), whole loop",
         "bound": "hbm", "achieved": dec_bytes / max(dec_ms, 1e-9) / 1e6, "peak": peak_hbm, "unit": "GB/s",
         "frac": dec_bytes / max(dec_ms, 1e-9) / 1e6 / peak_hbm, "launches_timed": int(dec_loops),
         "algorithmic_bytes_per_launch": dec_bytes / max(dec_loops, 1), "how": dec_how,
         "note": "bytes = passes x (10.5 MB joint output weights + per-stream rows); the 24 MB of decoder weights stay in the 126 MB L2 and "
                 "at 1024 streams a pass is two split-precision tensor-core GEMMs deep, so this entry is latency / tensor bound, not HBM bound"})
    if fe1h is not None and fe1h[0] > 0:
        line["roofline_hbm"].append(
            {"kernel": f"logmel_reg_kernel, one 1 h clip ({(hour.size - 400) // 160 + 1} frames in one launch)", "bound": "hbm", "achieved": fe1h[1] / fe1h[0] / 1e6,
             "peak": peak_hbm, "unit": "GB/s", "frac": fe1h[1] / fe1h[0] / 1e6 / peak_hbm, "launches_timed": int(fe1h[2]),
             "algorithmic_bytes_per_launch": fe1h[1] / max(fe1h[2], 1)})
    for e in engs:
        e.close()
    if not args.no_config3 and world == 1:
        line["config3_64streams_mixed_cache"] = config3_mixed(binding, model, args.precision, clips)
    if args.longform > 0 and world == 1 and not args.no_latency:
        line["config5_longform"] = longform(binding, model, args.precision, args.longform, args.longform_seconds, args.longform_blank_penalty)
    if not args.no_latency and world == 1:
        line["config1_offline_10s"] = config1_offline(binding, model, args.precision)
    if not args.no_latency and world == 1:
        line["latency_1stream"] = latency_one_stream(binding, model, args.precision, clips[0])
    if not args.no_cpu_baseline and world == 1:
        rtfx, cores, wall, a_s = cpu_path_rtfx(model, args.ref_streams, args.ref_chunks, 1)
        line["cpu_baseline"] = {"value": rtfx, "unit": "x real time", "cores": cores, "kind": "port",
                                "sample": f"{args.ref_streams} streams batched x {args.ref_chunks} chunks ({a_s:.2f} s audio), {wall:.1f} s of CPU work, "
                                          "C restatement of rust/features + PyTorch-CPU restatement of the NeMo modules (not the Rust/ORT binaries)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def longform(binding, model, precision, batch, seconds, blank_penalty=None):
    """BASELINE config 5 (long-form offline): `batch` clips of `seconds` s through pkb_offline_utterances -- full-utterance log-mel +
    per-feature normalisation + encoder over ALL frames (full self-attention) + TDT decode; host audio in, tokens out.
    Every clip is a different sequence of 10 s synthetic segments.  With random weights, attention over 45 000 frames averages the
    encoder output towards its mean and the joint then prefers blank almost everywhere (r1 measured 0 tokens per 32 h), which would
    leave the predictor / LSTM path out of the measurement; the reference's own PARAKEET_BLANK_PENALTY knob (parakeet_trt.cpp:3175-3178)
    is set for this run so that the decode emits tokens at a speech-like rate (reported as tokens_per_hour)."""
    from synth_audio import synth_clip
    n_samp = int(seconds * 16000)
    seg = 160000
    segs = [synth_clip(10.0, 4000 + k) for k in range(12)]
    rng = np.random.default_rng(7)
    audio = []
    for i in range(batch):
        order = rng.integers(0, len(segs), size=n_samp // seg + 1)
        audio.append(np.concatenate([segs[k] for k in order])[:n_samp].astype(np.float32))
    # host PCM lives in pinned memory, like the streaming arm's (e2e contract: inputs come from pinned host buffers); pageable sources made
    # the 7.4 GB of 32 x 1 h go through the driver's staging copies
    pinned = []
    try:
        import torch
        for i in range(batch):
            t = torch.from_numpy(audio[i]).pin_memory()
            pinned.append(t)
            audio[i] = t.numpy()
    except Exception:
        pinned = []
    t_enc = binding.load_library().pkb_encoded_length((n_samp - 400) // 160 + 1)
    # ~120 KB of work buffers per encoder frame: clips go through in groups that fit (4 one-hour clips = 22 GB)
    group = max(1, min(batch, int(4 * 45000 // max(t_enc, 1)) or 1))
    eng = binding.Engine(model, max_streams=batch, precision=precision, max_rows=group * t_enc + 64, contract_cache=0)
    sids = [eng.open() for _ in range(batch)]
    eng.offline_utterances(sids[:1], audio=[audio[0][:160000]], decode=True)       # warm-up (lazy buffers, first launches)
    eng.reset(sids[0])
    # blank penalty: bisect (on the first 10 minutes of clip 0, untimed) for 0.2 - 0.4 tokens per encoder frame = 2.5 - 5 tokens per second
    # (with these random weights the blank margin is nearly the same on every frame, so the token rate jumps from ~0 to ~1 per frame
    # within a fraction of a logit; when no penalty lands inside the window the smallest one that makes the decoder emit is used:
    # a token on every 80 ms frame is ~3x the rate of speech, i.e. the heavier decode load)
    cal = []
    if blank_penalty is None:
        lo_p, hi_p, pen, best = 0.0, 32.0, 16.0, None
        probe = audio[0][: min(n_samp, 600 * 16000)]
        for _ in range(8):
            pen = 0.5 * (lo_p + hi_p)
            eng.set_blank_penalty(pen)
            eng.offline_utterances(sids[:1], audio=[probe], per_feature_norm=True, decode=True)
            ratio = len(eng.tokens(sids[0])) / max(len(eng.last_steps(sids[0])), 1)
            cal.append((round(pen, 3), round(ratio, 3)))
            eng.reset(sids[0])
            if ratio >= 0.2 and (best is None or pen < best):
                best = pen
            if 0.2 <= ratio <= 0.4:
                break
            if ratio < 0.2:
                lo_p = pen
            else:
                hi_p = pen
        pen = best if best is not None else hi_p
        if not any(0.2 <= r <= 0.4 for _, r in cal):
            # no penalty lands inside the window (the transition is a step, and its position moves by a few tenths of a logit with the clip
            # length): stay clear of the step, on the emitting side -- the decoder then emits on (nearly) every frame
            pen += 1.5
    else:
        pen = blank_penalty
    eng.set_blank_penalty(pen)
    eng.profile_enable(True)
    t0 = time.perf_counter()
    for lo in range(0, batch, group):      # encode group by group (decode = 2: rows parked), then ONE batched decode of all clips
        eng.offline_utterances(sids[lo:lo + group], audio=audio[lo:lo + group], per_feature_norm=True, decode=2)
    assert eng.offline_decode_pending() == batch
    wall = time.perf_counter() - t0
    n_tok = sum(len(eng.tokens(s)) for s in sids)
    gemm_ms, gemm_flops, _ = eng.profile_read()
    att_ms, att_flops, att_l = eng.profile_read_class(4)
    dec_ms, _, _ = eng.profile_read_class(3)
    eng.profile_enable(False)
    eng.close()
    return {"workload": f"{batch} clips x {seconds:.0f} s (each a different sequence of 10 s synthetic segments) encoded in groups of {group}, decoded in one "
                        f"batched pass; whole-utterance offline (config 5 shape), encoder frames per clip {t_enc}",
            "rtfx_e2e": batch * seconds / wall, "wall_s": wall, "tokens": n_tok, "tokens_per_hour": n_tok / max(batch * seconds / 3600.0, 1e-9),
            "blank_penalty": pen, "blank_penalty_calibration": cal, "gemm_ms": gemm_ms, "gemm_tflops": gemm_flops / max(gemm_ms, 1e-9) / 1e9,
            "attention_ms": att_ms, "attention_tflops_algorithmic": att_flops / max(att_ms, 1e-9) / 1e9, "attention_launches": int(att_l),
            "decode_ms": dec_ms, "host_audio": "pinned" if pinned else "pageable"}


def config3_mixed(binding, model, precision, clips, n=64, steps=20):
    """BASELINE config 3: 64 concurrent streams, batched chunk step with PER-STREAM cache lengths, one B200.  Streams join at staggered
    times, so at the timed steps cache_last_channel_len ranges from a few rows to the saturated 256 inside one batch."""
    import torch
    eng = binding.Engine(model, max_streams=n, precision=precision, max_rows=8 * n)
    sids = np.array([eng.open() for _ in range(n)], np.int32)
    start = np.array([(j % 8) * 12 for j in range(n)])          # step at which stream j joins: 0, 12, ..., 84
    total = 16
    host = torch.empty((total, n, SAMPLES_PER_STEP), dtype=torch.float32).pin_memory()
    for k in range(total):
        idx = ((np.arange(n) * 977) % 12000)[:, None] + k * SAMPLES_PER_STEP + np.arange(SAMPLES_PER_STEP)[None, :]
        host[k] = torch.from_numpy(clips[(np.arange(n) % N_CLIPS)[:, None], idx])
    pushed = np.zeros(n, np.int64)

    def step(t):
        act = np.nonzero(start <= t)[0]
        # every active stream gets its next 0.24 s; a stream's own push counter selects the audio slice (all slices are equally valid audio)
        for grp in np.unique(pushed[act] % total):
            sel = act[pushed[act] % total == grp]
            rows = np.ascontiguousarray(host[int(grp)].numpy()[sel])
            eng.push_audio_batch(sids[sel], rows.ctypes.data, SAMPLES_PER_STEP, SAMPLES_PER_STEP)
        pushed[act] += 1
        return eng.step()
    t = 0
    while t < 88 + 2:                      # the first streams reach a saturated cache; the last joined 6 steps ago
        step(t)
        t += 1
    lens = [eng.cache_len(int(s)) for s in sids]
    for _ in range(3):
        assert step(t) == n
        t += 1
    ev0 = eng.event_record()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(t)
        t += 1
    ev1 = eng.event_record()
    dev_ms = eng.event_elapsed_ms(ev0, ev1)
    wall = time.perf_counter() - t0
    eng.close()
    return {"streams": n, "steps": steps, "ms_per_step_e2e": 1e3 * wall / steps, "ms_per_step_device": dev_ms / steps,
            "rtfx_e2e": n * AUDIO_S_PER_STEP * steps / wall, "cache_last_channel_len_min": int(min(lens)), "cache_last_channel_len_max": int(max(lens)),
            "distinct_cache_lengths": len(set(lens)),
            "what": "BASELINE config 3: 64 streams in one batched step, per-stream cache lengths (host audio in, tokens out)"}


def config1_offline(binding, model, precision):
    """BASELINE config 1 on the GPU: one synthetic 10 s clip, batch 1, host PCM in -> tokens out through pkb_offline_utterances."""
    from synth_audio import synth_clip
    pcm = synth_clip(10.0, 1234)
    eng = binding.Engine(model, max_streams=1, precision=precision, max_rows=256, contract_cache=0)
    s = eng.open()
    walls = []
    for _ in range(6):
        t0 = time.perf_counter()
        eng.offline_utterances([s], audio=[pcm], per_feature_norm=True, decode=True)
        walls.append(time.perf_counter() - t0)
        eng.reset(s)
    eng.close()
    w = float(np.median(walls[1:]))
    return {"wall_ms": 1e3 * w, "rtfx": 10.0 / w, "what": "BASELINE config 1: one 10 s clip, batch 1, host PCM in -> tokens out (whole-utterance path)"}


def latency_one_stream(binding, model, precision, clip, prefill=90, timed=60):
    """BASELINE config 2: one stream, per-chunk latency (push of 0.24 s of host audio -> tokens on the host), p50 / p95, measured in
    steady state: `prefill` untimed chunks first, so the 256-step attention cache is saturated (86 chunks fill it)."""
    eng = binding.Engine(model, max_streams=1, precision=precision)
    sid = np.array([eng.open()], np.int32)
    ring = 16 * SAMPLES_PER_STEP
    buf = np.ascontiguousarray(clip[:ring])
    eng.push_audio_batch(sid, buf[:SAMPLES_PER_STEP].ctypes.data, SAMPLES_PER_STEP, SAMPLES_PER_STEP)
    lat = []
    for k in range(1, prefill + timed + 1):
        off = (k % 16) * SAMPLES_PER_STEP
        seg = buf[off:off + SAMPLES_PER_STEP]
        t0 = time.perf_counter()
        eng.push_audio_batch(sid, seg.ctypes.data, SAMPLES_PER_STEP, SAMPLES_PER_STEP)
        n = eng.step()
        dt = 1e3 * (time.perf_counter() - t0)
        if k > prefill:
            assert n == 1
            lat.append(dt)
    cache_len = eng.cache_len(int(sid[0]))
    graphs = eng.graphs_built()
    eng.close()
    lat = np.array(lat)
    return {"p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)), "chunks": int(lat.size),
            "cache_last_channel_len": cache_len, "step_graphs": graphs}


if __name__ == "__main__":
    main()

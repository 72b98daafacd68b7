#!/usr/bin/env python
"""bench.py -- policy-net leaf evaluations per second (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                       # CPU arm: the oracle port on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...            # one rank per GPU (driver does this)

A "step" = one pass of the hot path (encode -> conv stack -> FC-155 -> softmax -> legal gather +
sort -> ranked child priors) over one batch of synthetic leaf positions per GPU.
Workload = BASELINE.json configs[3] (the batch sweep the metric is quoted on) at its top batch,
65,536 positions per GPU per step, leaf batches sharded one slice per GPU (weak scaling, no
collective on the data path).  The smaller configs (1 / 256 / 1024 positions) are parity-test
cases; their throughput (and the rest of the sweep, per-GPU batch sizes, whole-job
positions/s) is listed under "sweep_positions_per_s" for information.

Arithmetic of every number in the line unless a key says otherwise (`arithmetic` in `config`): the
configuration that meets north_star's accuracy gates on 100,000 positions
(tests/test_gpu_parity.py::test_centred_fp16_mode_meets_the_north_star_gates_on_100k_positions):
mode fp16c as created: centred fp16 operands, fp32 accumulate, the weights of the 192->192 layers rounded to
fp16 with error compensation against the calibration inputs' second moments (GPTQ-style, wann_b200/csrc/gptq.hpp),
ONE MMA per K-step (the MMA count of the bf16 kernel), no refinement pass (one kernel launch per evaluation).  Faster, less accurate
modes -- and the same mode with margin refinement, which makes the policy move equal the fp32
graph's on all 100,000 positions -- are extra keys under `modes`.

`value`  : whole-job positions/s, inputs resident in HBM, device pointers through the C ABI.
`e2e`    : same metric through the public host-buffer API (wann_eval_topk with pinned host
           arrays): every step copies its 65 B/position inputs H2D and its 241 B/position ranked
           priors D2H inside the timed region.
`roofline`: conv_stack_tc kernel, ALGORITHMIC FLOPs per launch (the extra MMAs of a split layer are
           overhead, not counted) / CUDA-event time of the launch, against the measured cuBLAS bf16
           BURST peak in MEASURED_PEAKS.json (the conservative denominator: over >= 1 s this kernel runs
           above cuBLAS's own sustained figure, which is shown beside it as `frac_of_sustained_peak`);
           `sustained` repeats the measurement over >= 3 s.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_POS_CONV = 2 * (442368 + 3 * 21233664 + 110592)      # conv1..conv5, SURVEY 8(d)
FLOP_PER_POS_ALL = FLOP_PER_POS_CONV + 2 * 9920                # + FC-155 = 128,527,744
BATCH = 65536
POOL_SETS = 32                                                 # 32 x (4.3 MB in + 15.8 MB out) > L2
IN_BYTES, OUT_BYTES = 65, 48 + 48 * 4 + 1
METRIC, UNIT = "policy_net_leaf_evals_per_sec", "positions/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
def make_positions(n, seed):
    """Synthetic leaf positions: seeded random legal play-out positions, tiled."""
    from wann_b200 import synthetic
    base_sq, base_bl = synthetic.random_positions(4096, seed=seed)
    reps = (n + 4095) // 4096
    rng = np.random.RandomState(seed)
    perm = rng.permutation(reps * 4096)[:n]
    return np.tile(base_sq, (reps, 1))[perm], np.tile(base_bl, reps)[perm]


_CPU_ARM = {}


def _cpu_arm_setup():
    """Pick, once, the fastest of the three CPU implementations of the path on this host."""
    if _CPU_ARM:
        return _CPU_ARM
    from oracle import c_oracle as C
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    C.set_threads(cores)           # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    w = O.synthetic_weights("trained_like", 1)
    sq, bl = O.random_positions_fast(2048, seed=99)
    sq, bl = np.tile(sq, (32, 1)), np.tile(bl, 32)              # up to 65,536 positions to draw from

    def run_numpy(s, b):
        pl = O.encode_planes(s, b)
        _, pr = O.forward_fp32(pl, w)
        return O.ranked_children(pr, O.legal_mask(s, b))

    def run_c(s, b):
        pl = C.encode(s, b)
        _, pr = C.forward(pl, w)
        return C.rank(pr, C.legal_mask(s, b))

    def run_torch(s, b):
        from oracle import torch_oracle as T
        _, pr = T.forward(C.encode(s, b), w, threads=cores)
        return C.rank(pr, C.legal_mask(s, b))

    best = None
    for name, fn in (("numpy+OpenBLAS", run_numpy), ("C+OpenMP", run_c), ("torch CPU/oneDNN", run_torch)):
        try:
            fn(sq[:64], bl[:64])
            t0 = time.perf_counter()
            fn(sq[:256], bl[:256])
            rate = 256 / (time.perf_counter() - t0)
        except Exception:  # noqa: BLE001
            continue
        if best is None or rate > best[0]:
            best = (rate, name, fn)
    # For context: the reference's OWN per-board Python encode + legal-move enumeration
    # (compiled from its .pyx into oracle/_ref), which sits in front of its TF session.
    extra = ""
    try:
        from oracle.ref_loader import load_ref
        ref = load_ref()
        if ref is not None:
            boards = [(O.board_from_squares(sq[i], 0 if bl[i] else 1), "Black" if bl[i] else "White")
                      for i in range(200)]
            t0 = time.perf_counter()
            for b, c in boards:
                ref.utils.convert_board_to_2d_matrix_POEB(b, c)
                ref.board_utils.enumerate_legal_moves(b, c)
            extra = "; reference's own Python encode+legal-move code: %.0f us/board (1 thread)" % (
                (time.perf_counter() - t0) / 200 * 1e6)
    except Exception:  # noqa: BLE001
        pass
    _CPU_ARM.update(cores=cores, sq=sq, bl=bl, rate=best[0], name=best[1], fn=best[2], extra=extra)
    return _CPU_ARM


def cpu_arm(sample_seconds=15.0):
    """The oracle port timed on the host cores: encode -> fp32 forward -> legal mask -> rank, on a
    bounded sample of the bench workload.  Returns (positions/s, cores, kind, sample description)."""
    a = _cpu_arm_setup()
    n = int(min(len(a["sq"]), max(256, a["rate"] * sample_seconds)))
    t0 = time.perf_counter()
    a["fn"](a["sq"][:n], a["bl"][:n])
    dt = time.perf_counter() - t0
    sample = "%d positions, oracle port (%s, fp32), %.1f s" % (n, a["name"], dt) + a["extra"]
    return n / dt, a["cores"], "port", sample


def run_reference(args, rank):
    if rank != 0:
        return
    vals = []
    sample = ""
    # the whole run stays within ~2.5 minutes whatever K and W are
    per_step = max(0.25, min(15.0, 150.0 / (args.warmup + args.steps)))
    if os.environ.get("WANN_B200_BENCH_CPU_SECONDS"):          # tests: a shorter sample per step
        per_step = float(os.environ["WANN_B200_BENCH_CPU_SECONDS"])
    for i in range(args.warmup + args.steps):
        v, cores, kind, sample = cpu_arm(sample_seconds=per_step)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[3] batch sweep, top batch: 65536 positions per GPU per step "
                                   "(CPU arm: bounded sample of the same workload per step)",
                       "weights": "synthetic trained_like seed 1 (reference shard absent)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "note": "CPU stand-in for the TF-1.0 session (TensorFlow is not installable here)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings
    written by C libraries to fd 1) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fp16c", choices=["bf16", "fp16", "fp16c", "fp32_tc"])
    ap.add_argument("--split-layers", type=int, default=None, help="fp16c: mask of hidden layers (bits 1..3) with lo weight stages; default 0")
    ap.add_argument("--gptq", type=int, default=1, help="fp16c: data-aware weight rounding (default 1)")
    ap.add_argument("--lo-n", type=int, default=64, help="fp16c: output positions covered by the lo stages")
    ap.add_argument("--lo-kgroups", type=int, default=1, help="fp16c: 64-channel input groups covered by the lo stages")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--refine-margin", type=float, default=None, help="logits; default 0 (no refinement pass)")
    ap.add_argument("--sustained-seconds", type=float, default=3.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 0)
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import torch
    import torch.distributed as dist
    import wann_b200
    from wann_b200 import weights as W
    from wann_b200.sharding import gather_counters, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world if world > 1 else 1
    if args.gpus != n_gpus and rank == 0:
        print("note: --gpus %d but WORLD_SIZE=%d; running on %d GPU(s)" % (args.gpus, world, n_gpus),
              file=sys.stderr)

    B = args.batch
    split_layers = 0 if args.split_layers is None else args.split_layers
    refine_margin = 0.0 if args.refine_margin is None else args.refine_margin
    ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), device=local_rank, mode=args.mode,
                                   refine_margin=refine_margin)
    if args.mode == "fp16c" and not args.gptq:
        ev.set_option("gptq", 0)
    if split_layers:
        ev.set_option("lo_n", args.lo_n)
        ev.set_option("lo_kgroups", args.lo_kgroups)
        ev.set_option("split_layers", split_layers)
    arithmetic = {"bf16": "bf16 operands, fp32 accumulate", "fp16": "fp16 operands, fp32 accumulate",
                  "fp32_tc": "fp16 hi/lo split operands (3 MMAs per K-step), fp32 accumulate",
                  "fp16c": "centred fp16 operands (a - bucket mean; conv1/conv5 weights hi+lo; 192->192 weights rounded with "
                           "error compensation, gptq=%d), fp32 accumulate" % args.gptq}[args.mode]
    arithmetic += "; split_layers=%d (lo_n=%d, lo_kgroups=%d); margin refinement %.3g logit (fp32-class re-evaluation of near-ties)" % (
        split_layers, args.lo_n, args.lo_kgroups, refine_margin)
    for opt in os.environ.get("WANN_BENCH_OPTIONS", "").split(","):      # experiments: "skew=5,fuse_tail=0"
        if "=" in opt:
            ev.set_option(opt.split("=")[0].strip(), int(opt.split("=")[1]))
    # whole-job pool of leaf positions, this rank's contiguous slice (weak scaling: B per GPU)
    pool_n = POOL_SETS * B
    lo, hi = shard_range(pool_n * n_gpus, rank, n_gpus)
    sq_np, bl_np = make_positions(B * 4, seed=20260601 + rank)
    sq_pool = torch.from_numpy(np.tile(sq_np, (POOL_SETS // 4, 1))).to(dev)
    bl_pool = torch.from_numpy(np.tile(bl_np, POOL_SETS // 4)).to(dev)
    assert sq_pool.shape[0] == hi - lo == pool_n
    outs = [(torch.empty((B, 48), dtype=torch.uint8, device=dev),
             torch.empty((B, 48), dtype=torch.float32, device=dev),
             torch.empty((B,), dtype=torch.uint8, device=dev)) for _ in range(POOL_SETS)]

    def step(i):
        k = i % POOL_SETS
        ev.eval_topk(sq_pool[k * B:(k + 1) * B], bl_pool[k * B:(k + 1) * B], out=outs[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) --------------------------------
    def timed_device_run(evaluator, steps, first):
        """K steps between barriers, CUDA events on the launching stream, clocks sampled meanwhile,
        conv-stack kernel bracketed by its own events (option time_kernels).  Max over ranks."""
        evaluator.set_option("time_kernels", 1)
        evaluator.kernel_times()
        l0 = evaluator.stats()["kernel_launches"]
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            k = (first + i) % POOL_SETS
            evaluator.eval_topk(sq_pool[k * B:(k + 1) * B], bl_pool[k * B:(k + 1) * B], out=outs[k])
        e1.record()
        barrier()
        ms_local = e0.elapsed_time(e1)
        clk = sampler.stop()
        n_launch = evaluator.stats()["kernel_launches"] - l0
        k_ms, k_launches = evaluator.kernel_times()
        evaluator.set_option("time_kernels", 0)
        tt = torch.tensor([ms_local], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return dict(ms=float(tt.item()), ms_local=ms_local, clocks=clk, launches=n_launch, conv_ms=k_ms,
                    conv_launches=max(k_launches, 1), steps=steps)

    def roofline_of(run, peak, peak_kind):
        per_launch_pos = B * run["steps"] / run["conv_launches"]
        ach = FLOP_PER_POS_CONV * per_launch_pos / (run["conv_ms"] / run["conv_launches"] * 1e-3) / 1e12
        return ach, per_launch_pos

    for i in range(args.warmup):
        step(i)
    barrier()
    main_run = timed_device_run(ev, args.steps, args.warmup)
    ms, ms_max, clocks, launches = main_run["ms_local"], main_run["ms"], main_run["clocks"], main_run["launches"]
    conv_ms, conv_launches = main_run["conv_ms"], main_run["conv_launches"]
    value = n_gpus * B * args.steps / (ms_max * 1e-3)
    # the same measurement over >= 3 s back to back: what the power-capped clock sustains
    sustained_run = None
    if args.sustained_seconds > 0:
        s_steps = max(args.steps, int(np.ceil(args.sustained_seconds * 1e3 / (ms_max / args.steps))))
        sustained_run = timed_device_run(ev, s_steps, args.warmup + args.steps)

    # ---------------- end to end through the host-buffer API (`e2e`) ----------------------
    h_sq = torch.from_numpy(sq_np[:B].copy()).pin_memory()
    h_bl = torch.from_numpy(bl_np[:B].copy()).pin_memory()
    h_out = (torch.empty((B, 48), dtype=torch.uint8).pin_memory(),
             torch.empty((B, 48), dtype=torch.float32).pin_memory(),
             torch.empty((B,), dtype=torch.uint8).pin_memory())
    h_args = (h_sq.numpy(), h_bl.numpy(), tuple(x.numpy() for x in h_out))
    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        ev.eval_topk(h_args[0], h_args[1], out=h_args[2])
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ev.eval_topk(h_args[0], h_args[1], out=h_args[2])      # returns when the priors are on the host
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * B * e2e_steps / float(t.item())

    # same, returning the full [N,155] probability rows (what the reference's evaluate() returns)
    h_probs = torch.empty((B, 155), dtype=torch.float32).pin_memory()
    for _ in range(2):
        ev.eval_probs(h_args[0], h_args[1], out=h_probs.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ev.eval_probs(h_args[0], h_args[1], out=h_probs.numpy())
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_probs_value = n_gpus * B * e2e_steps / float(t.item())

    # ---------------- BASELINE configs[4]: 8,192 independent game trees over 8 GPUs = 1,024 per GPU,
    # one leaf per tree per step, device-resident steady state through ONE C-ABI call per step
    # (wann_tree_step: evaluate -> choose a child -> apply it in place -> restart finished games)
    from wann_b200.treefarm import TreeFarm
    trees = {}
    for n_trees in (1024, 16384):
        farm = TreeFarm(ev, n_trees, device=dev, seed=17 + rank)
        for _ in range(40):
            farm.step_fused()                                   # spread the games over all plies
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tsteps = 200
        a.record()
        for _ in range(tsteps):
            farm.step_fused()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        trees[str(n_trees)] = {"trees_per_gpu": n_trees, "steps": tsteps,
                               "us_per_step": round(float(t.item()) * 1e3 / tsteps, 2),
                               "positions_per_s": round(n_gpus * n_trees * tsteps / (float(t.item()) * 1e-3), 1)}

    # ---------------- configs[4] with the REFERENCE'S OWN SEARCH per tree: a forest of array trees
    # (node-for-node the compiled reference's trees), 1,024 trees per GPU (and 8,192), the headline arithmetic.
    # DEVICE-RESIDENT forest: trees in HBM, one warp per tree, each
    # lock step = tree kernel -> conv stack -> child seeding -> tree kernel on one stream (wann_forest_run);
    # and, for comparison, the HOST forest (OpenMP tree code, one host-buffer evaluation call per step).
    forest_line = None
    if not args.no_sweep:
        from wann_b200 import SearchForest
        from wann_b200.forest import host_threads
        host_threads(max(1, (os.cpu_count() or 1) // max(1, world)))      # torchrun exports OMP_NUM_THREADS=1
        ev3 = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), device=local_rank, mode="fp16c")
        forest_line = {"arithmetic": "fp16c as created (the headline arithmetic), no refinement pass"}
        for key, n_trees, resident, sims in (("device_resident", 1024, True, 24), ("device_resident_8192", 8192, True, 12),
                                             ("host", 1024, False, 10)):
            fsq, fbl = sq_np[:n_trees], bl_np[:n_trees]
            forest = SearchForest(fsq, fbl, evaluator=ev3, device_resident=resident, node_capacity=16384)
            forest.run(ev3, 1)                                          # warm-up: first expansions, lane allocation
            barrier()
            t0 = time.perf_counter()
            fsteps, fevals = forest.run(ev3, sims)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sample = range(0, n_trees, max(1, n_trees // 64))
            nodes = float(np.mean([forest.info(t)["nodes"] for t in sample])) * n_trees
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            cnts = torch.tensor([fevals, nodes], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
            forest_line[key] = {"trees_per_gpu": n_trees, "simulations_per_tree": sims - 1, "lock_steps": fsteps,
                                "evaluations": int(cnts[0].item()), "tree_nodes_estimate": int(cnts[1].item()),
                                "evaluations_per_s": round(float(cnts[0].item()) / float(t.item()), 1),
                                "us_per_lock_step": round(dt * 1e6 / max(fsteps, 1), 1)}
            if not resident:
                forest_line[key]["host_threads_per_rank"] = host_threads()
            forest.close()
            # steady state (configs[4]: "each GPU serving its own leaf batches at steady state"): every tree keeps
            # searching -- a budget of lock steps instead of simulations, as the reference's search has a time budget
            forest = SearchForest(fsq, fbl, evaluator=ev3, device_resident=resident, node_capacity=32768)
            forest.run_steps(ev3, 40)
            barrier()
            t0 = time.perf_counter()
            ssteps, sevals = forest.run_steps(ev3, 150)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            forest.run(ev3, 0)                                          # finish the simulations in flight
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            cnts = torch.tensor([sevals], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
            forest_line[key]["steady_state"] = {"lock_steps": ssteps, "evaluations": int(cnts[0].item()),
                                                "evaluations_per_s": round(float(cnts[0].item()) / float(t.item()), 1),
                                                "rows_per_step": round(sevals / max(ssteps, 1), 1),
                                                "us_per_lock_step": round(dt * 1e6 / max(ssteps, 1), 1)}
            forest.close()
        forest_line["evaluations_per_s"] = forest_line["device_resident"]["steady_state"]["evaluations_per_s"]
        forest_line["note"] = ("evaluations_per_s = the device-resident forest of 1,024 trees per GPU at steady state; the "
                               "fixed-simulation runs beside it end with ever fewer active trees per lock step")
        ev3.close()

    # ---------------- the other arithmetic modes on the same workload (device resident, short runs):
    # faster and less accurate (bf16, fp16), or slower (margin refinement, lo weight stages: what the headline
    # mode needed before its weights were rounded with error compensation)
    modes = None
    if not args.no_sweep:
        modes = {}
        for name, mode, split, margin in (("bf16", "bf16", 0, 0.0), ("fp16", "fp16", 0, 0.0),
                                          ("fp16+refine0.02", "fp16", 0, 0.02),
                                          ("fp16c+refine0.01", "fp16c", 0, 0.01),
                                          ("fp16c+rtn", "fp16c", 0, 0.0),               # round-to-nearest weights
                                          ("fp16c+rtn+split14", "fp16c", 14, 0.0),      # + partial hi/lo split
                                          ("fp16c+rtn+fullsplit8", "fp16c", 8 | 192 << 8 | 3 << 16, 0.0)):
            split, lo_n, lo_kg = split & 255, (split >> 8) & 255 or 64, (split >> 16) or 1
            ev2 = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), device=local_rank, mode=mode,
                                            refine_margin=margin)
            if "+rtn" in name:
                ev2.set_option("gptq", 0)
            if split:
                ev2.set_option("lo_n", lo_n)
                ev2.set_option("lo_kgroups", lo_kg)
                ev2.set_option("split_layers", split)
            for i in range(3):
                ev2.eval_topk(sq_pool[i * B:(i + 1) * B], bl_pool[i * B:(i + 1) * B], out=outs[i])
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rsteps = max(3, min(args.steps, 40))
            a.record()
            for i in range(rsteps):
                k = (3 + i) % POOL_SETS
                ev2.eval_topk(sq_pool[k * B:(k + 1) * B], bl_pool[k * B:(k + 1) * B], out=outs[k])
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rec = {"mode": mode, "gptq": int(mode == "fp16c" and "+rtn" not in name), "split_layers": split,
                   "lo_n": lo_n if split else 0, "lo_kgroups": lo_kg if split else 0,
                   "refine_margin_logits": margin,
                   "value": n_gpus * B * rsteps / (float(t.item()) * 1e-3), "unit": UNIT, "steps": rsteps}
            if margin:
                ev2.eval_topk(sq_np[:B], bl_np[:B])                  # host calls read the device-side counter back:
                r0 = ev2.stats()["refined_positions"]                # the first drains the timed loop's total,
                ev2.eval_topk(sq_np[:B], bl_np[:B])                  # the second counts one batch
                rec["refined_fraction"] = round((ev2.stats()["refined_positions"] - r0) / float(B), 5)
            modes[name] = rec
            ev2.close()

    # ---------------- the HBM-bound byte kernels beside the path (north_star: "achieved HBM GB/s for
    # encode and epilogue"): encode 65 B -> 256 B, legal mask 65 B -> 155 B per position, device resident
    hbm_kernels = None
    if rank == 0 and not args.no_sweep:
        import ctypes
        from wann_b200 import _lib as L
        nbk = 2 * pool_n                                  # 4,194,304 positions: 1.3 / 0.9 GB of traffic per launch
        bk_sq, bk_bl = torch.cat([sq_pool, sq_pool]), torch.cat([bl_pool, bl_pool])
        d_planes = torch.empty((nbk, 8, 8, 4), dtype=torch.int8, device=dev)
        d_mask = torch.empty((nbk, 155), dtype=torch.uint8, device=dev)
        hbm_peak = 6533.5
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            hbm_peak = json.load(open(pk)).get("hbm_gbs", hbm_peak)
        hbm_kernels = {"positions": nbk, "peak_gbs": hbm_peak}
        s1 = ctypes.c_void_p(1)
        for name, fn, out_t, bpp in (("encode_kernel", ev._lib.wann_encode, d_planes, 65 + 256),
                                     ("legal_mask_kernel", ev._lib.wann_legal_mask, d_mask, 65 + 155)):
            call = lambda: L.check(fn(ev._ctx, bk_sq.data_ptr(), bk_bl.data_ptr(), nbk, out_t.data_ptr(),  # noqa: E731
                                      L.MEM_DEVICE, s1))
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                call()
            b.record()
            torch.cuda.synchronize()
            sec = a.elapsed_time(b) / 10 * 1e-3
            hbm_kernels[name] = {"bytes_per_position": bpp, "us": round(sec * 1e6, 2),
                                 "achieved_gbs": round(nbk * bpp / sec / 1e9, 1),
                                 "frac": round(nbk * bpp / sec / 1e9 / hbm_peak, 4)}
        del d_planes, d_mask, bk_sq, bk_bl
        # child construction (65+49 B in, 3,120 B out per position) and child seeding (65+48+192+1 B in,
        # 768+48+32 B out) on already-ranked positions (wann_expand_ranked)
        nck = min(pool_n, 1 << 17)
        d_idx = torch.empty((nck, 48), dtype=torch.uint8, device=dev)
        d_pri = torch.empty((nck, 48), dtype=torch.float32, device=dev)
        d_cnt = torch.empty((nck,), dtype=torch.uint8, device=dev)
        ev.eval_topk(sq_pool[:nck], bl_pool[:nck], out=(d_idx, d_pri, d_cnt))
        d_kids = torch.empty((nck, 48, 64), dtype=torch.int8, device=dev)
        d_flg = torch.empty((nck, 48), dtype=torch.uint8, device=dev)
        d_st = torch.empty((nck, 48, 4), dtype=torch.int32, device=dev)
        d_ps = torch.empty((nck, 8), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for name, kids_t, st_t, bpp in (("expand_children_kernel", d_kids, None, 65 + 49 + 48 * 64 + 48),
                                        ("seed_children_kernel", None, d_st, 65 + 48 + 192 + 1 + 768 + 48 + 32)):
            call = lambda: L.check(ev._lib.wann_expand_ranked(   # noqa: E731
                ev._ctx, sq_pool.data_ptr(), bl_pool.data_ptr(), nck, d_idx.data_ptr(), d_pri.data_ptr(), d_cnt.data_ptr(),
                kids_t.data_ptr() if kids_t is not None else None, d_flg.data_ptr(),
                st_t.data_ptr() if st_t is not None else None, d_ps.data_ptr() if st_t is not None else None, s1))
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                call()
            b.record()
            torch.cuda.synchronize()
            sec = a.elapsed_time(b) / 10 * 1e-3
            hbm_kernels[name] = {"positions": nck, "bytes_per_position": bpp, "us": round(sec * 1e6, 2),
                                 "achieved_gbs": round(nck * bpp / sec / 1e9, 1),
                                 "frac": round(nck * bpp / sec / 1e9 / hbm_peak, 4)}
        del d_kids, d_st

    # ---------------- optional: counters gathered with NCCL, sweep, CPU baseline ----------
    counters = gather_counters([B * args.steps, int(ms * 1e6)], device=dev)
    # configs[3]: the batch sweep 64 ... 65,536 positions PER GPU (plus 1 = configs[0], 256 = configs[2],
    # 1,024 = configs[1]), every rank on its own slice, device pointers, asynchronous back-to-back calls;
    # whole-job positions/s from the max-over-ranks CUDA-event time
    sweep = {}
    if not args.no_sweep:
        for nb in (1, 4, 16, 64, 256, 1024, 4096, 8192, 16384, 65536):   # policy_net_metrics.py:39-100
            o = tuple(x[:nb] for x in outs[0])
            for _ in range(3):
                ev.eval_topk(sq_pool[:nb], bl_pool[:nb], out=o)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            a.record()
            for r in range(reps):
                off = ((r * 4099) % (pool_n - nb)) if nb < pool_n else 0
                ev.eval_topk(sq_pool[off:off + nb], bl_pool[off:off + nb], out=o)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sweep[str(nb)] = round(n_gpus * nb * reps / (float(t.item()) * 1e-3), 1)
    if world > 1:
        dist.barrier()

    if rank == 0:
        burst, sustained, src = peaks()
        traffic, traffic_src = None, None
        summ = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(summ):
            rec = json.load(open(summ)).get("conv_stack_tc", {})
            traffic = rec.get("dram_bytes_per_launch")
            traffic_src = ("profiles/ncu_summary.json: %s (static value from the committed ncu --set full "
                           "capture, NOT measured in this run)" % rec.get("source", "r01 capture, bf16 mode"))

        def roof(run):
            ach, per_launch = roofline_of(run, None, None)
            # `frac` is ALWAYS against the burst figure, the higher of the two measured peaks: in a region of
            # >= 1 s this kernel runs above cuBLAS's own sustained rate under the same power cap (a fraction
            # above 1 says the denominator is not a ceiling), so the sustained figure is only shown beside it
            peak = burst
            return {"bound": "tensor", "kernel": "conv_stack_tc_kernel<%s>" % args.mode, "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_kind": "%s cuBLAS bf16 burst, the conservative denominator (timed region %.2f s; cuBLAS itself "
                                 "sustains %.1f TFLOP/s over 4 s under the same power cap: frac_of_sustained_peak)" % (
                                     src, run["ms"] * 1e-3, sustained),
                    "frac_of_burst_peak": ach / burst, "frac_of_sustained_peak": ach / sustained,
                    "burst_peak": burst, "sustained_peak": sustained,
                    "flop_per_position": FLOP_PER_POS_CONV,
                    "flop_note": "algorithmic FLOPs of conv1..conv5; the lo-weight MMAs of the partial hi/lo split "
                                 "(split_layers=%d, lo_n=%d, lo_kgroups=%d) and the refinement pass are overhead and not counted" % (
                                     split_layers, args.lo_n, args.lo_kgroups),
                    "positions_per_launch": per_launch,
                    "kernel_ms_per_launch": run["conv_ms"] / run["conv_launches"],
                    "kernel_share_of_step": run["conv_ms"] / run["ms_local"]}
        roofline = roof(main_run)
        roofline["traffic"] = traffic
        roofline["traffic_source"] = traffic_src
        sustained_rec = None
        if sustained_run is not None:
            sustained_rec = {"value": n_gpus * B * sustained_run["steps"] / (sustained_run["ms"] * 1e-3), "unit": UNIT,
                             "steps": sustained_run["steps"], "seconds": sustained_run["ms"] * 1e-3,
                             "clocks": sustained_run["clocks"], "roofline": roof(sustained_run)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16" if args.mode in ("fp16", "fp16c", "fp32_tc") else "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[3] batch sweep, top batch: %d positions per GPU per step, "
                                   "full path encode->conv stack->FC->softmax->legal gather+sort "
                                   "(ranked child priors out)" % B,
                       "mode": args.mode, "split_layers": split_layers, "lo_n": args.lo_n if split_layers else 0,
                       "lo_kgroups": args.lo_kgroups if split_layers else 0, "refine_margin_logits": refine_margin,
                       "arithmetic": arithmetic,
                       "accuracy": "this configuration on 100,000 positions vs the fp32 graph: see profiles/r02_accuracy_gptq_seed{1,2,3}.json "
                                   "and tests/test_gpu_parity.py::test_centred_fp16_mode_meets_the_north_star_gates_on_100k_positions",
                       "weights": "synthetic trained_like seed 1 (reference weight shard absent)",
                       "positions": "seeded random legal play-out positions",
                       "l2": "inputs/outputs rotate over %d buffer sets (%.0f MB) > 126 MB L2; 2 MB of "
                             "weights stay L2-resident by design" % (POOL_SETS, POOL_SETS * B * (IN_BYTES + OUT_BYTES) / 1e6),
                       "parallelism": "replicas x%d, leaf batches sharded per GPU, no collective" % n_gpus},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * IN_BYTES,
                    "d2h_bytes_per_step": B * OUT_BYTES, "api": "wann_eval_topk(WANN_MEM_HOST), pinned buffers",
                    "mode": args.mode, "split_layers": split_layers, "refine_margin_logits": refine_margin,
                    "note": "may exceed `value`: the host path runs 16,384-position chunks double-buffered on two "
                            "streams (copies hidden under the kernels), the power-capped kernel is faster per position "
                            "at 16,384 than at 65,536 positions per launch (see sweep_positions_per_s), and `value` "
                            "is taken with per-launch event bracketing (time_kernels) on"},
            "e2e_probs": {"value": e2e_probs_value, "unit": UNIT, "h2d_bytes_per_step": B * IN_BYTES,
                          "d2h_bytes_per_step": B * 155 * 4, "mode": args.mode,
                          "api": "wann_eval_probs(WANN_MEM_HOST): full [N,155] rows, the reference's evaluate() output"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "sustained": sustained_rec,
            "sweep_positions_per_s": sweep,
            "config5_game_trees": trees,
            "modes": modes,
            "config5_search_forest": forest_line,
            "hbm_kernels": hbm_kernels,
            "per_rank_counters": counters.tolist(),
        }
        if not args.no_cpu_baseline and n_gpus == 1:
            v, cores, kind, sample = cpu_arm()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                    "note": "CPU stand-in for the reference's TF-1.0 session "
                                            "(TensorFlow/weights not available); reported baseline only"}
        emit(line)
    ev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

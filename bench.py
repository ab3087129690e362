#!/usr/bin/env python
"""bench.py -- TS samples/sec over the full Langevin-dynamics trajectory (BASELINE.json metric).

A bench "step" is ONE PASS OF THE HOT PATH OVER ONE BATCH: the full `--ld-steps` (5000)
Langevin trajectory of one batch of `--batch` (100) synthetic Grambow-shaped reactions
(BASELINE.json configs[1]); value = reactions carried through the whole trajectory per
second, summed over ranks.  One process per GPU; reactions are sharded (every rank samples
its own batch: weak scaling), there is no collective on the data path.

  python bench.py [--gpus N --steps K --warmup W]           our CUDA path
  python bench.py --impl reference ...                      the reference algorithm on host cores
                                                            (oracle port; rank 0 only)
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "TS samples/sec (full LD trajectory)"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ld-steps", type=int, default=5000, help="Langevin steps per trajectory (sampling.py default)")
    ap.add_argument("--batch", type=int, default=100, help="reactions per GPU")
    ap.add_argument("--network", default="condensenc", choices=["condensenc", "dualenc"])
    ap.add_argument("--math", default=os.environ.get("TSDIFF_B200_MATH", "tf32"), choices=["fp32", "tf32"],
                    help="tf32: tcgen05 tensor cores, fp32 accumulate (DESIGN.md section 4 bounds); fp32: FFMA strict parity")
    ap.add_argument("--ref-ld-steps", type=int, default=8, help="sampler steps per bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_inputs(args, rank):
    from tsdiff_b200.synthetic import make_batch
    return make_batch(args.batch, seed=1000 + rank)


def make_models(args, device):
    from tsdiff_b200.config import QM9_DEFAULT_MODEL, TRAIN_CONFIG_MODEL
    from tsdiff_b200.models.epsnet import get_model
    cfg = TRAIN_CONFIG_MODEL if args.network == "condensenc" else QM9_DEFAULT_MODEL
    torch.manual_seed(0)
    m = get_model(cfg)
    if device is not None:
        m = m.to(device)
        m.math = args.math
    return m, cfg


# the LD knobs: sampling.py defaults for path B; path A needs explicit clips at random init
def ld_kwargs(args):
    if args.network == "condensenc":
        return dict(step_lr=1e-7, clip=1000)
    return dict(step_lr=1e-7, clip=10.0, clip_local=10.0)


# ----------------------------------------------------------------------------- reference arm
def oracle_sample_seconds(args, data, n_sampler_steps, repeats):
    """Times `repeats` bounded samples (the first n_sampler_steps Langevin steps of the real
    trajectory) of the reference algorithm restated in oracle/ on the host cores."""
    from oracle import tsdiff_oracle as O
    m, cfg = make_models(args, None)
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(7)
    noise = torch.randn(n_sampler_steps, data["atom_type"].numel(), 3, generator=gen)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            if args.network == "condensenc":
                O.dynamic_sampling_ld([params], cfg, data["atom_type"], data["r_feat"], data["p_feat"],
                                      data["pos_init"], data["bond_index"], data["bond_type"], data["batch"],
                                      n_sampler_steps, noise=noise, keep_traj=False, **ld_kwargs(args))
            else:
                O.dualenc_ld_sample(dict(params), cfg, data["atom_type"], data["pos_init"], data["bond_index"],
                                    data["bond_type"], data["batch"], n_sampler_steps, noise=noise, keep_traj=False,
                                    **ld_kwargs(args))
        times.append(time.perf_counter() - t0)
    return times


def workload_config(args, world):
    return {"workload": "full %d-step ld sampling, single checkpoint, batch_size %d synthetic Grambow-shape reactions "
                        "(10-25 atoms) per GPU" % (args.ld_steps, args.batch),
            "network": args.network, "weights": "random-init seed 0", "reactions_per_gpu": args.batch,
            "ld_steps": args.ld_steps, "math": args.math, "sharding": "reactions x%d, no collective" % world}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    data = build_inputs(args, 0)
    n = args.ref_ld_steps
    times = oracle_sample_seconds(args, data, n, args.warmup + args.steps)[args.warmup:]
    per_sampler_step = sum(times) / len(times) / n
    traj_seconds = per_sampler_step * args.ld_steps
    value = args.batch / traj_seconds
    sample = ("first %d of %d Langevin steps per bench step, reference-faithful dense graph build every step; "
              "extrapolated linearly to the full trajectory" % (n, args.ld_steps))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": traj_seconds * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args, 1),
           "us_per_eps_step": per_sampler_step * 1e6,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------- our arm
def flush_l2(buf):
    buf.zero_()


def build_runner(args, model, data_dev, keep_traj=True):
    """Device-resident sampler state: engine + CUDA-graph runner, inputs already in HBM."""
    from tsdiff_b200 import engine as E
    d = data_dev
    if args.network == "condensenc":
        eng = E.CondensedScoreEngine([model], d["atom_type"], d["r_feat"], d["p_feat"], d["bond_index"],
                                     d["bond_type"], d["batch"], math=args.math)
        sched, sigmas = E.ld_schedule(model.alphas, args.ld_steps, 1e-7)
        ch0, ch1 = eng.score_channels(1000)
    else:
        eng = E.DualScoreEngine(model, d["atom_type"], d["bond_index"], d["bond_type"], d["batch"], math=args.math)
        sched, sigmas = E.ld_schedule(model.alphas, args.ld_steps, 1e-7)
        ch0, ch1 = eng.score_channels(10.0, 10.0, 0.2)
    pos = (d["pos_init"] * sigmas[-1].to(d["pos_init"].device)).contiguous()
    runner = E.LangevinRunner(eng, ch0, ch1, sched, pos, seed=2022, keep_traj=keep_traj)
    return eng, runner


def api_call(args, model, data_host, device):
    """The call a user makes (sampling.py:169-209): host tensors -> device, dynamic_sampling /
    langevin_dynamics_sample, trajectory + final positions back on the host."""
    from tsdiff_b200.models.sampler import EnsembleSampler
    d = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in data_host.items()}
    if args.network == "condensenc":
        ens = EnsembleSampler([model])
        pos, traj = ens.dynamic_sampling(d["atom_type"], d["r_feat"], d["p_feat"], d["pos_init"], d["bond_index"],
                                         d["bond_type"], d["batch"], data_host["num_graphs"], extend_order=True,
                                         n_steps=args.ld_steps, sampling_type="ld", seed=2022, **ld_kwargs(args))
    else:
        pos, traj = model.langevin_dynamics_sample(d["atom_type"], d["pos_init"], d["bond_index"], d["bond_type"],
                                                   d["batch"], data_host["num_graphs"], extend_order=True,
                                                   n_steps=args.ld_steps, sampling_type="ld", seed=2022,
                                                   **ld_kwargs(args))
    return pos.cpu(), traj


def committed_traffic():
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the top kernels, from
    the committed `ncu --set full` capture (profiles/r1_ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def kernel_rooflines(args, eng, peaks, device):
    """Live CUDA-event timing (kernel alone, L2 flushed) of the two kernels that matter, on the
    engine's current late-trajectory edge list:
      * the dominant kernel: the fused filter network of one CFConv layer (two chained
        E x H x H tensor-core GEMMs; the same kernel also runs the node update) -> tensor roofline
      * the CFConv segmented aggregation (the HBM-bound message-passing kernel) -> HBM roofline."""
    from tsdiff_b200 import _lib as L
    lib = L.load()
    plan, h = eng.plan, eng.hidden
    e = plan.work_count()  # rows of the per-edge kernels: unordered pairs (path B) or directed edges
    e_dir = plan.edge_count()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # L2 flush = write pass + read pass over 256 MiB each: the read pass evicts the dirty lines of
    # the write pass, so their write-back does not compete with the timed kernel
    flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flush_r = torch.zeros(64 << 20, dtype=torch.float32, device=device)
    ws = eng.ws
    x = ws.edge[2]  # edge_attr of the last evaluation (E_cap, H)
    tmp, out = ws.edge[4], ws.edge[5]
    blocks = eng.members[0]["blocks"] if hasattr(eng, "members") else eng.blocks
    math = L.MATH[args.math]
    traffic = committed_traffic()

    def timed(fn, reps=10):
        ts = []
        for _ in range(3):
            fn()
        for _ in range(reps):
            flush_w.zero_()
            flush_r.sum()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            fn()
            t1.record()
            torch.cuda.synchronize()
            ts.append(t0.elapsed_time(t1) * 1e-3)
        return sum(ts) / len(ts)

    t_f = timed(lambda: L.check(lib.tsd_filter_network(C.byref(plan.c_work_batch), C.byref(plan.c_work_edges), L.ptr(x),
                                                       C.byref(blocks[0]), L.ptr(tmp), L.ptr(out), math, stream),
                                "tsd_filter_network"))
    # algorithmic = what the reference computes for this launch: two E x H x H layers over the DIRECTED edges
    # (SURVEY.md 8d); executed = the same over the unordered pairs (both directions share one row)
    flops = 2 * 2.0 * e_dir * h * h
    executed = 2 * 2.0 * e * h * h
    kname = "k_chain_tf32" if args.math == "tf32" else "k_gemm_ffma"
    tensor = {"bound": "tensor", "kernel": "%s: CFConv filter network nn2(ssp(nn0(edge_attr)))*C, 2 x (E x %d x %d), %s"
                                           % (kname, h, h, args.math),
              "achieved": flops / t_f / 1e12, "peak": peaks["tensor_burst"], "unit": "TFLOP/s",
              "frac": flops / t_f / 1e12 / peaks["tensor_burst"], "traffic": traffic.get(kname),
              "peak_source": peaks["source"] + " bf16 burst (tf32 tensor peak is half of it)",
              "us_per_launch": t_f * 1e6, "rows": e, "algorithmic_flops": flops, "executed_flops": executed,
              "executed_tflops": executed / t_f / 1e12,
              "note": "achieved = reference-equivalent flops (directed edges) / time; the kernel executes them once per "
                      "unordered pair" if e != e_dir else "no dedup",
              "algorithmic_bytes": 2 * e_dir * h * 4 + 2 * h * h * 4}
    x1 = ws.node[1]
    agg = ws.node[2]
    t_agg = timed(lambda: L.check(lib.tsd_cfconv_aggregate(C.byref(plan.c_work_batch), C.byref(plan.c_work_edges), h, L.ptr(x1),
                                                           L.ptr(out), L.ptr(agg), stream), "tsd_cfconv_aggregate"))
    n = plan.num_nodes
    # algorithmic: one filter row per directed edge + x1 once + output + CSR; with pair sharing
    # the filter rows live in a half-size buffer, so the DRAM traffic is lower than that
    nbytes = e_dir * h * 4 + 2 * n * h * 4 + e_dir * 8 + (n + 1) * 4
    hbm = {"bound": "hbm", "kernel": "k_cfconv_aggregate", "achieved": nbytes / t_agg / 1e9, "peak": peaks["hbm"],
           "unit": "GB/s", "frac": nbytes / t_agg / 1e9 / peaks["hbm"], "traffic": traffic.get("k_cfconv_aggregate"),
           "peak_source": peaks["source"], "us_per_launch": t_agg * 1e6, "algorithmic_bytes": nbytes}
    return tensor, hbm


def run_ours(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tsdiff_b200 has no CPU path (use --impl reference for the "
                         "host-core baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from tsdiff_b200 import _lib as L
    lib = L.load()
    peaks = measured_peaks()
    data = build_inputs(args, rank)
    model, cfg = make_models(args, device)
    data_dev = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}

    # ---- device-resident throughput: inputs in HBM, one replayed CUDA graph per Langevin step
    eng, runner = build_runner(args, model, data_dev)
    c0 = lib.tsd_launch_count()
    runner.use_graph, saved = False, runner.use_graph
    runner._one_step()  # eager step: counts our kernel launches per Langevin step
    torch.cuda.synchronize()
    launches_per_ld_step = lib.tsd_launch_count() - c0
    runner.use_graph = saved
    runner._reset()
    runner.prepare()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def one_trajectory():
        runner._reset()
        flush_l2(flush)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        runner.run()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) * 1e-3

    for _ in range(args.warmup):
        one_trajectory()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    times = [one_trajectory() for _ in range(args.steps)]
    barrier()
    clock_info = clocks.stop()
    elapsed = torch.tensor([sum(times)], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    mean_edges = None
    value = world * args.batch * args.steps / elapsed

    # ---- end to end through the public API: host inputs in, trajectory + positions out
    e2e = None
    if not args.no_e2e:
        pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}
        h2d = sum(v.numel() * v.element_size() for v in pinned.values() if torch.is_tensor(v))
        api_call(args, model, pinned, device)  # warm-up (untimed)
        iters = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            pos_host, traj = api_call(args, model, pinned, device)
        torch.cuda.synchronize()
        t_api = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t_api, op=dist.ReduceOp.MAX)
        d2h = pos_host.numel() * 4 + sum(t.numel() * 4 for t in traj)
        e2e = {"value": world * args.batch * iters / float(t_api.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "iters": iters,
               "includes": "H2D inputs, bond-order tables, graph capture, %d LD steps, D2H trajectory+positions"
                           % args.ld_steps}

    if rank != 0:
        return
    # ---- per-kernel rooflines (live, rank 0) and the CPU baseline (N = 1 only)
    runner._reset()
    runner.run(n_steps=min(args.ld_steps, 2000))  # a late-trajectory edge list (every pair inside the cutoff)
    mean_edges = eng.plan.edge_count()
    tensor_roof, hbm_roof = kernel_rooflines(args, eng, peaks, device)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        t_probe = oracle_sample_seconds(args, data, 1, 1)[0]
        n = int(max(2, min(40, args.cpu_seconds / max(t_probe, 1e-3))))
        t_sample = oracle_sample_seconds(args, data, n, 1)[0]
        per_step = t_sample / n
        cpu = {"value": args.batch / (per_step * args.ld_steps), "unit": UNIT, "cores": torch.get_num_threads(),
               "kind": "port", "us_per_eps_step": per_step * 1e6,
               "sample": "first %d of %d Langevin steps of the same batch on the host cores (oracle port of the "
                         "reference algorithm, dense graph build every step), extrapolated linearly" % (n, args.ld_steps)}
    ms_per_step = elapsed / args.steps * 1e3
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32" if args.math == "fp32" else "tf32", "data": "synthetic",
           "config": dict(workload_config(args, world), l2="flushed between timed trajectories (256 MiB write)",
                          edges_late_trajectory=mean_edges, nodes=int(data["atom_type"].numel()),
                          edge_capacity=eng.plan.edge_capacity,
                          network_rows_late_trajectory=eng.plan.work_count(),
                          dedup="per-edge networks run once per unordered atom pair (both directions are "
                                "bit-identical)" if eng.plan.upairs else "none"),
           "us_per_eps_step": ms_per_step * 1e3 / args.ld_steps,
           "gpu_launches": int(launches_per_ld_step) * args.ld_steps * args.steps,
           "launches_per_ld_step": int(launches_per_ld_step), "clocks": clock_info, "e2e": e2e,
           "roofline": tensor_roof, "roofline_message_passing": hbm_roof, "cpu_baseline": cpu,
           "published_reference_datum": "<=39.5 ms/step, ~0.51 samples/s (DDPM, unnamed GPU; BASELINE.md)"}
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Langevin-step time (CUDA-graph replays, late-trajectory edge count) for the filter-stack launch modes and node tiles
in ONE process: the tuning hooks are run-time switches, every mode re-captures its graph.
usage: python profiles/scripts/stack_modes.py [mode:tile[:pdl[:stack_ctas[:gemm_pdl[:atoms_per_node_cluster[:chain2[:stack2_ctas[:node_chain]]]]]]] ...]   (mode -1 = one filter kernel per block)"""
import sys, json, torch
sys.path.insert(0, '.')
import bench
from tsdiff_b200 import _lib as L

class A: pass
args = A(); args.batch = 100; args.network = 'condensenc'; args.math = 'tf32'; args.ld_steps = 5000; args.members = 1; args.mode = 'shard'
dev = torch.device('cuda:0')
lib = L.load()
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
dd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
specs = sys.argv[1:] or ["-1:0", "0:0"]
ref_pos = None
for spec in specs:
    f = [int(x) for x in spec.split(":")]
    mode, tile, pdl = f[0], f[1], (f[2] if len(f) > 2 else 1)
    lib.tsd_tune_node_pdl(pdl)
    grid = f[3] if len(f) > 3 else 0
    lib.tsd_tune_filter_stack_grid(grid)
    gpdl = f[4] if len(f) > 4 else 1
    lib.tsd_tune_gemm_pdl(gpdl)
    npc = f[5] if len(f) > 5 else 0
    lib.tsd_tune_node_npc(npc)
    chain2 = f[6] if len(f) > 6 else 1
    lib.tsd_tune_gemm_chain2(chain2)
    ctas2 = f[7] if len(f) > 7 else 0
    lib.tsd_tune_filter_stack_ctas2(ctas2)
    nchain = f[8] if len(f) > 8 else 0
    lib.tsd_tune_node_chain(nchain)
    lib.tsd_tune_filter_stack(mode)
    lib.tsd_tune_node_tile(tile)
    torch.manual_seed(0)
    eng, runner = bench.build_runner(args, [model], dd, keep_traj=False)
    runner.prepare(); runner.run(n_steps=1500); torch.cuda.synchronize()
    pos = runner.pos.clone() if hasattr(runner, "pos") else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(1500): runner.graph.replay()
    t1.record(); torch.cuda.synchronize()
    out = {"mode": mode, "tile": tile, "pdl": pdl, "stack_ctas": grid, "gemm_pdl": gpdl, "npc": npc, "chain2": chain2, "stack2_ctas": ctas2, "node_chain": nchain, "step_us": t0.elapsed_time(t1) / 1500 * 1e3,
           "pairs": eng.plan.work_count()}
    if pos is not None:
        if ref_pos is None: ref_pos = pos
        out["max_abs_pos_diff_vs_first_mode"] = float((pos - ref_pos).abs().max())
    out["node_chain_flag"] = lib.tsd_node_chain_flag()
    print(json.dumps(out), flush=True)
    del eng, runner

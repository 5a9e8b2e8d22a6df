#!/usr/bin/env python3
"""bench.py -- MCMC mutations/sec of the LMC chain loop on the bundled torus scene.

Workload (BASELINE.json configs[1]): torus scene, LMC (mala) mutation, path length (maxdepth) 8,
2^20 Markov chains per GPU, global cache off (every eligible MALA step evaluates its PSS
gradient), options of scenes/torus/lmc.xml.  One "step" = every chain of the job advanced by
MUTATIONS_PER_STEP iterations of the loop at src/mlt.cpp:91-170 (one `lmc_run_chains` call).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one rank per GPU)
    python bench.py --impl reference ...                      CPU arm: the reference's chain loop as
                                                              restated by the oracle, all host cores

Timed region (device arm): K steps bracketed by barrier + cuda synchronize, CUDA events on the
launching stream, max over ranks; it includes the final NCCL all-reduce of the fp32 film
(src/mlt.cpp:57->200 is the span the metric is defined on; MLTInit, scene load, BVH build are
setup).  Chain records + wavefront queues (4.5 KB + 1.4 KB per chain, x 2^20 = 6 GB) are far larger
than L2, so no explicit flush.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "langevin-mcmc_b200")
SCENE_XML = os.path.join(ROOT, "scenes", "torus", "lmc.xml")
METRIC = "MCMC mutations/sec (torus, path len 8)"
MAXDEPTH = 8
CHAINS_PER_GPU = 1 << 20
MUTATIONS_PER_STEP = 32
# canonical state words S_LMC(L) = 14 L + 21; bytes per mutation = 8 S + 48
ALGO_BYTES_PER_MUTATION = 8 * (14 * MAXDEPTH + 21) + 48   # 1112 B at L = 8
# dram__bytes_read.sum + dram__bytes_write.sum of the 48 kernels of one iteration over 2^20 chains
# (ncu, profiles/r01_launches_pervertex_2p20.csv / _summary.txt): 11.86 GB / 2^20 mutations
DRAM_BYTES_PER_MUTATION_NCU = 11.86e9 / (1 << 20)


def load_package():
    spec = importlib.util.spec_from_file_location("lmc_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lmc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6550.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def cpu_chain_rate(threads, budget_s, maxdepth=MAXDEPTH):
    """Times the CPU oracle's chain loop (restatement of src/mlt.cpp:60-196, same options, cache
    off) on `threads` host threads for roughly budget_s seconds.  Returns (mut/s, sample text)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle
    o = Oracle()
    h = o.load(SCENE_XML)
    o.set_option(h, "maxdepth", maxdepth)
    # when the reference's generated gradient code was built (oracle/_ref), the CPU arm evaluates
    # gradients with it (reverse mode, as the reference does) instead of the twin's evaluator
    ref_grad = o.use_reference_gradient(True) > 0
    chains, steps = 256 * threads, 32
    norm, init_ls = o.mlt_init(h, 300000, chains, 32)
    t0 = time.time()
    o.run_chains(h, chains, steps, norm, init_ls, threads=threads, want_trace=False, samples_per_chain=steps)
    dt = time.time() - t0
    rate = chains * steps / dt
    # size the measured run from the probe
    steps2 = int(max(32, min(4096, budget_s * rate / chains)))
    t0 = time.time()
    o.run_chains(h, chains, steps2, norm, init_ls, threads=threads, want_trace=False, samples_per_chain=steps2)
    dt = time.time() - t0
    return chains * steps2 / dt, "%d chains x %d mutations, torus maxdepth %d, %d threads, gradient = %s" % (
        chains, steps2, maxdepth, threads, "reference generated code (oracle/_ref)" if ref_grad else "oracle evaluator")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals, sample = [], ""
    for _ in range(args.warmup):
        cpu_chain_rate(threads, 1.0)
    for _ in range(args.steps):
        v, sample = cpu_chain_rate(threads, max(2.0, 40.0 / max(1, args.steps)))
        vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "mutations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "bundled torus scene (scenes/torus), synthetic chain seeds",
            "config": {"workload": "torus LMC maxdepth 8, cache off (BASELINE configs[1] options)", "chains": 256 * threads,
                       "note": "reference's own binary is unbuildable here (Embree/OIIO/Eigen/tup absent); this is the "
                               "oracle port of its chain loop on all host cores"},
            "cpu_baseline": {"value": value, "unit": "mutations/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "mutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains-per-gpu", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--mutations-per-step", type=int, default=MUTATIONS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the device arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lmc = load_package()
    scene = lmc.ParseScene(SCENE_XML)
    scene.options["maxdepth"] = MAXDEPTH
    n_local = args.chains_per_gpu
    total = n_local * world
    M, K, W = args.mutations_per_step, args.steps, args.warmup

    # ---- setup (untimed): MLTInit (init paths generated on rank 0's GPU, lmc_mlt_init_device), broadcast ----
    t_setup = time.time()
    stream = torch.cuda.current_stream()
    ctx = lmc.ChainContext(scene, local, stream=stream.cuda_stream)
    init_t = torch.zeros(total + 1, dtype=torch.float32, device="cuda")
    if rank == 0:
        norm, init_ls = ctx.mlt_init(max(300000, 4 * total), total, 65536)
        init_t[0] = norm
        init_t[1:] = torch.from_numpy(init_ls).cuda()
    if world > 1:
        dist.broadcast(init_t, 0)
    norm = float(init_t[0].item())
    init_ls = init_t[1:].cpu().numpy()
    setup_s = time.time() - t_setup

    film_t = torch.zeros(scene.height, scene.width, 3, dtype=torch.float32, device="cuda")
    ctx.film_bind(film_t.data_ptr())
    total_mut_per_chain = M * (K + W)
    ctx.begin(n_local, norm, init_ls, chain_base=rank * n_local, total_chains=total, samples_per_chain=total_mut_per_chain)
    for _ in range(W):
        ctx.run(M)
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    for _ in range(K):
        ctx.run(M)
    if world > 1:
        dist.all_reduce(film_t)           # the single NCCL all-reduce of the fp32 film
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    st = ctx.stats()
    kernel_ms.append(st["last_kernel_ms"])
    launches = st["kernel_launches"] - launches0 - 1   # minus the stats kernel itself
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mutations = float(total) * M * K
    value = mutations / (ms * 1e-3)

    # ---- e2e: the public API with host buffers: begin (H2D init scores) -> run -> film D2H ----
    e2e_steps = max(1, min(K, 2))
    pinned_init = torch.from_numpy(init_ls).pin_memory()
    host_film = torch.empty(scene.height, scene.width, 3, dtype=torch.float32).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(e2e_steps):
        ta = time.time()
        ctx.begin(n_local, norm, pinned_init.numpy(), chain_base=rank * n_local, total_chains=total, samples_per_chain=M)
        tb = time.time()
        ctx.run(M)
        if world > 1:
            dist.all_reduce(film_t)
        tc = time.time()
        host_film.copy_(film_t, non_blocking=False)
        if os.environ.get("LMC_BENCH_VERBOSE"):
            print("e2e step: begin call %.1f ms, run call %.1f ms, film copy (incl. wait) %.1f ms" % (
                (tb - ta) * 1e3, (tc - tb) * 1e3, (time.time() - tc) * 1e3), file=sys.stderr)
    torch.cuda.synchronize()
    e2e_s = time.time() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = float(total) * M * e2e_steps / e2e_s
    film_bytes = scene.height * scene.width * 3 * 4

    if rank == 0:
        peak, peak_src = measured_peak()
        per_gpu_rate = value / world
        # dominant kernel = k_chain_run: algorithmic bytes per launch / its mean launch time
        k_ms = kernel_ms[-1] if kernel_ms and kernel_ms[-1] > 0 else ms / K
        achieved = (float(n_local) * M * ALGO_BYTES_PER_MUTATION) / (k_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "mutations/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "bundled torus scene (scenes/torus), chains seeded by MLTInit, PCG seeds = chain ids",
                "config": {"workload": "torus, LMC (mala), maxdepth 8, 2^20 chains per GPU, global cache off",
                           "chains_per_gpu": n_local, "mutations_per_step": M, "l2": "inputs larger than L2 (6 GB of chain records + wavefront queues)",
                           "setup_s": round(setup_s, 2), "film_allreduce": world > 1},
                "e2e": {"value": e2e_value, "unit": "mutations/s", "h2d_bytes_per_step": int(4 * total),
                        "d2h_bytes_per_step": int(film_bytes), "steps": e2e_steps},
                "gpu_launches": int(launches),
                "clocks": sampler.result(),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": DRAM_BYTES_PER_MUTATION_NCU * float(n_local) * M,
                             "kernel": "one lmc_run_chains call = M chain-loop iterations of 48 launches each (k_wave_grad, k_trace, "
                                       "k_shade<P_CAM/G_CAM>, k_wave_finish+begin ...; shares in profiles/r01_launches_pervertex_2p20_summary.txt)",
                             "algorithmic_bytes_per_mutation": ALGO_BYTES_PER_MUTATION,
                             "kernel_ms_per_launch": k_ms, "peak_source": peak_src,
                             "per_gpu_mutations_per_s": per_gpu_rate},
                "stats": {"accepted": st["accepted"], "proposed": st["proposed"], "gradient_evals": st["gradient_evals"],
                          "gradient_nonfinite": st["gradient_nonfinite"]}}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sample = cpu_chain_rate(threads, 12.0)
            line["cpu_baseline"] = {"value": v, "unit": "mutations/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

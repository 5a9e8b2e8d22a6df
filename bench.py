#!/usr/bin/env python3
"""bench.py -- MCMC mutations/sec of the LMC chain loop on the bundled scenes.

Headline workload (BASELINE.json configs[1]): torus scene, LMC (mala) mutation, path length (maxdepth) 8,
2^20 Markov chains per GPU, global cache off (every eligible MALA step evaluates its PSS gradient -- the
reverse sweep in the reference's merge order, option adjointcompat = 1), options of scenes/torus/lmc.xml.
One "step" = every chain of the job advanced by MUTATIONS_PER_STEP iterations of the loop at
src/mlt.cpp:91-170 (one `lmc_run_chains` call).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one rank per GPU)
    python bench.py --impl reference ...                      CPU arm: the reference's chain loop as restated by
                                                              the oracle (TIMING build: -Ofast -march=native + libm,
                                                              gradients from the reference's own generated code
                                                              when oracle/_ref exists), all host cores

Timed region (device arm): K steps bracketed by barrier + cuda synchronize, CUDA events on the launching
stream, max over ranks; it includes the final NCCL all-reduce of the fp32 film (src/mlt.cpp:57->200 is the
span the metric is defined on; MLTInit, scene load, BVH build are setup).  Chain records + wavefront queues
(several GB at 2^20 chains) are far larger than L2, so no explicit flush.

The same JSON line carries, under "extra" -> "configs", short device-timed runs of the other BASELINE
configurations (veach-door LMC maxdepth 12; torus H2MC maxdepth 8; at N = 8 the door run IS configs[4]:
2^23 chains over 8 GPUs), each with its own roofline fraction.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "langevin-mcmc_b200")
METRIC = "MCMC mutations/sec (torus, path len 8)"
CHAINS_PER_GPU = 1 << 20
MUTATIONS_PER_STEP = 32


def algo_bytes(kind, L):
    """canonical state words S_LMC(L) = 14 L + 21, S_H2MC(L) = 8 L^2 + 6 L + 19;
    bytes per mutation = 8 S + 48 (state read + written once, two film splats)."""
    s = (14 * L + 21) if kind == "lmc" else (8 * L * L + 6 * L + 19)
    return 8 * s + 48


# name -> scene xml, option overrides, path length L, mutation kind
WORKLOADS = {
    "torus_lmc_L8": dict(xml=os.path.join("torus", "lmc.xml"), opts={"maxdepth": 8}, L=8, kind="lmc",
                         text="torus, LMC (mala), maxdepth 8, 2^20 chains per GPU, global cache off (BASELINE configs[1])"),
    "door_lmc_L12": dict(xml=os.path.join("veachdoor", "lmc.xml"), opts={"maxdepth": 12}, L=12, kind="lmc",
                         text="veach-door, LMC, maxdepth 12, 2^20 chains per GPU (BASELINE configs[2]; x8 GPUs = configs[4])"),
    "torus_h2mc_L8": dict(xml=os.path.join("torus", "h2mc.xml"), opts={"maxdepth": 8}, L=8, kind="h2mc",
                          text="torus, H2MC (Hessian preconditioner), maxdepth 8, 2^20 chains per GPU (BASELINE configs[3])"),
}
WORKLOADS["torus_lmc_L8_cache"] = dict(xml=os.path.join("torus", "lmc.xml"), opts={"maxdepth": 8, "globalcache": 1}, L=8, kind="lmc",
                                       text="torus, LMC, maxdepth 8, 2^20 chains per GPU, global cache ON (the reference's own operating mode: "
                                            "once a PSS dimension holds 3000 entries its chains query the cache instead of evaluating gradients)")
HEADLINE = "torus_lmc_L8"
# dram__bytes_read.sum + dram__bytes_write.sum of all kernels of one steady-state chain-loop iteration over 2^20 chains
# (ncu launch list of this very command, profiles/r02_bench_launches.csv: 13.4 GB per 48-launch iteration), per mutation
DRAM_BYTES_PER_MUTATION_NCU = 12800


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


# ------------------------------------------------------------------------------------------------
# CPU arm (the only place bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------------
def _timing_oracle():
    """oracle/liblmc_oracle_fast.so: the CPU restatement built like the reference (`g++ -Ofast -march=native`,
    src/Tupfile:17) with the platform libm -- the TIMING build.  Rebuilt for this host's ISA when a compiler is
    here (about 40 s); otherwise the prebuilt x86-64-v3 file is used.  Returns (ctypes-backed Oracle, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle
    path = os.path.join(ROOT, "oracle", "liblmc_oracle_fast.so")
    native = os.path.join(ROOT, "oracle", "liblmc_oracle_fast_native.so")
    desc = "-Ofast -march=x86-64-v3 (prebuilt)"
    if not os.path.exists(native) and os.environ.get("LMC_BENCH_NO_REBUILD") is None:
        try:
            subprocess.run(["g++", "-Ofast", "-march=native", "-std=c++17", "-fPIC", "-pthread", "-DLMC_TIMING_LIBM", "-w",
                            "-shared", "-o", native, os.path.join(ROOT, "oracle", "oracle_api.cpp"),
                            os.path.join(PKG_DIR, "csrc", "host", "host_scene.cpp"), "-lz", "-ldl"],
                           check=True, timeout=300, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            pass
    if os.path.exists(native):
        path, desc = native, "-Ofast -march=native (built on this host)"
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ROOT, "oracle_fast"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    o = Oracle.__new__(Oracle)
    o.L = ctypes.CDLL(path)
    o.L.lmco_scene_load.restype = ctypes.c_void_p
    o.L.lmco_last_error.restype = ctypes.c_char_p
    return o, desc


def cpu_chain_rate(o, threads, budget_s, workload=HEADLINE):
    """Times the CPU oracle's chain loop (restatement of src/mlt.cpp:60-196, same options, cache off) on
    `threads` host threads for roughly budget_s seconds.  Returns (mut/s, sample text)."""
    wl = WORKLOADS[workload]
    h = o.load(os.path.join(ROOT, "scenes", wl["xml"]))
    for k, v in wl["opts"].items():
        o.set_option(h, k, v)
    # when the reference's generated gradient code was built (oracle/_ref), the CPU arm evaluates
    # gradients with it (reverse mode, as the reference does) instead of the twin's evaluator
    ref_grad = wl["kind"] == "lmc" and o.use_reference_gradient(True) > 0
    chains, steps = 256 * threads, 16
    norm, init_ls = o.mlt_init(h, 300000, chains, 32)
    t0 = time.time()
    o.run_chains(h, chains, steps, norm, init_ls, threads=threads, want_trace=False, samples_per_chain=steps)
    rate = chains * steps / (time.time() - t0)
    steps2 = int(max(16, min(4096, budget_s * rate / chains)))      # size the measured run from the probe
    t0 = time.time()
    o.run_chains(h, chains, steps2, norm, init_ls, threads=threads, want_trace=False, samples_per_chain=steps2)
    dt = time.time() - t0
    return chains * steps2 / dt, "%d chains x %d mutations, %s, %d threads, gradient = %s" % (
        chains, steps2, workload, threads, "reference generated code (oracle/_ref, ispc -O3 fast-math)" if ref_grad
        else "oracle evaluator")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    o, build = _timing_oracle()
    vals, sample = [], ""
    for _ in range(args.warmup):
        cpu_chain_rate(o, threads, 1.0)
    for _ in range(args.steps):
        v, sample = cpu_chain_rate(o, threads, max(2.0, 40.0 / max(1, args.steps)))
        vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "mutations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "bundled torus scene (scenes/torus), chains seeded by MLTInit (300000 init paths), PCG seeds = chain ids",
            "config": {"workload": WORKLOADS[HEADLINE]["text"], "chains": 256 * threads,
                       "note": "the reference's own binary is unbuildable here (Embree/OIIO/Eigen/tup absent); this is the "
                               "oracle port of its chain loop, timing build " + build + ", on all host cores; the "
                               "reference's published torus LMC run is 4.31 M mutations/s on 32 cores (135 k/s/core, cache on)"},
            "cpu_baseline": {"value": value, "unit": "mutations/s", "cores": threads, "kind": "port", "sample": sample,
                             "build": build},
            "e2e": {"value": value, "unit": "mutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# device arm
# ------------------------------------------------------------------------------------------------
def device_run(lmc, torch, dist, workload, n_local, M, K, W, rank, world, local, want_e2e):
    """W warm-up + K timed steps of `workload`; returns a dict of measurements (rank-local except ms = max)."""
    wl = WORKLOADS[workload]
    scene = lmc.ParseScene(os.path.join(ROOT, "scenes", wl["xml"]))
    for k, v in wl["opts"].items():
        scene.options[k] = v
    total = n_local * world
    stream = torch.cuda.current_stream()
    # ---- setup (untimed): MLTInit with the init paths generated on rank 0's GPU (lmc_mlt_init_device), broadcast.
    # (The sharded form, lmc_mlt_init_device_part on every rank + all-gather + lmc_mlt_init_finish, gives the same bits but
    # is slower here -- 2.0 s against 0.7 s at N = 2: the cost is moving and scanning ~2.4 lsScores per init sample on the
    # host, not generating the paths.) ----
    t_setup = time.time()
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
    if world > 1:
        # the film communicator lives behind the C ABI (lmc_comm_unique_id / lmc_comm_init_rank / lmc_allreduce_film):
        # rank 0 creates the NCCL id, torch.distributed is only the side channel that hands it to the other ranks
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.from_numpy(lmc.comm_unique_id()))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, uid.cpu().numpy())
        ctx.allreduce_film()              # untimed: the first collective of a communicator sets up its channels (film is still zero)
        torch.cuda.synchronize()
    ctx.begin(n_local, norm, init_ls, chain_base=rank * n_local, total_chains=total, samples_per_chain=M * (K + W))
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
    ev0.record(stream)
    for _ in range(K):
        ctx.run(M)
    if world > 1:
        ctx.allreduce_film()              # the single NCCL all-reduce of the fp32 film (lmc_allreduce_film)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    st = ctx.stats()
    launches = st["kernel_launches"] - launches0 - 1   # minus the stats kernel itself
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    res = dict(ms=ms, value=float(total) * M * K / (ms * 1e-3), launches=int(launches), kernel_ms=st["last_kernel_ms"],
               clocks=sampler.result(), setup_s=setup_s, stats=st, scene=scene)

    if want_e2e:
        # ---- e2e through the public API with HOST buffers: lmc_chains_begin uploads this job's init scores from
        # pinned memory (H2D), then K steps of lmc_run_chains, each followed by a read of the step's result -- the
        # progressive film, like the reference's reportIntervalSpp dump (src/mlt.cpp:171-193) -- into pinned host
        # memory (D2H).  Chains continue across the K steps (same large-step schedule as the device-timed region).
        pinned_init = torch.from_numpy(init_ls).pin_memory()
        host_film = torch.empty(scene.height, scene.width, 3, dtype=torch.float32).pin_memory()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        ctx.begin(n_local, norm, pinned_init.numpy(), chain_base=rank * n_local, total_chains=total, samples_per_chain=M * K)
        for k in range(K):
            ctx.run(M)
            if world > 1 and k == K - 1:
                ctx.allreduce_film()      # once, at the end: an all-reduce is in place, the chains keep splatting
            host_film.copy_(film_t, non_blocking=False)
        torch.cuda.synchronize()
        e2e_s = time.time() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        res["e2e"] = {"value": float(total) * M * K / e2e_s, "unit": "mutations/s",
                      "h2d_bytes_per_step": int(4 * total / K), "d2h_bytes_per_step": int(scene.height * scene.width * 3 * 4),
                      "steps": K, "note": "begin (H2D init scores, once) + K x (run + film D2H); chains continue across steps"}
    ctx.close()
    del film_t
    torch.cuda.empty_cache()
    return res


def roofline(res, workload, n_local, M, K, peak, peak_src, world):
    wl = WORKLOADS[workload]
    b = algo_bytes(wl["kind"], wl["L"])
    k_ms = res["kernel_ms"] if res["kernel_ms"] and res["kernel_ms"] > 0 else res["ms"] / K
    achieved = (float(n_local) * M * b) / (k_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (DRAM_BYTES_PER_MUTATION_NCU * float(n_local) * M) if (DRAM_BYTES_PER_MUTATION_NCU and workload == HEADLINE) else None,
            "kernel": "one lmc_run_chains call = M chain-loop iterations (k_wave_grad, k_trace, k_shade<...>, "
                      "k_wave_finish+begin ...; per-kernel shares in profiles/)",
            "algorithmic_bytes_per_mutation": b, "kernel_ms_per_launch": k_ms, "peak_source": peak_src,
            "per_gpu_mutations_per_s": res["value"] / world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains-per-gpu", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--mutations-per-step", type=int, default=MUTATIONS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
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
    n_local, M, K, W = args.chains_per_gpu, args.mutations_per_step, args.steps, args.warmup
    peak, peak_src = measured_peak()

    res = device_run(lmc, torch, dist, HEADLINE, n_local, M, K, W, rank, world, local, want_e2e=True)
    extra = {}
    if not args.no_extra_configs:
        for name in ("door_lmc_L12", "torus_h2mc_L8", "torus_lmc_L8_cache"):
            m2 = max(4, M // 2)
            r2 = device_run(lmc, torch, dist, name, n_local, m2, 2, 1, rank, world, local, want_e2e=False)
            extra[name] = {"workload": WORKLOADS[name]["text"], "value": r2["value"], "unit": "mutations/s", "n_gpus": world,
                           "steps": 2, "warmup": 1, "mutations_per_step": m2, "ms_per_step": r2["ms"] / 2,
                           "gpu_launches": r2["launches"], "roofline": roofline(r2, name, n_local, m2, 2, peak, peak_src, world),
                           "accepted": r2["stats"]["accepted"], "proposed": r2["stats"]["proposed"],
                           "gradient_evals": r2["stats"]["gradient_evals"]}
            if WORKLOADS[name]["opts"].get("globalcache"):
                extra[name].update({k: r2["stats"][k] for k in ("cache_queries", "cache_hits", "cache_count")})

        # the operating point of an actual render of this scene (profiles/r01_render_check.txt: 65 536 long chains keep the
        # start-up bias of the always-accepted first large step small): far fewer chains than the SMs can hold threads for
        rp = device_run(lmc, torch, dist, HEADLINE, 65536, 64, 2, 1, rank, world, local, want_e2e=False)
        extra["torus_lmc_L8_render_point"] = {"workload": "torus, LMC, maxdepth 8, 65 536 chains per GPU (chain count of a converged render)",
                                              "value": rp["value"], "unit": "mutations/s", "n_gpus": world, "steps": 2, "warmup": 1,
                                              "mutations_per_step": 64, "ms_per_step": rp["ms"] / 2, "gpu_launches": rp["launches"]}

    if rank == 0:
        st = res["stats"]
        line = {"metric": METRIC, "value": res["value"], "unit": "mutations/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": res["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "bundled torus scene (scenes/torus), chains seeded by MLTInit, PCG seeds = chain ids",
                "config": {"workload": WORKLOADS[HEADLINE]["text"], "chains_per_gpu": n_local, "mutations_per_step": M,
                           "gradient": "reverse sweep in the reference's merge order (adjointcompat = 1)",
                           "l2": "inputs larger than L2 (GBs of chain records + wavefront queues)",
                           "setup_s": round(res["setup_s"], 2), "film_allreduce": world > 1},
                "e2e": res["e2e"], "gpu_launches": res["launches"], "clocks": res["clocks"],
                "roofline": roofline(res, HEADLINE, n_local, M, K, peak, peak_src, world),
                "stats": {"accepted": st["accepted"], "proposed": st["proposed"], "gradient_evals": st["gradient_evals"],
                          "gradient_nonfinite": st["gradient_nonfinite"]},
                "extra": {"configs": extra}}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            o, build = _timing_oracle()
            v, sample = cpu_chain_rate(o, threads, 12.0)
            line["cpu_baseline"] = {"value": v, "unit": "mutations/s", "cores": threads, "kind": "port", "sample": sample,
                                    "build": build}
            try:       # the parity build (-O2, no fast-math, double-precision deterministic math) for comparison
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from conftest import Oracle
                vp, _ = cpu_chain_rate(Oracle(), threads, 4.0)
                line["cpu_baseline"]["parity_build_value"] = vp
            except Exception:
                pass
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

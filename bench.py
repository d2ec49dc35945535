#!/usr/bin/env python3
"""bench.py -- throughput of the descriptor search, in strand-nucleotides
scanned per second (BASELINE.json: "Gnt/s scanned (both strands) per
descriptor at 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gpumotif|reference]
                    [--descr trna] [--mnt 1024]

A step is one pass of the hot path over one batch: every start offset on both
strands of a synthetic database of `--mnt` Mnt per GPU (i.i.d. uniform acgt,
1 Mnt records -- SURVEY.md 8d "syn_1G"), descriptor test/trna.descr (the one
BASELINE.json's target is quoted on).  One process per GPU; ranks own disjoint
databases (weak scaling); no collective on the data path.

value  kernel + hit gather + host sort with the packed database RESIDENT in HBM
e2e    the same call made with HOST buffers (pinned characters): H2D copy,
       device pack, search, candidates back on the host -- every step
Both are timed with CUDA events on the library's stream, max over ranks.

--impl reference times the reference's own CPU implementation (the binary built
from its sources, oracle/_ref/rnamotif, else the oracle port) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "strand-nt scanned per second (both strands)"
UNIT = "G strand-nt/s"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def load_plan(name):
    import helpers
    return helpers.load_plan(name)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------ CPU arm

def _write_sample(path, n_records, rec_nt, seed):
    from rnamotif_b200 import synth
    ids, seq, off = synth.random_records(seed, [rec_nt] * n_records)
    synth.write_fastn(path, ids, seq, off)
    return int(off[-1])


def cpu_reference_run(descr, procs, rec_per_proc, rec_nt, seed=1001):
    """Time the reference CPU implementation over `procs` files of
    rec_per_proc x rec_nt nucleotides, one process per file (what mrnamotif
    does over MPI, src/mrnamotif.c:884-921).  Returns (strand_nt, seconds, kind)."""
    ref_bin = os.path.join(REF_DIR, "rnamotif")
    descr_file = os.path.join(REF_DIR, "data", "test", descr + ".descr")
    plan = load_plan(descr)
    strands = 2 if int(np.frombuffer(plan, dtype=np.int32, count=9)[8]) else 1
    with tempfile.TemporaryDirectory() as tmp:
        files, total = [], 0
        for p in range(procs):
            f = os.path.join(tmp, "part%03d.fastn" % p)
            total += _write_sample(f, rec_per_proc, rec_nt, seed + p)
            files.append(f)
        if os.path.exists(ref_bin) and os.path.exists(descr_file):
            env = dict(os.environ, EFNDATA=os.path.join(REF_DIR, "data", "efndata"))
            t0 = time.perf_counter()
            ps = [subprocess.Popen([ref_bin, "-descr", descr + ".descr", f],
                                   cwd=os.path.dirname(descr_file), env=env,
                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for f in files]
            rcs = [p.wait() for p in ps]
            dt = time.perf_counter() - t0
            if any(rcs):
                raise RuntimeError("reference rnamotif failed: %r" % rcs)
            return total * strands, dt, "reference"
        # the reference binary did not travel: time the plain-C port instead
        from rnamotif_b200 import fastn, oracle_port
        dbs = [fastn.read_fastn(f) for f in files]

        def work(db):
            oracle_port.scan_db(plan, db[2], db[3], strands == 2)

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(db,)) for db in dbs]  # ctypes releases the GIL
        [t.start() for t in th]
        [t.join() for t in th]
        dt = time.perf_counter() - t0
        return total * strands, dt, "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, cores)
    # ~1.3 M strand-nt/s/core for trna: 4 x 1 Mnt per process ~ 6 s per step
    rec_per_proc, rec_nt = 4, 1_000_000
    times, work, kind = [], 0, "reference"
    for i in range(args.warmup + args.steps):
        w, dt, kind = cpu_reference_run(args.descr, procs, rec_per_proc, rec_nt, seed=1001 + 97 * i)
        if i >= args.warmup:
            times.append(dt)
            work += w
    total_t = sum(times)
    val = work / total_t / 1e9
    sample = f"{procs} processes x {rec_per_proc} x {rec_nt} nt synthetic acgt per step, default flags (-O2.5 prefilter on)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"test/{args.descr}.descr over synthetic uniform acgt, 1 Mnt records, both strands",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm

class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # clocks are evidence, not a dependency
            self.reasons.add("unavailable:" + type(e).__name__)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from rnamotif_b200 import gpumotif

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libgpumotif has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    plan = load_plan(args.descr)
    strands = 2 if gpumotif.plan_field(plan, 8) else 1
    rec_nt = 1_000_000
    n_rec = max(1, args.mnt)
    total = n_rec * rec_nt
    rec_off = np.arange(n_rec + 1, dtype=np.int64) * rec_nt

    # synthetic database of this rank: uniform acgt characters, generated on the
    # device (seeded per rank), copied once to pinned host memory for the e2e leg
    g = torch.Generator(device="cuda")
    g.manual_seed(1001 + rank)
    lut = torch.tensor(list(b"acgt"), dtype=torch.uint8, device="cuda")
    d_chars = torch.empty(total, dtype=torch.uint8, device="cuda")
    for o in range(0, total, 1 << 27):
        n = min(1 << 27, total - o)
        d_chars[o:o + n] = lut[torch.randint(0, 4, (n,), generator=g, device="cuda")]
    h_chars = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    h_chars.copy_(d_chars)
    torch.cuda.synchronize()

    ms = gpumotif.MotifSearch(plan, device=local)
    if args.tile:
        ms.set_tile(args.tile)
    stream = torch.cuda.ExternalStream(ms.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    # ---- resident leg -------------------------------------------------------
    ms.set_device_chars(d_chars.data_ptr(), rec_off)
    kernel_ms, filter_ms, launches, hits_n, filter_launches, survivors = [], [], 0, 0, 0, 0

    def step_resident():
        nonlocal launches, hits_n, filter_launches, survivors
        ms.scan(0, total, strands, copy=False)
        st = ms.stats()
        kernel_ms.append(st.kernel_ms)
        filter_ms.append(st.filter_ms)
        launches += st.n_launches
        filter_launches += st.n_filter_launches
        survivors = st.n_survivors
        hits_n = st.n_hits

    for _ in range(args.warmup):
        step_resident()
    kernel_ms.clear()
    filter_ms.clear()
    launches = filter_launches = 0
    sampler = ClockSampler(local)
    sampler.start()
    t_res = timed(step_resident, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    k_ms = float(np.mean(kernel_ms))
    gpu_launches = launches

    # ---- end-to-end leg: host characters in, candidates out, every step -------
    h2d = d2h = 0

    phases = {}

    def step_e2e():
        nonlocal h2d, d2h
        ms.upload_ptr(h_chars.data_ptr(), rec_off)
        ms.scan(0, total, strands, copy=False)  # candidates are in host memory (library buffer)
        st = ms.stats()
        h2d, d2h = st.h2d_bytes, st.d2h_bytes
        phases.update(h2d_ms=st.h2d_ms, pack_ms=st.pack_ms, kernel_ms=st.kernel_ms, d2h_ms=st.d2h_ms,
                      sort_ms=st.sort_ms)

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    t_e2e = timed(step_e2e, args.steps)

    work = float(total) * strands * world  # strand-nt per step, all ranks
    value = work * args.steps / (t_res / 1e3) / 1e9
    e2e = work * args.steps / (t_e2e / 1e3) / 1e9

    if rank == 0:
        pk, which = peaks()
        # The dominant kernel.  On the worklist path (strong level-0 filter, e.g. trna)
        # it is the sieve kernel gm_search_kernel<1,*,2>: it alone reads the packed
        # database (0.5 B per nt, once for both strands) and writes the worklist; the
        # enumeration kernel gm_dfs_kernel touches only the survivors' windows.  On the
        # fused path one kernel does both.  Its average launch duration comes from CUDA
        # events the library records around every launch on its stream.
        split = filter_launches > 0
        if split and float(np.sum(filter_ms)) < 0.5 * float(np.sum(kernel_ms)):
            # worklist path whose enumeration kernel dominates (weak filters: pk1, trna.general)
            dom_name = "gm_dfs_kernel<FULL> (enumeration of the filter's survivors)"
            dom_ms = (float(np.sum(kernel_ms)) - float(np.sum(filter_ms))) / filter_launches
            per_launch_nt = total * args.steps / filter_launches
            alg_bytes = survivors / max(filter_launches / args.steps, 1) * (32 + 64) + \
                hits_n * (32 + 8 * ms.n_descr) / max(filter_launches / args.steps, 1)
            split = False
            filter_share = (float(np.sum(kernel_ms)) - float(np.sum(filter_ms))) / t_res
        elif split:
            filter_share = float(np.sum(filter_ms)) / t_res
            dom_name = "gm_search_kernel<1,FULL,PF> (level-0 sieve / prefilter)"
            dom_ms = float(np.sum(filter_ms)) / filter_launches            # per launch
            per_launch_nt = total * args.steps / filter_launches
            alg_bytes = per_launch_nt * 0.5 + survivors / max(filter_launches / args.steps, 1) * 32
        else:
            filter_share = float(np.sum(kernel_ms)) / t_res
            dom_name = "gm_search_kernel<0,FULL,PF> (fused filter + enumeration)"
            dom_ms = k_ms / max(gpu_launches / args.steps, 1)
            per_launch_nt = total * args.steps / max(gpu_launches, 1)
            alg_bytes = per_launch_nt * 0.5 + hits_n * (32 + 8 * ms.n_descr) / max(gpu_launches / args.steps, 1)
        achieved = alg_bytes / (dom_ms / 1e3) / 1e9
        traffic = None
        tj = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tj):
            t = json.load(open(tj))
            if t["workload"] == {"descr": args.descr, "mnt": args.mnt}:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]  # bytes per launch, ncu
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "algorithmic_bytes": alg_bytes,
                "peak_source": which,
                "kernel": dom_name, "kernel_ms": dom_ms, "kernel_share_of_step": filter_share,
                "all_kernels_ms_per_step": k_ms,
                "note": "the search is integer-issue bound, not HBM bound (SURVEY.md F8); see issue and profiles/"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"test/{args.descr}.descr over syn_{args.mnt}M: {n_rec} x 1 Mnt uniform acgt per GPU, "
                                   f"{strands} strand(s)", "l2": "input (packed) larger than L2" if total / 2 > 126e6
                       else "input smaller than L2; each step re-reads it after the hit gather",
                       "candidates_per_step_rank0": int(hits_n)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": t_e2e / args.steps, "phases_ms_last_step": phases},
            "gpu_launches": int(gpu_launches),
            "roofline": roof,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu:
            w, dt, kind = cpu_reference_run(args.descr, 1, 16, 1_000_000)
            line["cpu_baseline"] = {"value": w / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "16 x 1 Mnt synthetic acgt, one process, default flags"}
            # the reference's pair-rule evaluations per strand-nt on this input class (oracle count)
            try:
                from rnamotif_b200 import oracle_port, synth
                ids, seq, off = synth.random_records(5, [200_000])
                _, st = oracle_port.scan_db(plan, seq, off, strands == 2)
                W = st.n_pair_evals / max(st.n_starts, 1)
                sm_clk = (line["clocks"]["sm_mhz"] or 1965.0) * 1e6
                lane_peak = 148 * 128 * sm_clk
                line["issue"] = {"pair_evals_per_strand_nt": W,
                                 "pair_evals_per_s": W * total * strands / (k_ms / 1e3),
                                 "int_lane_peak_per_s": lane_peak,
                                 "frac_of_lane_peak": W * total * strands / (k_ms / 1e3) / lane_peak,
                                 "note": "the reference's pair-rule evaluations per strand-nt (oracle count) x measured "
                                         "strand-nt/s; the sieve does the same tests 32 starts to a word, so this is "
                                         "the algorithm-level rate, not an instruction count"}
            except Exception as e:
                line["issue"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    ms.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpumotif", choices=["gpumotif", "reference"])
    ap.add_argument("--descr", default="trna")
    ap.add_argument("--mnt", type=int, default=1024, help="Mnt of synthetic sequence per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tile", type=int, default=0, help="starts per tile (0 = library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

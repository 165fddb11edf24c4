#!/usr/bin/env python
"""bench.py -- measures the association hot path on B200(s) and prints ONE JSON line (rank 0).

Headline (`value`): marginal-scan throughput in GB/s of 2-bit genotype data (algorithmic bytes =
n_snps * n_samples / 4) on BASELINE.json configs[1] -- 5 000 cases / 5 000 controls x 500 000 SNPs,
allelic + genotypic chi-square -- with the store resident in HBM; roofline = HBM bandwidth.
`e2e`: the reference's `select_cc_maf` call surface (algorithms/maf_func.cpp:238-269) end to end through
the C-ABI with HOST buffers: case/control masks from host memory -> compaction (K0) -> scan (K1) ->
counts + statistics copied back to pinned host memory, every step.
`pairwise`: the exhaustive SNP x SNP screen (configs[2], 2 000/2 000 x 50 000 SNPs = 1 249 975 000 pairs)
in pairs/s. Its roofline fraction stays defined on the integer-popcount formula (4 AND+POPC
word-cells per 32 samples per pair) whichever engine runs; since the tensor-core engine (tcgen05 kind::i8) does the
same counting as a GEMM, that fraction exceeds 1 and a second object, `roofline.tensor`, gives the int8 MAC rate
against the tensor-core peak. Plus its own e2e (computeBoost call surface) and CPU baseline.

N > 1 (torchrun): weak scaling for the marginal scan (each rank scans its own 500 000-SNP shard, no
collective on the data path); the pairwise screen shards tile pairs in chunks over the ranks and
all-gathers the hit lists over NCCL inside the timed region.

--impl reference: the reference's own CPU implementation (oracle/_ref, the unmodified sources) on all
host cores, same metric, on a bounded sample of the same workload.
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
sys.path.insert(0, ROOT)

SEED = 20121127
MARGINAL = dict(n_snps=500_000, n_samples=10_000, n_case=5_000)       # BASELINE.json configs[1]
PAIRWISE = dict(n_snps=50_000, n_samples=4_000, n_case=2_000)         # BASELINE.json configs[2]
PAIRWISE_CFG3 = dict(n_snps=500_000, n_samples=10_000, n_case=5_000)  # BASELINE.json configs[3] (multi-GPU runs)
NCU_TRAFFIC_MARGINAL = 1.2930e9       # dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r1k_marginal_scan_full.md)
NCU_TRAFFIC_PAIRWISE = 1.9517e10      # same for pair_screen_mma_kernel at configs[2] on one GPU (profiles/r1h_pair_screen_mma_full.md)
CPU_MARGINAL_SAMPLE_SNPS = 2_000
CPU_PAIRWISE_SAMPLE_SNPS = 1_500


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    The scan's timed region is milliseconds long, far below what `nvidia-smi -lms` can resolve, so the same NVML
    counters nvidia-smi prints are polled in-process (pynvml, ~0.5 ms period); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.idx, self.uuid, self.rows, self.proc, self.nvml, self.source = gpu_index, uuid, [], None, None, None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                h = pynvml.nvmlDeviceGetHandleByIndex(int(ids[self.idx]) if self.idx < len(ids) else self.idx)
            self.nvml, self.handle, self.source = pynvml, h, "nvml"
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(reasons_fn(h))
                watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((time.perf_counter(), [str(self.idx), sm, self.max_sm, watts, mask]))
            except Exception:
                pass
            time.sleep(0.0005)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                mask = sum(bit for (bit, _), v in zip(self.REASONS, f[4:8]) if v.lower().startswith("active"))
                self.rows.append((time.perf_counter(), [f[0], float(f[1]), float(f[2]), float(f[3]), mask]))
            except (ValueError, IndexError):
                pass

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        where = "timed region"
        if not rows:   # region shorter than one sampling period: the nearest samples around it
            rows = [r for _, r in sorted(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))[:3]]
            where = "nearest samples to the timed region"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[1] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[4]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][2], "reasons": sorted(n for b, n in self.REASONS if mask & b),
                "samples": len(rows), "power_w_max": round(max(r[3] for r in rows), 1), "source": self.source, "window": where}

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()


def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run this rank's host thread on the CPUs next to its GPU (NVML's ideal affinity), so that the
    pinned host buffers of the e2e path land on that NUMA node instead of wherever torchrun started the process."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import libgwaspp_b200 as gw

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)     # pinned result buffers are first-touched on the GPU's own node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks, peak_src = measured_peaks()
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    K, W = args.steps, args.warmup
    out = {}

    # ------------------------------------------------------------------ marginal scan (headline)
    M, N, NCASE = args.snps or MARGINAL["n_snps"], args.samples or MARGINAL["n_samples"], args.cases or MARGINAL["n_case"]
    st = gw.GenoStore(M, N, device=local)
    st.set_stream(stream.cuda_stream)
    st.simulate(SEED + rank)                                   # each rank owns a different 500k-SNP shard
    pheno = gw.simulate_phenotype(SEED, N, NCASE)
    case_mask, ctrl_mask = gw.stream_masks(pheno)
    st.select_case_control(case_mask=case_mask, ctrl_mask=ctrl_mask)
    d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
    d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
    bytes_per_step = M * N / 4.0                               # algorithmic bytes (2 bits per genotype)

    for _ in range(W):
        st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0, t0 = gw.launch_count(), time.perf_counter()
    ev0.record(stream)
    for _ in range(K):
        st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
    ev1.record(stream)
    barrier()
    t1, launches = time.perf_counter(), gw.launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / K
    value = world * bytes_per_step / (ms_per_step * 1e-3) / 1e9
    clocks = sampler.window(t0, t1) if rank == 0 else None

    # per-launch duration of the dominant kernel, CUDA events on the launching stream (inside the library)
    kms = []
    for _ in range(min(K, 10)):
        st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
        kms.append(st.last_scan_ms())
    k_ms = float(np.mean(kms))
    achieved = bytes_per_step / (k_ms * 1e-3) / 1e9
    traffic = args.traffic_bytes
    if traffic is None and (M, N, NCASE) == (MARGINAL["n_snps"], MARGINAL["n_samples"], MARGINAL["n_case"]):
        traffic = NCU_TRAFFIC_MARGINAL          # ncu --set full capture of this launch shape, profiles/r1k_marginal_scan_full.md
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic,
                "kernel": "marginal_scan_kernel", "kernel_ms": round(k_ms, 4), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_step}

    # e2e: select_cc_maf call surface with host buffers
    h_counts = torch.empty((M, 8), dtype=torch.int32, pin_memory=True)
    h_stats = torch.empty((M, 8), dtype=torch.float64, pin_memory=True)
    e2e_steps = max(3, min(K, 10))
    for it in range(2 + e2e_steps):
        if it == 2:
            barrier()
            te0 = time.perf_counter()
        st.select_case_control(case_mask=case_mask, ctrl_mask=ctrl_mask)
        st.marginal_scan_into(0, M, counts=h_counts, stats=h_stats, on_device=False)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - te0) / e2e_steps)
    e2e_kernel_ms = st.last_scan_ms()
    wr = (st.P // 2 + 3) // 4 * 4                              # device words per raw plane
    h2d = int(3 * 4 * wr + 2 * 4 * (12 * wr + 1) + 4 * ((NCASE + 31) // 32 + (N - NCASE + 31) // 32))   # 3 masks + compaction tables
    d2h = int(h_counts.numel() * 4 + h_stats.numel() * 8)
    e2e = {"value": round(world * bytes_per_step / e2e_s / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s * 1e3, 3), "steps": e2e_steps,
           "what": "gwasdev_select_case_control(host masks) + gwasdev_marginal_scan(host pinned outputs): the first scan after a "
                   "selection runs marginal_scan_masked_kernel on the raw rows (select fused into the scan, no compaction), "
                   "in pieces whose D2H copies overlap the next piece's scan",
           "kernel": "marginal_scan_masked_kernel", "kernel_span_ms": round(e2e_kernel_ms, 4),
           "host_cpus": (f"{numa[0]}-{numa[-1]} ({len(numa)}, NVML affinity of the GPU)" if numa else "unbound"),
           "bound": "PCIe D2H of the 96 B/SNP results"}
    # sanity: the device and host paths agree, and the scan did real work
    assert torch.equal(h_counts, d_counts.cpu()) and int(h_counts[:, :4].sum(1).min()) == NCASE
    st.close()
    del d_counts, d_stats, h_counts, h_stats
    torch.cuda.empty_cache()

    out.update({
        "metric": "marginal-scan GB/s vs HBM peak", "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bit-planes (popcount) + f64 statistics", "data": "synthetic",
        "config": {"workload": f"configs[1]: {NCASE} cases / {N - NCASE} controls x {M} SNPs per GPU, marginal allelic + "
                               "genotypic chi-square scan (counts + statistics written)",
                   "generator": "simulate_data.cpp restated, affy6 panel of maf_spectrum.tab, seed 20121127",
                   "l2": f"inputs larger than L2 ({bytes_per_step / 1e6:.0f} MB streamed per step vs 126 MB L2)",
                   "parallelism": "one process per GPU, SNP-range shards, no data-path collective"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    })

    # ------------------------------------------------------------------ biobank scale (configs[4]), optional
    if args.biobank and rank == 0:
        out["biobank"] = biobank_gpu(args, gw, torch, local, stream, peaks, peak_src)

    # ------------------------------------------------------------------ pairwise screen
    if not args.no_pairwise:
        out["pairwise"] = pairwise_gpu(args, gw, torch, dist, rank, world, local, stream, barrier, max_over_ranks,
                                       sum_over_ranks, sampler)
        if world == 1 and not args.pw_snps and not args.no_missing:
            out["pairwise_missing_calls"] = pairwise_missing_gpu(gw, local, stream)
        if world > 1 and not args.no_cfg3 and not args.pw_snps:
            # BASELINE.json configs[3]: the north star's target problem, sharded by tile pairs over the ranks
            out["pairwise_configs3"] = pairwise_gpu(args, gw, torch, dist, rank, world, local, stream, barrier, max_over_ranks,
                                                    sum_over_ranks, sampler, shape=PAIRWISE_CFG3)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline_marginal(N, NCASE, threads=1, steps=3)
            out["cpu_baseline"] = cb
            out["configs0_file_to_statistics"] = configs0_file_to_statistics(gw, local)
            if "pairwise" in out:
                out["pairwise"]["cpu_baseline"] = cpu_baseline_pairwise(args.pw_samples or PAIRWISE["n_samples"],
                                                                        args.pw_cases or PAIRWISE["n_case"])
        sampler.stop()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def configs0_file_to_statistics(gw, local):
    """BASELINE configs[0] -- 1 000 cases / 1 000 controls x 10 000 SNPs, the case the reference runs on a CPU -- from the
    genotype FILE to the per-SNP statistics on the host: 80 MB of TPED text -> device loader -> selection -> scan -> host
    results, next to the reference doing the same with its own reader and select_cc_maf on one host core (oracle/_ref)."""
    import tempfile
    import oracle
    M, N, NCASE = 10_000, 2_000, 1_000
    pheno = gw.simulate_phenotype(SEED, N, NCASE)
    with gw.GenoStore(M, N, device=local) as st:            # the cohort: device generator -> rows -> text
        st.simulate(SEED)
        rows = st.get_rows()
        st.select_case_control(pheno)
        want = st.marginal_scan(mi=False)["counts"]
    P = (rows.shape[1] - 1) // 2
    p1 = np.unpackbits(np.ascontiguousarray(rows[:, 1:1 + P]).view(np.uint8), axis=1, bitorder="little")[:, :N]
    p2 = np.unpackbits(np.ascontiguousarray(rows[:, 1 + P:]).view(np.uint8), axis=1, bitorder="little")[:, :N]
    lut = np.array([np.frombuffer(t, np.uint8).view(np.uint32)[0] for t in (b"0\t0\t", b"A\tA\t", b"A\tC\t", b"C\tC\t")], np.uint32)
    code = p1 + 2 * p2
    d = tempfile.mkdtemp()
    tped, tfam = os.path.join(d, "c0.tped"), os.path.join(d, "c0.tfam")
    with open(tped, "wb") as f:
        for r in range(M):
            f.write(b"0\trs%d\t0\t%d\t" % (r, r) + lut[code[r]].view(np.uint8).tobytes()[:-1] + b"\n")
    with open(tfam, "w") as f:
        for i, ph in enumerate(pheno):
            f.write(f"F{i}\tI{i}\t0\t0\t1\t{int(ph)}\n")
    res = {"workload": f"configs[0]: {NCASE} cases / {N - NCASE} controls x {M} SNPs, {os.path.getsize(tped) / 1e6:.1f} MB TPED file (page cache) "
                       "-> per-SNP case/control counts + statistics on the host"}
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        with gw.GenoStore.from_tped(tped, device=local) as st:     # table sized and loaded in one pass over the file
            t2 = time.perf_counter()
            st.select_case_control(pheno)
            got = st.marginal_scan(mi=False)
            t3 = time.perf_counter()
        cur = {"create_load_ms": round(1e3 * (t2 - t0), 2), "select_scan_ms": round(1e3 * (t3 - t2), 2),
               "total_ms": round(1e3 * (t3 - t0), 2)}
        if best is None or cur["total_ms"] < best["total_ms"]:
            best = cur
    assert np.array_equal(got["counts"], want), "counts from the loaded file differ from the generated cohort's"
    res["b200"] = best
    if oracle.have_ref():
        t0 = time.perf_counter()
        R = oracle.Ref(tped=tped, tfam=tfam, level=5)
        t1 = time.perf_counter()
        R.run("select_cc_maf")
        t2 = time.perf_counter()
        ref_counts = np.stack([R.cc_dist(r, 1)[0] for r in range(0, M, 101)])
        assert np.array_equal(ref_counts, want[::101]), "the reference's counts differ"
        res["reference"] = {"load_ms": round(1e3 * (t1 - t0), 1), "select_cc_maf_ms": round(1e3 * (t2 - t1), 1),
                            "total_ms": round(1e3 * (t2 - t0), 1), "cores": 1, "kind": "reference",
                            "what": "TpedGenotypeFile + TfamAnnotationFile readers, compute(select_cc_maf) at --comp-level 5"}
    for fn in (tped, tfam):
        os.remove(fn)
    os.rmdir(d)
    return res


def biobank_gpu(args, gw, torch, local, stream, peaks, peak_src):
    """BASELINE configs[4]: 200 000 samples x 1 000 000 SNPs. (a) HBM-resident: the whole 50 GB scan layout in one
    launch. (b) streamed in sample blocks: raw rows of one block at a time from pinned host memory -> H2D -> compaction
    -> counts added into a device accumulator, statistics from the summed counts at the end; bounded to --bb-stream-snps
    SNPs so that the pinned host copy stays small (the streamed rate is PCIe-bound and scales linearly in SNPs)."""
    N, NCASE, M = 200_000, 100_000, args.bb_snps
    res = {"workload": f"configs[4]: {NCASE} cases / {N - NCASE} controls x {M} SNPs"}
    pheno = gw.simulate_phenotype(SEED, N, NCASE)
    with gw.GenoStore(M, N, device=local) as st:
        st.set_stream(stream.cuda_stream)
        st.simulate(SEED)
        st.select_case_control(pheno)
        d_counts = torch.empty((M, 8), dtype=torch.int32, device="cuda")
        d_stats = torch.empty((M, 8), dtype=torch.float64, device="cuda")
        kms = []
        for it in range(8):
            st.marginal_scan_into(0, M, counts=d_counts, stats=d_stats)
            if it >= 3:
                kms.append(st.last_scan_ms())
        k_ms = float(np.mean(kms))
        gbs = M * N / 4.0 / (k_ms * 1e-3) / 1e9
        res["resident"] = {"value": round(gbs, 1), "unit": "GB/s", "kernel_ms": round(k_ms, 3), "bytes": M * N / 4.0,
                           "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4), "peak_source": peak_src}
        ref_counts = d_counts[: args.bb_stream_snps].cpu().numpy().view(np.uint32)
        del d_counts, d_stats
    torch.cuda.empty_cache()
    # streamed: host copy of the first Ms SNPs, block by block
    Ms, B = args.bb_stream_snps, args.bb_block
    blocks = [(s0, min(B, N - s0)) for s0 in range(0, N, B)]
    host = []
    for s0, nb in blocks:                                   # setup (untimed): generate each block, keep its raw rows pinned
        with gw.GenoStore(Ms, nb, device=local) as st:
            st.simulate_block(SEED, s0, N)
            rows = torch.from_numpy(st.get_rows()).pin_memory()
            host.append(rows)
    acc = torch.zeros((Ms, 8), dtype=torch.int32, device="cuda")
    stores = {nb: gw.GenoStore(Ms, nb, device=local) for nb in {nb for _, nb in blocks}}
    times = []
    for rep in range(3):
        acc.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for (s0, nb), rows in zip(blocks, host):
            st = stores[nb]
            st.put_rows(rows.numpy())
            st.select_case_control(pheno[s0:s0 + nb])
            st.marginal_accumulate(acc, on_device=True)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    for st in stores.values():
        st.close()
    assert np.array_equal(acc.cpu().numpy().view(np.uint32), ref_counts), "streamed counts differ from the resident scan"
    t = min(times)
    h2d = sum(int(r.numel()) * 2 for r in host)
    res["streamed"] = {"value": round(Ms * N / 4.0 / t / 1e9, 2), "unit": "GB/s of 2-bit genotypes, end to end from pinned host memory",
                       "snps": Ms, "sample_blocks": len(blocks), "block_samples": B, "seconds": round(t, 4),
                       "h2d_bytes": h2d, "h2d_gbs": round(h2d / t / 1e9, 2),
                       "check": "summed counts bit-identical to the resident scan"}
    return res


def pairwise_missing_gpu(gw, local, stream):
    """configs[2] with 1 % of the calls missing (what real genotype data looks like): every 64-SNP block has missing calls,
    so the whole screen takes the reference's 9-cell branch -- here the four-plane tensor-core kernel; the 9-cell AND+POPC
    kernel it replaces is timed beside it."""
    M, N, NCASE = PAIRWISE["n_snps"], PAIRWISE["n_samples"], PAIRWISE["n_case"]
    res = {"workload": f"configs[2] shape with 1 % missing calls: {NCASE}/{N - NCASE} samples x {M} SNPs ({M * (M - 1) // 2} pairs), "
                       "9-cell tables (compressed_genotype_table5.cpp:1000-1067) + KSA statistic, threshold 30"}
    with gw.GenoStore(M, N, device=local) as st:
        st.set_stream(stream.cuda_stream)
        st.simulate(SEED, missing_rate=0.01)
        st.select_case_control(gw.simulate_phenotype(SEED, N, NCASE))
        for name, engine in (("tensor_cores_four_planes", 0), ("and_popc_nine_cells", 1)):
            st.set_pair_engine(engine)
            ms, hits = [], None
            for it in range(4):
                hits, s = st.pairwise_scan(30.0)
                if it:
                    ms.append(s.screen_ms)
            k_ms = float(np.mean(ms))
            res[name] = {"value": round(s.pairs_tested / (k_ms * 1e-3), 1), "unit": "pairs/s", "kernel_ms": round(k_ms, 3),
                         "tiles_with_missing_calls": int(s.tiles_nine_cell), "hits": int(len(hits))}
        cells = 9 * ((NCASE + 31) // 32 + (N - NCASE + 31) // 32)
        peak, _ = gw.popc_peak(local)
        res["and_popc_nine_cells"]["frac_of_popc_roofline"] = round(res["and_popc_nine_cells"]["value"] * cells / peak, 4)
        res["tensor_cores_four_planes"]["x_popc_roofline"] = round(res["tensor_cores_four_planes"]["value"] * cells / peak, 4)
    return res


def pairwise_gpu(args, gw, torch, dist, rank, world, local, stream, barrier, max_over_ranks, sum_over_ranks, sampler, shape=None):
    if shape is None:
        M, N, NCASE = args.pw_snps or PAIRWISE["n_snps"], args.pw_samples or PAIRWISE["n_samples"], args.pw_cases or PAIRWISE["n_case"]
        K, W = max(1, min(args.steps, args.pw_steps)), min(args.warmup, 3)
    else:   # configs[3]: every pass is 0.4 s (8 GPUs) to 1.4 s (2 GPUs) of tensor-core work
        M, N, NCASE = shape["n_snps"], shape["n_samples"], shape["n_case"]
        K, W = max(1, min(args.steps, 2)), 1
    st = gw.GenoStore(M, N, device=local)
    st.set_stream(stream.cuda_stream)
    st.simulate(SEED)                                          # the store is replicated on every rank
    pheno = gw.simulate_phenotype(SEED, N, NCASE)
    case_mask, ctrl_mask = gw.stream_masks(pheno)
    st.select_case_control(case_mask=case_mask, ctrl_mask=ctrl_mask)
    cap = 1 << 20
    d_hits = torch.empty((cap, 2), dtype=torch.int64, device="cuda")     # 16-byte gwasdev_hit records
    from libgwaspp_b200 import multi_gpu as mg

    def step():
        n, stats = st.pairwise_scan(30.0, shard=rank, n_shards=world, capacity=cap, hits=d_hits, on_device=True)
        if world > 1:   # top-k / hit gather over NCCL: counts, then the padded hit buffers (one collective each)
            counts = mg.gather_counts(int(n), world, torch.device("cuda", local))
            mx = max(1, max(counts))
            mg.gather_records(d_hits[:mx], world)
            n = int(sum(counts))
        return n, stats

    for _ in range(W):
        step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0, t0 = gw.launch_count(), time.perf_counter()
    ev0.record(stream)
    screen_ms, pairs, cells, n_hits, cand, engine = [], 0, 0, 0, 0, 0
    for _ in range(K):
        n_hits, s = step()
        screen_ms.append(s.screen_ms)
        pairs, cells, cand, engine = s.pairs_tested, s.word_cells, s.candidates, s.engine
    ev1.record(stream)
    barrier()
    t1, launches = time.perf_counter(), gw.launch_count() - l0
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / K
    total_pairs = sum_over_ranks(float(pairs))
    value = total_pairs / (ms_per_step * 1e-3)
    k_ms = float(np.mean(screen_ms))
    peak_cells, clk = gw.popc_peak(local)
    achieved = cells / (k_ms * 1e-3)
    peaks, peak_src = measured_peaks()
    # tensor-core view of the same launch: one int8 MAC per (cell, sample), 4 cells per pair
    macs = float(pairs) * 4.0 * N
    tensor_peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))      # int8 runs at twice the bf16 rate on the same datapath
    tensor = {"bound": "tensor", "achieved": round(2.0 * macs / (k_ms * 1e-3) / 1e12, 1), "peak": round(tensor_peak, 1),
              "unit": "TOP/s (int8, 2 ops per MAC)", "frac": round(2.0 * macs / (k_ms * 1e-3) / 1e12 / tensor_peak, 4),
              "peak_source": f"2 x bf16_tflops of {peak_src} (nominal int8 dense 4 500)",
              "algorithmic_macs_per_launch": int(macs)}
    kernel = {2: "pair_screen_mma_kernel (tcgen05.mma cta_group::2 kind::i8)", 1: "pair_screen_kernel<false> (AND+POPC)"}.get(engine, "?")
    res = {
        "metric": "pairwise SNP x SNP tests/sec", "value": round(value, 1), "unit": "pairs/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": round(ms_per_step, 3), "scaling": "strong", "hits": int(n_hits), "candidates": int(cand),
        "config": {"workload": f"{'configs[3]' if (M, N) == (500_000, 10_000) else 'configs[2]'}: exhaustive pairwise epistasis {NCASE}/{N - NCASE} samples x {M} SNPs "
                               f"({M * (M - 1) // 2} pairs), 3x3x2 contingency + KSA statistic, threshold 30",
                   "engine": {2: "tensor cores", 1: "AND+POPC"}.get(engine, "?"),
                   "parallelism": f"128x128 SNP tile pairs dealt in chunks of 64 over {world} rank(s); NCCL all_gather of hits"},
        "roofline": {"bound": "int_popc", "achieved": round(achieved / 1e12, 4), "peak": round(peak_cells / 1e12, 4),
                     "unit": "T word-cells/s (32-bit AND+POPC)", "frac": round(achieved / peak_cells, 4),
                     "traffic": NCU_TRAFFIC_PAIRWISE if (engine == 2 and world == 1 and (M, N, NCASE) == (PAIRWISE["n_snps"], PAIRWISE["n_samples"], PAIRWISE["n_case"])) else None,
                     "kernel": kernel, "kernel_ms": round(k_ms, 3),
                     "peak_source": f"register-only __popc microbenchmark run in this process (clock attr {clk:.0f} MHz); "
                                    "nominal 148 SMs x 16 POPC/clk x 1.965 GHz = 4.65",
                     "algorithmic_word_cells_per_launch": int(cells),
                     "tensor": tensor if engine == 2 else None},
        "gpu_launches": int(launches),
        "clocks": sampler.window(t0, t1) if rank == 0 else None,
    }
    # e2e: computeBoost call surface: masks from host -> select -> margins -> screen -> G-test -> hits on host
    e2e_steps = 3
    for it in range(1 + e2e_steps):
        if it == 1:
            barrier()
            te0 = time.perf_counter()
        st.select_case_control(case_mask=case_mask, ctrl_mask=ctrl_mask)
        hits, s = st.pairwise_scan(30.0, shard=rank, n_shards=world)
        if len(hits):
            st.gtest(hits["i"], hits["j"])
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - te0) / e2e_steps)
    res["e2e"] = {"value": round(total_pairs / e2e_s, 1), "unit": "pairs/s", "ms_per_step": round(e2e_s * 1e3, 2),
                  "h2d_bytes_per_step": int(64 * ((N // 32 + 4) // 4 * 4) + 8 * len(hits)),
                  "d2h_bytes_per_step": int(32 * max(1, len(hits))), "steps": e2e_steps,
                  "what": "select_case_control + pairwise_scan (margins, screen, fp64 re-score, sort) + gtest, host buffers"}
    st.close()
    return res


# ---------------------------------------------------------------------------------------------------
# CPU baselines (the reference itself when oracle/_ref is present, else the C port)
# ---------------------------------------------------------------------------------------------------
def _ref_marginal_sample(seed, n_snps, N, NCASE):
    import oracle
    O = oracle.Oracle()
    codes, pheno = O.simulate(seed, n_snps, N, NCASE)
    if oracle.have_ref():
        R = oracle.Ref(n_snps, N, 5)
        R.add_codes(codes)
        R.set_case_control(pheno)
        return "reference", R, None
    rows = O.pack_codes(codes)
    return "port", O, (rows, pheno, N)


def _marginal_step(kind, obj, extra):
    """one select_cc_maf-equivalent pass: (t_select, t_scan) seconds."""
    if kind == "reference":
        return obj.time_phase(0, 1), obj.time_phase(2, 1)
    rows, pheno, N = extra
    t0 = time.perf_counter()
    sel, nca, nco = obj.select(rows, N, pheno)
    t1 = time.perf_counter()
    obj.cc_counts_selected(sel, nca, nco)
    return t1 - t0, time.perf_counter() - t1


def cpu_baseline_marginal(N, NCASE, threads=1, steps=3):
    kind, obj, extra = _ref_marginal_sample(SEED, CPU_MARGINAL_SAMPLE_SNPS, N, NCASE)
    _marginal_step(kind, obj, extra)
    ts = [_marginal_step(kind, obj, extra) for _ in range(steps)]
    t_sel, t_scan = float(np.median([t[0] for t in ts])), float(np.median([t[1] for t in ts]))
    b = CPU_MARGINAL_SAMPLE_SNPS * N / 4.0
    return {"value": round(b / t_scan / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": kind,
            "sample": f"{CPU_MARGINAL_SAMPLE_SNPS} SNPs x {N} samples of the same cohort; value = pre-selected per-SNP counts + "
                      "MinorAlleleFrequency loop (select_cc_maf body without its per-SNP timer/stream writes), "
                      f"{t_scan * 1e3:.1f} ms; selectCaseControl alone {t_sel * 1e3:.1f} ms",
            "e2e_value": round(b / (t_sel + t_scan) / 1e9, 5)}


def cpu_baseline_pairwise(N, NCASE):
    import oracle
    O = oracle.Oracle()
    n = CPU_PAIRWISE_SAMPLE_SNPS
    codes, pheno = O.simulate(SEED, n, N, NCASE)
    pairs = n * (n - 1) // 2
    if oracle.have_ref():
        R = oracle.Ref(n, N, 5)
        R.add_codes(codes)
        R.set_case_control(pheno)
        t = R.time_phase(3, 1)
        kind = "reference"
        what = "compute(computeBoost): selectCaseControl + computeMargins + pre-screen + G-test"
    else:
        rows = O.pack_codes(codes)
        t0 = time.perf_counter()
        sel, nca, nco = O.select(rows, N, pheno)
        mar = O.margins(sel, nca, nco)
        O.boost_screen(sel, mar, nca, nco)
        t = time.perf_counter() - t0
        kind, what = "port", "oracle select + margins + boost_screen"
    return {"value": round(pairs / t, 1), "unit": "pairs/s", "cores": 1, "kind": kind,
            "sample": f"first {n} SNPs x {N} samples ({pairs} pairs): {what}, {t:.2f} s"}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation on all host cores
# ---------------------------------------------------------------------------------------------------
def _ref_worker(widx, N, NCASE, warmup, steps, start_evt, q):
    kind, obj, extra = _ref_marginal_sample(SEED + 1000 + widx, CPU_MARGINAL_SAMPLE_SNPS, N, NCASE)
    for _ in range(warmup):
        _marginal_step(kind, obj, extra)
    q.put(("ready", widx, kind))
    start_evt.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        _marginal_step(kind, obj, extra)
    q.put(("done", widx, time.perf_counter() - t0))


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build_oracle()
    N, NCASE = args.samples or MARGINAL["n_samples"], args.cases or MARGINAL["n_case"]
    cores = args.ref_procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    q, evt = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_ref_worker, args=(w, N, NCASE, args.warmup, args.steps, evt, q)) for w in range(cores)]
    for p in procs:
        p.start()
    kind = "port"
    for _ in range(cores):
        _, _, kind = q.get()
    t0 = time.perf_counter()
    evt.set()
    times = [q.get()[2] for _ in range(cores)]
    wall = time.perf_counter() - t0
    for p in procs:
        p.join()
    b = CPU_MARGINAL_SAMPLE_SNPS * N / 4.0
    value = cores * args.steps * b / max(times) / 1e9
    sample = (f"each of {cores} processes: {args.steps} steps of selectCaseControl + per-SNP case/control counts + "
              f"MinorAlleleFrequency over {CPU_MARGINAL_SAMPLE_SNPS} SNPs x {N} samples (select_cc_maf call surface)")
    line = {
        "impl": "reference", "metric": "marginal-scan GB/s vs HBM peak", "value": round(value, 5), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(max(times) / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 words + LUT popcount (reference)", "data": "synthetic",
        "config": {"workload": f"configs[1] shape: {NCASE} cases / {N - NCASE} controls, bounded sample of "
                               f"{CPU_MARGINAL_SAMPLE_SNPS} SNPs per process"},
        "cpu_baseline": {"value": round(value, 5), "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 5), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(wall, 2),
    }
    if not args.no_pairwise:
        line["pairwise"] = {"impl": "reference", **cpu_baseline_pairwise(args.pw_samples or PAIRWISE["n_samples"],
                                                                         args.pw_cases or PAIRWISE["n_case"])}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--snps", type=int, default=0)
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--cases", type=int, default=0)
    ap.add_argument("--pw-snps", type=int, default=0)
    ap.add_argument("--pw-samples", type=int, default=0)
    ap.add_argument("--pw-cases", type=int, default=0)
    ap.add_argument("--pw-steps", type=int, default=5)
    ap.add_argument("--no-pairwise", action="store_true")
    ap.add_argument("--no-missing", action="store_true", help="skip the section on a cohort with missing calls")
    ap.add_argument("--no-cfg3", action="store_true", help="N > 1: skip the configs[3] (5k/5k x 500k SNPs) pairwise section")
    ap.add_argument("--biobank", action="store_true", help="add the configs[4] biobank-scale section (needs ~110 GB of HBM)")
    ap.add_argument("--bb-snps", type=int, default=1_000_000)
    ap.add_argument("--bb-stream-snps", type=int, default=100_000)
    ap.add_argument("--bb-block", type=int, default=16_384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram__bytes_read+write per launch of the scan kernel from the committed ncu capture "
                         "(default: profiles/r1k_marginal_scan_full.md, valid for the default configs[1] shape only)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()

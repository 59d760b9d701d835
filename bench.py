#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native prover hot path.

Metric (BASELINE.json): BN254 G1 MSM throughput, Mpts/s, 2^20 synthetic random points / scalars per
GPU (configs[1]); weak scaling over N GPUs (each rank owns a contiguous point range of the same
size, one 128-byte partial sum per rank is gathered and folded -- SURVEY.md 8e).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 20]

One JSON line on stdout (rank 0).  `value` is timed with CUDA events on the library's stream with
points and scalars resident in HBM; `e2e` is the same MSM through the reference-facing C-ABI call
with the scalars in pinned HOST memory (H2D inside the timed region, affine result read back);
`roofline` is the bucket-accumulation kernel against the integer-MAD peak measured live on the same
GPU (MEASURED_PEAKS.json carries no integer peak); `cpu_baseline` is the oracle's restatement of the
reference's commit() loop timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bn254_g1_msm_throughput"
UNIT = "Mpts/s"
SEED_SCALARS = 0x5EED0001
SEED_POINTS = 0x5EED0002
MACS_PER_FP_MUL = 136                      # 8-limb CIOS: 8*(8+1+8) limb-MACs   
MODEL_ACC_MACS_PER_POINT = 160 * MACS_PER_FP_MUL  # SURVEY model: 16 windows x 10 Fp-mul (XYZZ mixed add) = 21 760


def windows_for(c):
    return (255 + c - 1) // c


def msm_macs_per_point(n):
    return MODEL_ACC_MACS_PER_POINT + 14 * MACS_PER_FP_MUL * (1 << 20) / n  # + bucket reduction, amortised


# ------------------------------------------------------------------ clocks sampling
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ reference arm (CPU)
def _ref_points(n):
    """n distinct valid G1 points, cheaply: P_0 = k*G, P_{i+1} = P_i + Q (one affine add each)."""
    from oracle import bn254
    p = bn254.g1_mul(bn254.G1, 0x1234567)
    q = bn254.g1_mul(bn254.G1, 0x7654321)
    out = []
    for _ in range(n):
        out.append(p)
        p = bn254.g1_add(p, q)
    return out


def _ref_worker(args):
    from oracle import bn254
    pts, scalars = args
    return bn254.g1_msm(pts, scalars)


def run_reference(args):
    """The reference's CPU path for this metric: kzg.commit's loop of affine double-and-add scalar
    multiplications (/root/reference/zkp/plonk/kzg.py:59-67 over py_ecc), as restated by
    oracle/bn254.g1_msm (py_ecc is not installable in this image: SURVEY F3), on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import bn254, synthetic
    cores = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    # size a step so the whole run takes about 60-90 s: ~13 ms per point per core
    per_point_s = 0.013
    budget_s = 75.0
    sample = int(budget_s * cores / ((steps + warmup) * per_point_s))
    sample = max(cores * 4, min(sample, 1 << 14))
    sample -= sample % cores
    pts = _ref_points(sample)
    scal = synthetic.scalars(SEED_SCALARS, sample)
    chunk = sample // cores
    jobs = [(pts[i * chunk:(i + 1) * chunk], scal[i * chunk:(i + 1) * chunk]) for i in range(cores)]
    with mp.Pool(cores) as pool:
        def step():
            parts = pool.map(_ref_worker, jobs)
            acc = None
            for p in parts:
                acc = bn254.g1_add(acc, p)
            return acc
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = step()
        dt = time.perf_counter() - t0
    assert bn254.g1_is_on_curve(res)
    value = sample * steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": "BN254 G1 MSM, 2^%d synthetic points per GPU (configs[1])" % args.log_n,
                   "step": "bounded sample of %d points of that workload" % sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d points/step, oracle/bn254.g1_msm = commit()-style affine double-and-add "
                                   "(reference kzg.py:59-67 on py_ecc semantics; py_ecc itself is not installable here)" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm (GPU)
def cpu_baseline(nat, table, scalars_handle):
    """Oracle (port of the reference's commit loop) on one host core, bounded sample; also a parity
    check of the GPU result on that sample."""
    from oracle import bn254
    sample = 2048                        # ~10 s of single-core CPU work at ~5 ms per point
    raw = nat.table_download(table, 0, sample)
    pts = [nat.g1_from_bytes(raw[64 * i:64 * i + 64]) for i in range(sample)]
    scal = nat.fr_vec_from_bytes(nat.scalars_download(scalars_handle, 0, sample))
    t0 = time.perf_counter()
    want = bn254.g1_msm(pts, scal)
    dt = time.perf_counter() - t0
    got = nat.g1_msm_dev(table, 0, scalars_handle, 0, sample)
    if got != want:
        raise SystemExit("PARITY FAILURE: GPU MSM differs from the oracle on the cpu_baseline sample")
    return {"value": sample / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
            "sample": "first %d points/scalars of the workload, oracle/bn254.g1_msm (reference kzg.commit loop, "
                      "affine double-and-add, py_ecc semantics); GPU result on the same sample is bit-identical" % sample}


def ntt_extra(nat, log_n):
    """Forward Fr NTT of 2^log_n resident elements: achieved integer and HBM rates."""
    import ctypes
    from interactive_zkp_study_b200 import _lib
    n = 1 << log_n
    omega = pow(5, (nat.R_MOD - 1) >> log_n, nat.R_MOD)
    h = nat.scalars_generate(0x5EED0004, n)
    wb = nat.fe_bytes(omega)

    def go():
        nat.check(_lib.lib().zkp_fr_ntt_dev(h.handle, 0, log_n, nat.buf(wb), 0, None))
    for _ in range(3):
        go()
    best = 1e9
    for _ in range(10):
        nat.timer_start()
        go()
        best = min(best, nat.timer_stop())
    h.free()
    macs = 68.0 * n * log_n
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {"n": n, "ms": best, "gelem_per_s": n / best / 1e6, "t_mac_per_s": macs / best / 1e9,
            "hbm_gbs_algorithmic": 64.0 * n / best / 1e6, "hbm_frac": 64.0 * n / best / 1e6 / hbm,
            "hbm_peak_gbs": hbm, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
            "note": "68*n*log2(n) limb-MACs and 64 B/element (one read + one write); the 254-bit NTT is integer bound (10.6 MAC/B needed)"}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ.setdefault("ZKP_B200_DEVICE", str(local_rank))
    from interactive_zkp_study_b200 import native as nat
    info = nat.device_info()

    n = 1 << args.log_n                 # points per GPU
    steps, warmup = args.steps, args.warmup
    G1 = nat.g1_bytes((1, 2))
    # this rank's point range [rank*n, (rank+1)*n): P_i = s_i * G with s_i from the synthetic stream
    # (known discrete logs -> O(n) check of any MSM), scalars k_i from a second stream
    def stream(seed, first, count):
        # element i of the global stream == element (i - first) of a stream with a shifted counter
        return nat.scalars_generate((seed + 64 * first) & ((1 << 64) - 1), count)
    s_h = stream(SEED_POINTS, rank * n, n)
    table = nat.g1_fixed_base_mul_dev(G1, s_h, n)
    plain_ms = None
    pre_c = 0
    if not args.plain:
        # first the plain-table number (no precomputation), for the record
        kk = stream(SEED_SCALARS + 0x9000, rank * n, n)
        for _ in range(3):
            nat.g1_msm_dev_partial(table, 0, kk, 0, n)
        nat.sync()
        nat.timer_start()
        for _ in range(5):
            nat.g1_msm_dev_partial(table, 0, kk, 0, n)
        plain_ms = nat.timer_stop() / 5
        kk.free()
        # static table -> window-precomputed layout, once (like loading the SRS)
        t0 = time.perf_counter()
        pre_c = nat.table_precompute(table)     # window width picked by the library for this table size
        precompute_s = time.perf_counter() - t0
    W_actual = windows_for(pre_c if pre_c else 16)
    n_vec = 4                            # rotate scalar vectors so no step reuses cached digits
    k_h = [stream(SEED_SCALARS + 0x1000 * v, rank * n, n) for v in range(n_vec)]

    def barrier():
        nat.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    exchange = None
    if world > 1:
        from interactive_zkp_study_b200 import sharded
        exchange = sharded.PartialExchange()   # CUDA send/recv buffers of the one 128 B-per-rank gather

    def sharded_step(k):
        """Local Pippenger -> XYZZ partial written into the NCCL send buffer -> all-gather over NVLink ->
        rank 0 folds straight out of the receive buffer (interactive_zkp_study_b200/sharded.py)."""
        return sharded.g1_msm_sharded(exchange, table, k, n)

    def step_resident(i):
        k = k_h[i % n_vec]
        if world == 1:
            return nat.g1_msm_dev(table, 0, k, 0, n)
        return sharded_step(k)

    # ---- correctness of the exact workload before timing (size-independent check)
    if rank == 0 and world == 1 and args.verify and args.log_n <= 22:
        from oracle import bn254, synthetic
        s = synthetic.scalars(SEED_POINTS, n)
        k = synthetic.scalars(SEED_SCALARS, n)
        want = bn254.g1_mul(bn254.G1, sum(a * b for a, b in zip(k, s)) % bn254.R)
        if step_resident(0) != want:
            raise SystemExit("PARITY FAILURE: 2^%d MSM != (sum k_i s_i) * G" % args.log_n)

    if world > 1 and args.verify and n * world <= (1 << 23):
        # sharded MSM == (sum over ALL ranks' ranges of k_i s_i) * G, checked on rank 0.  The expected value is
        # seconds of pure-Python work on rank 0: the other ranks wait for it on the HOST (a gloo barrier), not
        # inside a collective spinning on their GPUs -- with the peers parked in NCCL for that long, the gathers
        # of the following ~40 steps measured 3.4 ms instead of 0.1 ms (2 and 4 ranks).
        want = None
        if rank == 0:
            from oracle import bn254, synthetic
            s = synthetic.scalars(SEED_POINTS, n * world)
            k = synthetic.scalars(SEED_SCALARS, n * world)
            want = bn254.g1_mul(bn254.G1, sum(a * b for a, b in zip(k, s)) % bn254.R)
            del s, k
        host_group = dist.new_group(backend="gloo")
        dist.barrier(group=host_group)
        got = step_resident(0)
        if rank == 0 and got != want:
            raise SystemExit("PARITY FAILURE: sharded MSM over %d GPUs != (sum k_i s_i) * G" % world)

    # clocks are sampled from here to the end of the timed region: the region itself lasts < 0.1 s and
    # nvidia-smi delivers a sample every ~50-100 ms, so the integer-peak microbenchmarks (also a
    # saturating integer load) and the warm-up steps are included to get a meaningful median
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    peak = {}
    if rank == 0:
        peak = {"imad_wide_u32": nat.imad_peak(0), "imad_lo": nat.imad_peak(1), "imad_hi_u32": nat.imad_peak(2),
                "fp_mul_chain": nat.imad_peak(3)}

    # ---- resident-input timing (value)
    for i in range(warmup):
        step_resident(i)
    nat.msm_profile(True)
    barrier()
    launches0 = nat.launch_count()
    acc_us = 0.0
    msm_us = 0.0
    nat.timer_start()
    for i in range(steps):
        step_resident(warmup + i)
        acc_us += nat.msm_last_profile("accumulate")
        msm_us += nat.msm_last_profile(None)
    ms = nat.timer_stop()
    barrier()
    launches = nat.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    nat.msm_profile(False)
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    total_points = n * world
    value = total_points * steps / (ms * 1e-3) / 1e6

    # for the record: the same MSMs submitted as ONE pipelined batch call (two streams: the tail of MSM k
    # overlaps the accumulation of MSM k+1), the call shape of a prover that commits in groups
    pipelined = None
    if world == 1:
        items = [(k_h[i % n_vec], 0, 0, n) for i in range(min(steps, 8))]
        nat.g1_msm_dev_batch(table, items)
        nat.timer_start()
        nat.g1_msm_dev_batch(table, items)
        bms = nat.timer_stop()
        pipelined = {"msms_per_call": len(items), "ms_per_msm": bms / len(items),
                     "value": n * len(items) / (bms * 1e-3) / 1e6, "unit": UNIT, "call": "zkp_g1_msm_dev_batch"}

    # ---- end-to-end timing: scalars in pinned host memory, H2D inside, affine result read back
    pinned = nat.PinnedBuffer(32 * n)
    pinned.write(nat.scalars_download(k_h[0], 0, n))

    e2e_scalars = nat.scalars_alloc(n) if world > 1 else None

    def step_e2e():
        if world == 1:
            return nat.g1_msm_table(table, 0, pinned.addr, n)
        nat.scalars_upload(e2e_scalars, 0, pinned.addr, n)   # H2D of this rank's scalar shard (inside the timed region)
        return sharded_step(e2e_scalars)

    for _ in range(max(1, warmup // 2)):
        step_e2e()
    barrier()
    if os.environ.get("ZKP_BENCH_DEBUG") and world > 1:      # diagnosis: host-side split of the e2e step, untimed
        tu = tm = tg = tc = 0.0
        for _ in range(steps):
            t0 = time.perf_counter()
            nat.scalars_upload(e2e_scalars, 0, pinned.addr, n)
            t1 = time.perf_counter()
            nat.g1_msm_dev_partial(table, 0, e2e_scalars, 0, n, out_addr=exchange.send_addr)
            t2 = time.perf_counter()
            exchange.all_gather()
            t3 = time.perf_counter()
            if rank == 0:
                nat.g1_combine_partials(exchange.recv_addr, world)
            t4 = time.perf_counter()
            tu += t1 - t0; tm += t2 - t1; tg += t3 - t2; tc += t4 - t3
        sys.stderr.write("[rank %d] tight loop per step: upload %.2f, partial MSM %.2f, gather %.2f, combine %.2f ms\n"
                         % (rank, tu / steps * 1e3, tm / steps * 1e3, tg / steps * 1e3, tc / steps * 1e3))
        barrier()
    nat.timer_start()
    for _ in range(steps):
        step_e2e()
    e2e_ms = nat.timer_stop()
    if os.environ.get("ZKP_BENCH_DEBUG"):
        sys.stderr.write("[rank %d] e2e loop: %.2f ms for %d steps (device events)\n" % (rank, e2e_ms, steps))
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = total_points * steps / (e2e_ms * 1e-3) / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (bucket accumulation) against the measured integer peak
    acc_s = acc_us / steps * 1e-6
    peak_gmacs = max(peak["imad_hi_u32"], peak["imad_wide_u32"])
    acc_macs = n * W_actual * 10 * MACS_PER_FP_MUL   # the mixed additions this launch really performs
    achieved = acc_macs / acc_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "accumulate_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "imad", "kernel": "msm_accumulate_kernel<Fp>", "achieved": achieved, "peak": peak_gmacs / 1e3,
        "unit": "T limb-MAC/s (one IMAD.WIDE.U32 = 32x32+64->64)", "frac": achieved / (peak_gmacs / 1e3),
        "traffic": traffic,
        "algorithmic_macs_per_launch": acc_macs,
        "algorithmic_note": "%d windows x 10 Fp-mul (XYZZ mixed add) x 136 limb-MAC per point; the SURVEY 8d model "
                            "(16 windows, 21 760 MAC/pt) is used for whole_msm.frac" % W_actual,
        "kernel_ms": acc_s * 1e3, "kernel_share_of_step": acc_us / max(msm_us, 1e-9),
        "peak_source": "measured live on this GPU by zkp_imad_peak (dependent-free IMAD.HI.U32 / IMAD.WIDE.U32 chains); "
                       "MEASURED_PEAKS.json has no integer peak",
        "peaks_gmacs": peak,
        "whole_msm": {"macs_per_point_model": msm_macs_per_point(n),
                      "frac": (n * msm_macs_per_point(n) / (msm_us / steps * 1e-6) / 1e12) / (peak_gmacs / 1e3)},
        "hbm_view": {"algorithmic_bytes_per_launch": n * W_actual * 68,
                     "achieved_gbs": n * W_actual * 68 / acc_s / 1e9,
                     "note": "windows x (64 B point gather + 4 B index) per point: far below the HBM roof, the kernel is integer bound"},
    }
    base = cpu_baseline(nat, table, k_h[0]) if world == 1 else None
    extras = {}
    if world == 1 and args.extras and args.log_n <= 20:
        # second half of the BASELINE metric: Groth16 prove ms @2^20 constraints (config 3), and the Fr NTT
        for h in [table] + k_h:
            h.free()
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import groth16_large
            extras["groth16_prove_2^%d" % args.log_n] = groth16_large.run(args.log_n, 3, verify=True, quiet=True)
        except Exception as e:  # never lose the headline line
            extras["groth16_prove_error"] = repr(e)
        try:
            import qap_large
            extras["groth16_r1cs_prove_2^%d" % args.log_n] = qap_large.run(args.log_n, 3, quiet=True)
        except Exception as e:
            extras["groth16_r1cs_prove_error"] = repr(e)
        try:
            import plonk_large
            extras["plonk_prove_2^%d" % args.log_n] = plonk_large.run(args.log_n, 3, verify=True, quiet=True)
        except Exception as e:
            extras["plonk_prove_error"] = repr(e)
        try:
            extras["fr_ntt"] = ntt_extra(nat, args.log_n)
        except Exception as e:
            extras["fr_ntt_error"] = repr(e)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (254-bit modular integers, Montgomery)", "data": "synthetic",
        "config": {
            "workload": "BN254 G1 MSM, 2^%d synthetic points per GPU (configs[1]); P_i = s_i*G, scalars uniform in [0,r)" % args.log_n,
            "points_per_gpu": n, "total_points": total_points,
            "sharding": "contiguous point ranges, one 128 B XYZZ partial per rank gathered over NCCL" if world > 1 else "single GPU",
            "table": ("window-precomputed static table T[w][i] = 2^(%d w) P_i, %d windows, %.0f MiB per GPU, built once in %.0f ms "
                      "(zkp_g1_table_precompute; SRS/CRS tables are fixed per circuit)" % (pre_c, W_actual, W_actual * n * 64 / 2**20, precompute_s * 1e3))
                     if pre_c else "plain affine table, 64 B/point",
            "plain_table": None if plain_ms is None else {"ms_per_step": plain_ms, "value": n / plain_ms / 1e3, "unit": UNIT,
                                                          "note": "same MSM without the precomputed windows (16 windows + Horner), this rank only"},
            "l2": "working set (point table + scalars + 2x60 MiB digit/index arrays + bucket partials) exceeds the 126 MB L2; "
                  "a different scalar vector every step",
            "device": info["name"],
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / steps,
                "h2d_bytes_per_step": 32 * n * world, "d2h_bytes_per_step": 64,
                "path": "zkp_g1_msm_table: device-resident point table (static SRS), scalars from pinned host memory"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": base,
        "pipelined_batch": pipelined,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 of the points per GPU")
    ap.add_argument("--no-verify", dest="verify", action="store_false")
    ap.add_argument("--plain", action="store_true", help="do not precompute the window table")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the Groth16-prove and NTT extras")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- images/sec of the 2-stack hourglass train step (BASELINE.json configs[1]:
S=2, C=256, bs=24 per GPU, 256x256, fwd + sum-of-stacks MSE + bwd + RMSprop [+ grad all-reduce]).

    python bench.py [--gpus N --steps K --warmup W]          our arm (one process per GPU under torchrun)
    python bench.py --impl reference [...]                   the reference's CPU implementation, host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec 2-stack HG bs24 256x256 train step"
BYTES_PER_IMAGE_STEP = 0.892e9      # algorithmic HBM bytes / image / train step at S=2


def bytes_per_image_step(stacks):
    """0.892 GB/image at S=2 and 3.15 GB/image at S=8 (the 7.81 ms floor of configs[4] at bs 16); the stem and
    the three residuals in front of the stacks are the constant part, every stack adds the same amount."""
    per_stack = (3.15e9 - BYTES_PER_IMAGE_STEP) / 6.0
    return BYTES_PER_IMAGE_STEP + (stacks - 2) * per_stack
FLOP_PER_IMAGE_STEP = 50.0e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super(ClockSampler, self).__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(args, rank):
    from pose_adv_aug_b200 import synth
    x = synth.make_images(args.batch, args.res, seed=100 + rank)
    t = synth.make_heatmaps(args.batch, args.res, 16, seed=200 + rank)
    return x, t


def make_weights(args):
    """Seeded reference-initialiser-shaped weights, keyed by the drop-in module's own state_dict schema (== the
    reference's, tests/test_abi_cpu.py); both arms load the same dict.  Nothing under oracle/ is touched here."""
    from pose_adv_aug_b200 import synth
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    return synth.make_state_dict(synth.schema_of(M.create_hg(args.stacks, 1, 16, args.chan)), seed=1, perturb_bn=False)


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref when present, else the oracle port)
# ----------------------------------------------------------------------------------------------
class CpuStep(object):
    def __init__(self, args, sd):
        import torch
        from oracle import make_ref
        from oracle import hg_oracle as O
        self.torch, self.O, self.args = torch, O, args
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        ref, _ = make_ref.load()
        self.kind = "reference" if ref is not None else "port"
        if ref is not None:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                self.net = ref.create_hg(num_stacks=args.stacks, num_modules=1, num_classes=16, chan=args.chan)
            self.net.load_state_dict(sd)
            self.net.train()
            self.opt = torch.optim.RMSprop(self.net.parameters(), lr=2.5e-4, alpha=0.99, eps=1e-8, momentum=0,
                                           weight_decay=0)
        else:
            from collections import OrderedDict
            self.sd = OrderedDict((k, v.clone()) for k, v in sd.items())
            self.sq = OrderedDict((k, torch.zeros_like(v)) for k, v in self.sd.items() if O.is_trainable(k))

    def step(self, x, t):
        """One reference train step (stack-hg.py:153-165) on CPU; returns (outputs, loss)."""
        if self.kind == "reference":
            outs = self.net(x)
            loss = 0
            for o in outs:
                tmp = (o - t) ** 2
                loss = loss + tmp.sum() / tmp.numel()
            self.opt.zero_grad()
            loss.backward()
            self.opt.step()
            return [o.detach() for o in outs], float(loss)
        outs, loss, _, _ = self.O.train_step(self.sd, x, t, self.args.stacks, 1, square_avg=self.sq)
        return outs, float(loss)


def cpu_sample(args, sd, x, t, budget_s, steps, warmup):
    """Times reference train steps on the host cores at the FULL per-GPU batch of the workload (the batch size changes the
    BatchNorm statistics and the per-image cost, so a smaller sample batch would be a different workload); the sample is
    bounded through the NUMBER of steps: at most `steps` timed steps after `warmup`, fewer when budget_s would be exceeded
    (never fewer than one).  Returns the measured rate and what the sample was."""
    cpu = CpuStep(args, sd)
    sb = args.batch
    b = min(2, sb)
    cpu.step(x[:b], t[:b])                      # thread pool / allocator warm-up on a tiny batch (untimed)
    t0 = time.time()
    outs, loss = cpu.step(x[:sb], t[:sb])       # first full-batch step: timed, counts as warm-up unless it is all we can afford
    first = time.time() - t0
    n_warm = max(0, min(warmup, int(budget_s / max(first, 1e-6)) - 1) - 1)
    for _ in range(n_warm):
        cpu.step(x[:sb], t[:sb])
    left = budget_s - first * (1 + n_warm)
    n_timed = max(0, min(steps, int(left / max(first, 1e-6))))
    if n_timed == 0:
        dt, n_timed, n_warm_total = first, 1, 0
    else:
        t0 = time.time()
        for _ in range(n_timed):
            outs, loss = cpu.step(x[:sb], t[:sb])
        dt = (time.time() - t0) / n_timed
        n_warm_total = 1 + n_warm
    return {"value": sb / dt, "unit": "images/s", "cores": cpu.cores, "kind": cpu.kind,
            "sample": "%d timed train steps (fwd+MSE+bwd+RMSprop, %d full-batch warm-up) at batch %d of the same "
                      "S=%d C=%d %dx%d workload, torch CPU fp32, %d threads" %
                      (n_timed, n_warm_total, sb, args.stacks, args.chan, args.res, args.res, cpu.cores),
            "s_per_step": dt, "sample_batch": sb, "timed_steps": n_timed}


def cpu_config1(steps=5):
    """BASELINE.json configs[0] -- the reference's own CPU-runnable case: 1-stack hourglass, nFeat=128, bs=2, 64x64,
    forward + MSE on the host cores (reference model when oracle/_ref travelled, else the oracle port)."""
    import argparse as _ap
    import torch
    a = _ap.Namespace(stacks=1, chan=128, batch=2, res=64)
    sd = make_weights(a)
    x, t = build_inputs(a, 0)
    cpu = CpuStep(a, sd)

    def fwd_mse():
        with torch.no_grad():
            if cpu.kind == "reference":
                outs = cpu.net(x)
            else:
                outs, _ = cpu.O.hg_forward(cpu.sd, x, 1, 1, training=True)
            return float(sum(((o - t) ** 2).sum() / o.numel() for o in outs))
    fwd_mse()
    t0 = time.time()
    for _ in range(steps):
        loss = fwd_mse()
    dt = (time.time() - t0) / steps
    return {"images_per_s": a.batch / dt, "ms": dt * 1e3, "loss": loss, "kind": cpu.kind, "cores": cpu.cores,
            "workload": "1-stack hourglass nFeat=128, bs=2, 64x64 synthetic, CPU fwd+MSE"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    sd = make_weights(args)
    x, t = build_inputs(args, 0)
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    r = cpu_sample(args, sd, x, t, budget_s=150.0, steps=steps, warmup=warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": r["timed_steps"], "warmup": warmup, "steps_requested": steps, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, 1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "torch": torch.__version__}
    print(json.dumps(line))


def config_dict(args, world):
    return {"workload": "%d-stack hourglass train step (fwd + sum-of-stacks MSE + bwd + RMSprop%s), S=%d C=%d, "
                        "bs=%d/GPU, %dx%d synthetic MPII-shaped batch" %
                        (args.stacks, " + 1 NCCL grad all-reduce" if world > 1 else "", args.stacks, args.chan, args.batch,
                         args.res, args.res),
            "stacks": args.stacks, "chan": args.chan, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "res": args.res, "parallelism": "dp%d" % world,
            "l2": "no explicit flush: each step streams >5 GB of activations (>> 126 MB L2) between reuses"}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pose_adv_aug_b200 import dist as hdist
    from pose_adv_aug_b200 import HourglassTrainer, get_lib
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    rank, world, local = hdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = get_lib()
    assert lib.cdll.hgk_device_ok() == 1, "libhgk needs an sm_100 GPU"
    pk = peaks()
    sd = make_weights(args)
    x, t = build_inputs(args, rank)
    net = M.create_hg(args.stacks, 1, 16, args.chan)
    net.load_state_dict(sd)
    tr = HourglassTrainer(net, args.batch, args.res, device=dev, use_graph=not args.no_graph, n_streams=args.streams,
                          n_low=args.low_streams)
    xp, tp = x.pin_memory(), t.pin_memory()
    tr.x.copy_(xp)
    tr.t.copy_(tp)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        tr.step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    # ---- timed region: K device-resident steps ----
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        tr.step_resident()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    # ---- e2e: public API with host buffers (H2D of the batch + D2H of the loss every step) ----
    e2e_steps = max(3, min(2 * args.steps, 40))      # wall-clock timed (host in the loop): more steps, less jitter
    for _ in range(3):                      # untimed warm-up of the double-buffered input path (staging buffers, copy stream)
        tr.prefetch(xp, tp)
        tr.step_prefetched()
    barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    last = 0.0
    # public API with host buffers, the upload of the next batch issued while the current step computes (what a
    # DataLoader with pin_memory does for the reference loop's `.cuda(async=True)`): every timed step contains the H2D
    # copy of ITS inputs (issued one step earlier on the copy stream; the first one is issued inside the region) and the
    # D2H read of its loss
    # ... read back (4 bytes into pinned memory) while the NEXT step already runs: every step's loss reaches the host inside
    # the timed region, one step late, and the GPU does not idle across the host's read
    tr.prefetch(xp, tp)
    for i in range(e2e_steps):
        tr.step_prefetched(loss_to_host=True)
        if i + 1 < e2e_steps:
            tr.prefetch(xp, tp)
        if i > 0:
            last = tr.pop_loss()
    last = tr.pop_loss()
    torch.cuda.synchronize()
    barrier()
    e2e_s = time.time() - t0
    # the same with the loss read right after its own step (loss.item(): the host waits for the step, then launches the next)
    barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    tr.prefetch(xp, tp)
    for i in range(e2e_steps):
        loss_dev = tr.step_prefetched()
        if i + 1 < e2e_steps:
            tr.prefetch(xp, tp)
        last = float(loss_dev.item())
    torch.cuda.synchronize()
    barrier()
    e2e_sync_s = time.time() - t0
    if os.environ.get("HGK_BENCH_E2E_REPEAT"):        # developer: repeatability of the two input paths
        for rep in range(3):
            torch.cuda.synchronize(); ta = time.time()
            tr.prefetch(xp, tp)
            for i in range(e2e_steps):
                loss_dev = tr.step_prefetched()
                if i + 1 < e2e_steps:
                    tr.prefetch(xp, tp)
                last = float(loss_dev.item())
            torch.cuda.synchronize(); tb = time.time()
            for _ in range(e2e_steps):
                last = float(tr.step(xp, tp).item())
            torch.cuda.synchronize(); tc = time.time()
            sys.stderr.write("e2e repeat %d: prefetched %.3f ms/step, serial %.3f ms/step\n" %
                             (rep, (tb - ta) / e2e_steps * 1e3, (tc - tb) / e2e_steps * 1e3))
    # the same without lookahead: copies, then the step, then the loss read, all serial
    barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(e2e_steps):
        last = float(tr.step(xp, tp).item())
    torch.cuda.synchronize()
    barrier()
    e2e_serial_s = time.time() - t0
    w2 = time.time()
    sampler.stop()
    tms = torch.tensor([ms, e2e_s * 1e3, e2e_serial_s * 1e3, e2e_sync_s * 1e3], device=dev, dtype=torch.float64)
    by_rank = None
    if world > 1:
        mine = torch.tensor([ms / args.steps], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        by_rank = [round(float(v), 4) for v in allr]
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_serial_ms, e2e_sync_ms = float(tms[0]), float(tms[1]), float(tms[2]), float(tms[3])
    clocks = sampler.summary(w0, w1)
    # ---- dominant kernel, timed live with CUDA events on the launching stream ----
    roof = dominant_kernel_roofline(tr, pk, torch)
    if rank != 0:
        if world > 1:
            tr.close()          # the step graph holds captured NCCL kernels: release it before the communicator goes
            dist.destroy_process_group()
        return
    value = world * args.batch * args.steps / (ms / 1e3)
    ms_step = ms / args.steps
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, world),
            "roofline": roof,
            "step_roofline": {"bound": "hbm", "achieved": args.batch * bytes_per_image_step(args.stacks) / (ms_step / 1e3) / 1e9,
                              "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "frac": args.batch * bytes_per_image_step(args.stacks) / (ms_step / 1e3) / 1e9 / pk["hbm_gbs"],
                              "note": "whole step, algorithmic %.3f GB/image/step of the fused plan; peak %s"
                                      % (bytes_per_image_step(args.stacks) / 1e9, pk["src"])},
            "e2e": {"value": world * args.batch * e2e_steps / (e2e_ms / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": int(xp.numel() * 4 + tp.numel() * 4), "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "serial_value": world * args.batch * e2e_steps / (e2e_serial_ms / 1e3),
                    "sync_loss_value": world * args.batch * e2e_steps / (e2e_sync_ms / 1e3),
                    "api": "HourglassTrainer.prefetch(images_pinned, heatmaps_pinned) [next batch, copy stream] + "
                           "step_prefetched(loss_to_host=True) + pop_loss() [every step's loss copied to pinned host memory "
                           "and read one step later, while the next step runs]; sync_loss_value = the same with "
                           "step_prefetched() -> loss.item() right after each step; serial_value = HourglassTrainer.step("
                           "images_pinned, heatmaps_pinned) -> loss.item() with the copies in front of the step"},
            "gpu_launches": tr.launches_per_step * args.steps, "launches_per_step": tr.launches_per_step,
            "cuda_graph": not args.no_graph, "graph_streams": args.streams, "low_priority_streams": args.low_streams, "clocks": clocks, "loss": last, "conv_path": M.CONV_PATH}
    if world > 1:
        line["ms_per_step_by_rank"] = by_rank        # device-timed per rank (equal when every step ends in a collective)
        if os.environ.get("HGK_AR_SKIP", "0") == "1":
            line["diagnostic"] = "HGK_AR_SKIP=1: NO gradient all-reduce (timing diagnostic: each rank at its own pace; not a valid bench line)"
        line["allreduce"] = {"in_graph": bool(tr.ar_in_graph), "graph_launches_per_step": 1 if tr.ar_in_graph else 2,
                             "buckets_mb": [round((hi - lo) * 4e-6, 2) for lo, hi in tr.bucket_ranges()],
                             "note": "NCCL sum of the flat gradient buffer, bucketed back to front on a communication stream "
                                     "and captured inside the step graph" if tr.ar_in_graph else
                                     "one eager NCCL all-reduce between two graphs"}
    if world == 1 and not args.no_cpu_baseline:
        # ---- cpu_baseline leg (the only place this arm executes anything under oracle/): the reference's CPU step on
        #      the host cores, timed on a bounded sample, and -- as the checker -- its heat-maps against ours ----
        try:
            r = cpu_sample(args, sd, x, t, budget_s=25.0, steps=2, warmup=0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["parity"] = parity_vs_cpu(args, sd, x, t, torch, dev)
            try:
                line["cpu_baseline"]["config1"] = cpu_config1()
            except Exception as e:
                line["cpu_baseline"]["config1"] = {"failed": str(e)}
        except Exception as e:      # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %s" % e}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        tr.close()
        dist.destroy_process_group()


def parity_vs_cpu(args, sd, x, t, torch, dev):
    """fp32 heat-map max-abs-err (BASELINE.json metric, second half) on a 2-image sample."""
    from collections import OrderedDict
    from oracle import hg_oracle as O
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    n = 2
    outs_ref, loss_ref, _, _ = O.train_step(OrderedDict((k, v.clone()) for k, v in sd.items()), x[:n], t[:n],
                                            args.stacks, 1)
    net = M.create_hg(args.stacks, 1, 16, args.chan)
    net.load_state_dict(sd)
    net.to(dev).train()
    with torch.no_grad():
        outs = net(x[:n].to(dev))
    err = max(float((a.cpu() - b).abs().max()) for a, b in zip(outs, outs_ref))
    rel = max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(outs, outs_ref))
    return {"heatmap_max_abs_err": err, "heatmap_rel_to_max": rel, "sample": "train-mode forward, %d images" % n,
            "oracle": "oracle/hg_oracle.py fp32 CPU"}


def dominant_kernel_roofline(tr, pk, torch):
    """Times the FLOP-dominant launch class of the step (3x3 conv, C/2 -> C/2 at the top resolution,
    forward) with CUDA events, cycling over the plan's distinct launches of that class so that
    inputs are not L2-resident between repetitions."""
    plan = tr.plan
    cand = []
    for rec in plan.fwd:
        a = rec[1]
        if rec[2] == "conv_nhwc":
            N, H, W, Cin, k, Cout = a[4], a[5], a[6], a[7], a[9], a[12]
        elif rec[2] in ("conv_tc_nhwc", "conv_tc_bn_nhwc", "conv_tc_bn_x2_nhwc", "conv_tc_x2_nhwc"):
            N, H, W, Cin, k, Cout = a[4], a[5], a[6], a[7], a[10], a[12]
        else:
            continue
        if k == 3:
            cand.append((2.0 * N * H * W * Cin * Cout * 9, rec))
    if not cand:
        return None
    top = max(c[0] for c in cand)
    recs = [r for f, r in cand if f == top]
    # several layer shapes can share the top FLOP count (64->64 @128x128 and 128->128 @64x64): keep the most frequent one
    import collections
    shape_of = lambda r: tuple(r[1][4:8])
    common = collections.Counter(shape_of(r) for r in recs).most_common(1)[0][0]
    recs = [r for r in recs if shape_of(r) == common]
    stream = torch.cuda.current_stream().cuda_stream
    for r in recs:
        r[0](*r[1], stream)
    torch.cuda.synchronize()
    reps = 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for r in recs:
            r[0](*r[1], stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(recs))
    a = recs[0][1]
    N, H, W, Cin, Cout = a[4], a[5], a[6], a[7], a[12]
    tc = recs[0][2] in ("conv_tc_nhwc", "conv_tc_bn_nhwc", "conv_tc_bn_x2_nhwc", "conv_tc_x2_nhwc")
    x2 = recs[0][2] == "conv_tc_bn_x2_nhwc"
    achieved = top / (ms / 1e3) / 1e12
    tile = tc and H % 16 == 0 and W % 16 == 0
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch: read from the newest committed `ncu --set full`
    # summary (profiles/ncu_traffic.json, written by tools/ncu_summary.py with the commit it was captured at); None when
    # no capture of this layer shape is committed
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for ent in tj.get("entries", []):
            if ent.get("shape") == [N, H, W, Cin, Cout] and ent.get("ksize") == 3 and ent.get("direction") == "fwd":
                traffic = ent["dram_bytes_per_launch"]
                traffic_src = "%s @ %s (%s)" % (ent.get("file"), ent.get("commit"), ent.get("kernel"))
    except Exception:
        pass
    return {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch, from the committed ncu --set full
            # capture profiles/r2z_kernels_ncu.txt (24x64x64, 128->128: 51.6 MB read + 7.3 MB written back before the
            # kernel ended; the 50 MB output largely stays dirty in the 126 MB L2)
            "traffic": traffic, "traffic_src": traffic_src,
            "kernel": "%s 3x3 %d->%d @ %dx%dx%d (fwd, %d launches of this FLOP class/step)" %
                      ((("conv_tc2_kernel (tcgen05 image-tile kernel, TF32 + 2xBF16)" if x2 else
                         "conv_tc2_kernel (tcgen05 image-tile kernel, 3xTF32)") if tile else "conv_tc_kernel (tcgen05, 3xTF32)")
                       if tc else "conv_igemm_simt (fp32 FFMA)", Cin, Cout, N, H, W, len(recs)),
            "ms_per_launch": ms, "flop_per_launch": top, "algorithmic_bytes_per_launch": 4.0 * N * H * W * (Cin + Cout),
            # tensor-core work in TF32-equivalents: 3 TF32 MMAs per product (3xTF32), or 1 TF32 + 2 BF16 at twice the rate = 2
            "mma_flop_per_launch": top * ((2 if x2 else 3) if tc else 1),
            "frac_of_own_tensor_ceiling": achieved / (pk["bf16_tflops"] / (4.0 if x2 else 6.0)) if tc else None,
            "peak_src": pk["src"] + " dense bf16 burst; achieved counts algorithmic fp32 FLOPs (per fp32-class product the kernel "
                        "issues " + ("one TF32 MMA at 1/2 the bf16 rate and two bf16 MMAs: its own ceiling is peak/4)" if x2 else
                                     "3 TF32 MMAs at 1/2 the bf16 rate: its own ceiling is peak/6)")}


def run_config3(args):
    """BASELINE.json configs[2]: one agent-augmented joint-train iteration (joint-train-pose-s-r-agent.py:245-296) at bs 24:
    half-hourglass forward (hg.train(), agent.eval()) -> ASN (scale, rotation) distributions -> on-GPU sampling ->
    the batch re-augmented with the sampled scales / rotations ON THE GPU from resident photographs (agent.AgentBatchLoader =
    the reference's load_batch_data DataLoader round trip) -> full hourglass train step (CUDA graph) -> PCK on the GPU; plus the agent update
    (train_agent_sr, :323-410: hg.eval(), agent.train(), KL loss, ASN backward, flat RMSprop) timed separately."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from pose_adv_aug_b200 import synth, HourglassTrainer, FlatRMSprop, agent
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    from pose_adv_aug_b200.pylib import Evaluation, HumanPts
    dev = torch.device("cuda", 0)
    S, C, N, R = 2, 256, 24, 256
    net = M.create_hg(S, 1, 16, C)
    net.load_state_dict(synth.make_state_dict(synth.schema_of(net), seed=1, perturb_bn=False))
    asn = M.create_asn(C, C, 7, 7, is_aug=True)
    asn.load_state_dict(synth.make_state_dict(synth.schema_of(asn), seed=2, perturb_bn=False))
    asn.to(dev)
    tr = HourglassTrainer(net, N, R, device=dev, use_graph=True)
    x = synth.make_images(N, R, seed=100).to(dev)
    tr.x.copy_(x)
    tr.t.copy_(synth.make_heatmaps(N, R, 16, seed=200))
    pts = torch.randint(4, 60, (N, 16, 2), device=dev).float()
    np.random.seed(0)
    # the re-augmentation of the batch with the agent's sampled (scale, rotation) (load_batch_data, joint-train...:425-450):
    # 24 MPII-sized synthetic photographs resident in HBM, cropped / rotated / resized on the GPU (agent.AgentBatchLoader)
    rng = np.random.default_rng(5)
    photos = [torch.from_numpy(np.ascontiguousarray(np.transpose(synth.make_photo(720, 1280, 500 + k), (2, 0, 1)))).to(dev)
              for k in range(N)]
    annos = []
    for k in range(N):
        j = np.concatenate([rng.uniform(300, 980, (16, 1)), rng.uniform(150, 570, (16, 1)), np.ones((16, 1))], axis=1)
        annos.append({"joint_self": j.tolist(), "objpos": [float(rng.uniform(400, 880)), float(rng.uniform(250, 470))],
                      "scale_provided": float(rng.uniform(0.8, 3.0)), "normalizer": float(rng.uniform(40, 120))})
    loader = agent.AgentBatchLoader(photos, annos)
    img_index = list(range(N))

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def iteration_synthetic():
        net.train(); asn.eval()
        with torch.no_grad():
            ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
        agent.sample_scale_rotation(ps, pr)
        hm, _ = HumanPts.pts2heatmap(pts, [64, 64])
        tr.t.copy_(hm)
        tr.step_resident()
        return Evaluation.accuracy(tr.heatmaps()[-1], tr.t, list(range(16)))

    def iteration():
        net.train(); asn.eval()
        with torch.no_grad():
            ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
        _, _, si, ri = agent.sample_scale_rotation(ps, pr)
        img, hm, c, s_, r_, gp, nz = loader.load_batch(si.tolist(), ri.tolist(), img_index)      # 48 indices D2H, as the reference
        tr.x.copy_(img)
        tr.t.copy_(hm)
        tr.step_resident()
        return Evaluation.accuracy(tr.heatmaps()[-1], tr.t, list(range(16)))

    def load_only():
        loader.load_batch([3] * N, [2] * N, img_index)

    opt = FlatRMSprop(asn, lr=2.5e-4)
    tgt = torch.softmax(torch.randn(N, 7, device=dev), dim=1)

    def agent_update():
        net.eval(); asn.train()
        ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
        loss = F.kl_div(torch.log(F.softmax(ps, dim=1) + 1e-7), tgt, reduction="mean") * 7 + \
            F.kl_div(torch.log(F.softmax(pr, dim=1) + 1e-7), tgt, reduction="mean") * 7
        opt.zero_grad()
        loss.backward()
        opt.step()

    steps, warmup = args.steps, max(args.warmup, 3)
    ms = timed(iteration, steps, warmup)
    ms_syn = timed(iteration_synthetic, steps, warmup)
    ms_load = timed(load_only, steps, warmup)
    ms_up = timed(agent_update, steps, warmup)
    print(json.dumps({
        "metric": "images/sec 2-stack HG + ASN agent joint-train iteration bs24 256x256", "value": N / ms * 1e3,
        "unit": "images/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: half-hg(train BN)+ASN(eval) forward, on-GPU (s, r) sampling, the batch "
                               "re-augmented on the GPU with the sampled scales / rotations from 24 resident 720x1280 photographs "
                               "(AgentBatchLoader: flip, colour, crop / rotate / resize, target rendering), full 2-stack train step "
                               "(CUDA graph), GPU PCK; S=2 C=256 bs=24 256x256",
                   "batch_per_gpu": N, "parallelism": "dp1"},
        "without_loader": {"ms_per_step": ms_syn, "what": "the same iteration on a fixed pre-cropped batch (targets rendered from "
                                                         "fixed joint coordinates): the number of the earlier rounds"},
        "loader": {"ms_per_batch": ms_load, "what": "AgentBatchLoader.load_batch alone (host geometry + ~10 launches per image)"},
        "agent_update": {"ms_per_update": ms_up, "images_per_s": N / ms_up * 1e3,
                         "what": "train_agent_sr: half-hg(eval) + ASN(train) forward + KL + ASN backward + flat RMSprop (module path)"},
        "gpu_launches": None, "launches_per_train_step": tr.launches_per_step}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5],
                    help="BASELINE.json configuration (1-based): 2 = the bench line (default), 3 = agent-augmented joint-train "
                         "iteration, 5 = 8-stack hourglass bs 16")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--stacks", type=int, default=2)
    ap.add_argument("--chan", type=int, default=256)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--conv-path", type=int, default=None)
    ap.add_argument("--streams", type=int, default=10)
    ap.add_argument("--low-streams", type=int, default=3, help="of --streams, low-priority streams reserved for weight gradients")
    args = ap.parse_args()
    if args.config == 5:
        global METRIC
        METRIC = "images/sec 8-stack HG bs16 256x256 train step"
        args.stacks, args.batch = 8, 16
    if args.config == 3 and args.impl == "ours":
        args.steps = 10 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        run_config3(args)
        return
    if args.impl == "reference":
        args.steps = 2 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
        return
    args.steps = 20 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else args.warmup
    if args.conv_path is not None:
        from pose_adv_aug_b200.models import asn_stacked_hg as M
        M.CONV_PATH = args.conv_path
    run_ours(args)


if __name__ == "__main__":
    main()

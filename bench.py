#!/usr/bin/env python
"""Benchmark of MINTIME's hot path on B200:  videos/sec for 16-frame 224x224 clips through
EfficientNet-B0 -> Size-Invariant TimeSformer (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the reference's CPU path (oracle port) on the host cores

A "step" is one forward of the hot path over one batch of synthetic clips (BASELINE.json configs[1]:
batch=32 16-frame 1-identity clips per GPU, inference).  `value` times the step with inputs resident
in HBM; `e2e` times the public nn.Module API fed from pinned HOST buffers (H2D of the clip and its
masks + D2H of the logits inside the timed region).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "videos_per_sec_16f_224px"
UNIT = "videos/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi polling during the timed region (B200_PROFILING.md 'clocks' recipe)."""

    def __init__(self, index: int, period_ms: int = 50):
        self.period_ms = period_ms
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.index = index
        self.window = "warm-up + timed region"

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.period_ms)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the upper half of the samples = clocks under load (the poller also sees idle gaps)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": self.window}


def cpu_oracle_videos_per_sec(batch: int, frames: int, identities, steps: int, warmup: int):
    """The reference's own CPU path on the host cores (fp32 eager, all threads): the UNMODIFIED reference modules
    staged in oracle/_ref (kind "reference") when they are there, else the oracle port (kind "port").
    Returns (videos/s, seconds per step, cores, kind)."""
    from mintime_b200 import synth
    from mintime_b200.spec import default_tsf_config
    from oracle import ref_arm

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = default_tsf_config(num_frames=frames)
    esd = synth.make_effnet_state_dict(1234)
    tsd = synth.make_tsf_state_dict(cfg, 4321)
    meta = synth.make_batch_meta(batch, frames, identities, seed=1234)
    clip = synth.make_frames(batch, frames, seed=1234, mask=meta["mask"])
    if ref_arm.available():
        kind = "reference"
        ext, model = ref_arm.build_modules(esd, tsd, cfg, require_attention=False)
        step = lambda: ref_arm.forward(ext, model, clip, meta)
    else:
        from oracle import mintime_oracle as orc
        kind = "port"
        step = lambda: orc.hot_path_forward(esd, tsd, cfg, clip, meta["mask"], meta["identities_mask"],
                                            meta["size_embedding"], meta["positions"])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            step()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec, cores, kind


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if getattr(args, "mode", "infer") == "train":
        # reference arm of `--mode train`: the oracle's training step (extractor forward + autograd + SGD) on the host
        from bench_train import cpu_train_videos_per_sec
        b = min(args.cpu_batch, 2)
        steps = max(1, min(args.steps, 3))
        v, sec, cores, kind = cpu_train_videos_per_sec(b, args.frames, args.identities, steps=steps)
        sample = (f"{b} clips x {args.frames} frames per training step, {steps} timed steps after 1 warm-up (bounded sample), "
                  + ("unmodified reference modules staged in oracle/_ref" if kind == "reference" else "oracle port"))
        emit({"impl": "reference", "metric": "train_videos_per_sec_16f_224px", "value": v, "unit": UNIT, "n_gpus": args.gpus,
              "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": "BASELINE.json configs[3]: train.py step (frozen extractor) on the host cores",
                         "batch_per_gpu": args.batch, "frames": args.frames, "precision": "f32"},
              "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    b = args.cpu_batch
    v, sec, cores, kind = cpu_oracle_videos_per_sec(b, args.frames, args.identities, args.steps, args.warmup)
    sample = (f"{b} clips x {args.frames} frames per step (bounded sample of the batch={args.batch} workload), "
              + ("unmodified reference modules staged in oracle/_ref" if kind == "reference" else "oracle port"))
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def north_star(kernels, scopes, peaks):
    """The two kernel targets BASELINE.json's north_star names, as measured in this run:
    the divided-attention kernel against the tensor peak, the MBConv stack against the HBM roofline on the strict
    algorithmic bytes of SURVEY.md 8(d)."""
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_attention.json")) as fh:
            ncu = json.load(fh)
    except Exception:
        pass
    attn = {}
    for k in kernels:
        if k["name"].startswith("fused_attn"):
            attn[k["name"]] = {"ms_per_launch": k["ms_per_launch"], "launches_per_step": k["launches_per_step"],
                               "tflops_algorithmic": k["tflops"], "frac_of_burst_peak": k["tflops"] / peaks["tensor_burst"],
                               "ncu_tensor_pipe_pct": ncu.get(k["name"], {}).get("sm__pipe_tensor_cycles_active_pct"),
                               "target_frac": 0.70}
    return {"divided_attention_kernel": attn or None,
            "mbconv_stack": dict(scopes.get("mbconv_stack", {}), target_frac=0.60) if "mbconv_stack" in scopes else None,
            "extractor": scopes.get("extractor")}


def workload_config(args, cpu=False):
    ids = ",".join(map(str, args.identities))
    return {
        "workload": f"BASELINE.json configs[1]: batch={args.batch} {args.frames}-frame {ids}-identity synthetic 224x224 "
                    f"clips per GPU, inference (EfficientNet-B0 eval -> SizeInvariantTimeSformer, channels=1280)",
        "batch_per_gpu": args.batch, "frames": args.frames, "identities": ids, "precision": "f32" if cpu else args.precision,
        "launch": "n/a" if cpu else ("eager nn.Module calls" if getattr(args, "no_graph", False)
                                     else "one CUDA-graph replay per step (mintime_b200.graphed.GraphedHotPath)"),
        "timing": "CUDA events on the launch stream, max over ranks; per-step input (308 MB fp32 clip batch) and "
                  "activations exceed the 126 MB L2" if not cpu else "wall clock, torch CPU eager, all host threads",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--identities", type=lambda s: [int(x) for x in s.split(",")], default=[1])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=4, help="clips per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attention-maps", action="store_true", help="also return the CLS attention maps (config 5)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: BASELINE configs[1] (default, the metric's config); train: configs[3], the train.py step")
    ap.add_argument("--unfrozen", action="store_true",
                    help="--mode train: the extractor trains too (train.py:155-170: .train(), batch-stat BN, drop-connect, "
                         "gradients) instead of --freeze_backbone")
    ap.add_argument("--unfreeze-blocks", type=int, default=-1,
                    help="--mode train --unfrozen: train only the last k MBConv blocks (train.py --extractor_unfreeze_blocks)")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-path measurement printed beside the bf16 number")
    ap.add_argument("--no-graph", action="store_true",
                    help="call the nn.Modules eagerly instead of replaying the step as a CUDA graph (GraphedHotPath)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 while we run (e.g. NCCL's
    # version banner) is routed to stderr, the real stdout is restored just before the JSON is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)

    if args.impl == "reference":
        run_reference(args, emit)
        return
    if args.mode == "train":
        from bench_train import run_train
        run_train(args, emit, ClockSampler, load_peaks)
        return

    import mintime_b200
    from mintime_b200 import _lib, synth
    from mintime_b200.spec import default_tsf_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback for the product path); "
                         "use --impl reference for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B, f = args.batch, args.frames
    cfg = default_tsf_config(num_frames=f)
    ext = mintime_b200.EfficientNet.from_name("efficientnet-b0", precision=args.precision)
    # conditioned extractor weights: the recipe of the end-to-end bf16 fixtures (tests/golden/cond_*.npz), so the run
    # that is timed can also be CHECKED against the reference (`parity` below); throughput does not depend on values
    esd = synth.make_effnet_state_dict(1234, conditioned=True)
    ext.load_state_dict(esd)
    ext = ext.to(dev).eval()
    model = mintime_b200.SizeInvariantTimeSformer(config=cfg, require_attention=args.attention_maps,
                                                  precision=args.precision)
    model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321))
    model = model.to(dev).eval()

    # every rank processes its own shard of clips (weak scaling: B clips per GPU, no data-path collective)
    meta = synth.make_batch_meta(B, f, args.identities, seed=1234 + rank)
    clip_u8 = synth.make_frames(B, f, seed=1234 + rank, mask=meta["mask"], dtype=torch.uint8)
    clip_dev = clip_u8.to(dev).float()                                   # resident fp32 NHWC, 0..255 (train.py:334)
    meta_dev = {k: v.to(dev) for k, v in meta.items()}
    host = {"clip": clip_u8.pin_memory(), **{k: v.pin_memory() for k, v in meta.items()}}
    logits_host = torch.empty((B, 1), dtype=torch.float32).pin_memory()

    def step_resident():
        with torch.no_grad():
            x = clip_dev.view(B * f, 224, 224, 3).permute(0, 3, 1, 2)    # train.py:341
            feats = ext(x)
            feats = feats.reshape(B, f, 1280, 7, 7)                      # train.py:354
            return model(feats, mask=meta_dev["mask"], size_embedding=meta_dev["size_embedding"],
                         identities_mask=meta_dev["identities_mask"], positions=meta_dev["positions"])

    # The step is launch-gap sensitive (~185 kernels of 20-150 us): the serving / benchmark entry point is
    # mintime_b200.graphed.GraphedHotPath, which captures extractor + model once and replays the CUDA graph
    # (8.43 -> 8.01 ms per step at B=32).  --no-graph times the eager nn.Module calls instead.
    use_graph = not args.no_graph
    graph_kernels = 0
    if use_graph:
        from mintime_b200.graphed import GraphedHotPath
        hot_res = GraphedHotPath(ext, model, B, f, frame_dtype=torch.float32, device=dev)
        hot_res.static["videos"].copy_(clip_dev)
        for k in ("mask", "identities_mask", "size_embedding", "positions"):
            hot_res.static[k].copy_(meta_dev[k])
        graph_kernels = hot_res.kernels_per_replay
        step_eager = step_resident

        def step_resident():                                             # noqa: F811  (timed variant)
            return hot_res.replay()

    # End-to-end loop = what a data loader + the public API do: every step's clip (uint8 NHWC) and masks go
    # pinned host -> HBM on a copy stream, double buffered so the copy of step i+1 overlaps the compute of
    # step i; every step ends with the D2H of its logits and a stream sync (the caller reads them).
    keys = ("clip", "mask", "identities_mask", "size_embedding", "positions")
    copy_stream = torch.cuda.Stream(device=dev)
    if use_graph:
        hots = [GraphedHotPath(ext, model, B, f, frame_dtype=torch.uint8, device=dev) for _ in range(2)]
        slots = [{"clip": h.static["videos"], **{k: h.static[k] for k in keys[1:]}} for h in hots]
    else:
        slots = [{k: torch.empty_like(host[k], device=dev) for k in keys} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"slot": 0, "primed": False}

    def issue_h2d(s):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[s])                              # the compute that last read this slot
            for k in keys:
                slots[s][k].copy_(host[k].view_as(slots[s][k]) if k == "clip" else host[k], non_blocking=True)
            ready[s].record(copy_stream)

    def step_e2e():
        if not state["primed"]:
            issue_h2d(0)
            state["primed"] = True
        s = state["slot"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[s])                                         # this step's inputs have landed
        issue_h2d(1 - s)                                                 # next step's inputs, overlapped
        with torch.no_grad():
            m = slots[s]
            if use_graph:
                out = hots[s].replay()
            else:
                x = m["clip"].view(B * f, 224, 224, 3).permute(0, 3, 1, 2)
                feats = ext(x).reshape(B, f, 1280, 7, 7)
                out = model(feats, mask=m["mask"], size_embedding=m["size_embedding"], identities_mask=m["identities_mask"],
                            positions=m["positions"])
            logits = out[0] if isinstance(out, tuple) else out
            done[s].record(cur)
            logits_host.copy_(logits, non_blocking=True)                 # D2H of the step's result
            cur.synchronize()                                            # the caller reads the logits
        state["slot"] = 1 - s
        return logits_host

    from mintime_b200 import dist as mdist

    def barrier():
        mdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        barrier()
        return mdist.max_over_ranks(ms, device=dev)     # device time, max over ranks

    lib = _lib.load()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # before the warm-up: nvidia-smi's own start-up (NVML init) stays out of the timed region
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    launches0 = lib.mt_prof_launch_count()
    t_wall = time.perf_counter()
    ms = timed(step_resident, args.steps, 0)
    launches = lib.mt_prof_launch_count() - launches0
    if use_graph:
        launches = graph_kernels * args.steps                            # kernels inside the replayed graph
    if rank == 0 and time.perf_counter() - t_wall < 0.4:
        # nvidia-smi cannot sample faster than ~20 ms: keep the SAME step running (untimed) until the poller has
        # seen at least ~0.4 s of load, so the clocks / throttle reasons describe this workload
        sampler.window = "warm-up + timed region + untimed continuation of the same step (region shorter than the poller needs)"
        while time.perf_counter() - t_wall < 0.4:
            step_resident()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, args.warmup)

    # ---- parity of the TIMED run: logits of the measured step (+ attention maps from one extra call) against what the
    # unmodified reference returned for clips 0/9/18/31 of this batch (tests/golden/cond_bench_b32_clips.npz)
    parity = None
    if rank == 0 and B == 32 and f == 16 and list(args.identities) == [1]:
        try:
            import numpy as np
            g = dict(np.load(os.path.join(ROOT, "tests", "golden", "cond_bench_b32_clips.npz")))
            idx = torch.tensor(g["clips"].tolist())
            out = step_resident()
            lg = (out[0] if isinstance(out, tuple) else out).float().cpu()
            model.require_attention = True
            with torch.no_grad():
                x = clip_dev.view(B * f, 224, 224, 3).permute(0, 3, 1, 2)
                _, (sa, ta) = model(ext(x).reshape(B, f, 1280, 7, 7), mask=meta_dev["mask"],
                                    size_embedding=meta_dev["size_embedding"],
                                    identities_mask=meta_dev["identities_mask"], positions=meta_dev["positions"])
            model.require_attention = args.attention_maps
            heads, N = cfg["model"]["heads"], 1 + f * 49
            sel = lambda m: m.float().cpu().view(B, heads, 1, N)[idx].reshape(-1).double()
            rel = lambda a, r: float((a - torch.from_numpy(r).reshape(-1).double()).norm() /
                                     torch.from_numpy(r).double().norm())
            dl = float((lg[idx] - torch.from_numpy(g["tsf.logits"])).abs().max())
            ds, dt = rel(sel(sa), g["tsf.space_attn"]), rel(sel(ta), g["tsf.time_attn"])
            tol = {"fp32": (1e-4, 1e-3), "bf16": (1e-2, 5e-3)}[args.precision]
            parity = {"against": "unmodified fp32 reference on clips 0/9/18/31 of this batch "
                                 "(tests/golden/cond_bench_b32_clips.npz)",
                      "max_abs_logit_err": dl, "space_attn_rel_l2": ds, "time_attn_rel_l2": dt,
                      "tolerance": {"max_abs_logit_err": tol[0], "attn_rel_l2": tol[1]},
                      "pass": bool(dl <= tol[0] and ds <= tol[1] and dt <= tol[1])}
        except Exception as e:                                           # never let the check hide the measurement
            parity = {"error": repr(e)}

    # ---- the exact (fp32, FFMA kernels) path on the same workload, beside the bf16 number
    fp32_path = None
    if rank == 0 and world == 1 and args.precision == "bf16" and not args.no_fp32:
        ext32 = mintime_b200.EfficientNet.from_name("efficientnet-b0", precision="fp32")
        ext32.load_state_dict(esd)
        ext32 = ext32.to(dev).eval()
        model32 = mintime_b200.SizeInvariantTimeSformer(config=cfg, precision="fp32")
        model32.load_state_dict(synth.make_tsf_state_dict(cfg, 4321))
        model32 = model32.to(dev).eval()

        def step32():
            with torch.no_grad():
                x = clip_dev.view(B * f, 224, 224, 3).permute(0, 3, 1, 2)
                return model32(ext32(x).reshape(B, f, 1280, 7, 7), mask=meta_dev["mask"],
                               size_embedding=meta_dev["size_embedding"], identities_mask=meta_dev["identities_mask"],
                               positions=meta_dev["positions"])
        ms32 = timed(step32, 3, 1)
        fp32_path = {"value": B / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32, "steps": 3,
                     "note": "MT_PREC_FP32: the exact path (FFMA kernels, eager module calls), same batch"}
        del ext32, model32

    # per-kernel CUDA-event times over 2 more steps of the same workload (mt_prof_* hooks in the library)
    lib.mt_prof_reset()
    lib.mt_prof_enable(1)
    prof_steps = 2
    for _ in range(prof_steps):
        (step_eager if use_graph else step_resident)()                   # (event-bracketed launches: eager calls)
    torch.cuda.synchronize()
    lib.mt_prof_enable(0)
    prof = _lib.profile_collect()
    lib.mt_prof_reset()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores, kind = cpu_oracle_videos_per_sec(args.cpu_batch, f, args.identities, 5, 1)
        what = ("the unmodified reference modules staged in oracle/_ref, fp32 torch-CPU eager" if kind == "reference"
                else "oracle = fp32 torch-CPU restatement of the reference")
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{args.cpu_batch} clips x {f} frames per step, 5 timed steps after 1 warm-up ({what}), "
                         f"{sec:.2f} s/step"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    peaks = load_peaks()
    ridge = peaks["tensor"] * 1e12 / (peaks["hbm"] * 1e9)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            traffic = json.load(fh)
    except Exception:
        pass
    kernels = []
    scopes = {}
    total_ms = sum(p[1] for p in prof if "(strict bytes)" not in p[0]) or 1.0
    for name, ms_t, fl, by, cnt in prof:
        sec = ms_t * 1e-3
        if "(strict bytes)" in name:              # nested scopes over whole groups of launches, not kernels
            scopes[name.split(" ")[0]] = {"bytes_strict": by / cnt, "ms": ms_t / cnt, "gbs": by / sec / 1e9,
                                          "frac": by / sec / 1e9 / peaks["hbm"], "peak_hbm_gbs": peaks["hbm"]}
            continue
        bound = "tensor" if (by > 0 and fl / by >= ridge and name.startswith(("gemm", "fused_attn"))) else "hbm"
        ach = fl / sec / 1e12 if bound == "tensor" else by / sec / 1e9
        # every entry is ONE event-bracketed launch at full clocks: the burst figure is the tensor denominator
        peak = peaks["tensor_burst"] if bound == "tensor" else peaks["hbm"]
        kernels.append({"name": name, "share": ms_t / total_ms, "ms_per_launch": ms_t / cnt, "launches_per_step": cnt / prof_steps,
                        "bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                        "frac": ach / peak, "tflops": fl / sec / 1e12, "gbs": by / sec / 1e9})
    dom = kernels[0] if kernels else None
    roofline = None
    if dom:
        t = traffic.get(dom["name"])
        roofline = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": t, "kernel": dom["name"], "share_of_step": dom["share"],
                    "peak_source": f"{peaks['source']} ({'burst bf16 cuBLAS' if dom['bound'] == 'tensor' else 'copy'} "
                                   f"figure of MEASURED_PEAKS.json: the kernel is timed alone, event-bracketed)",
                    "ms_per_launch": dom["ms_per_launch"]}
    n = world
    h2d = int(host["clip"].numel() * host["clip"].element_size() +
              sum(host[k].numel() * host[k].element_size() for k in ("mask", "identities_mask", "size_embedding", "positions")))
    out = {
        "metric": METRIC, "value": n * B / (ms * 1e-3), "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": n * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(logits_host.numel() * 4), "ms_per_step": ms_e2e,
                "input": "uint8 NHWC clips + masks/positions from pinned host memory, H2D of step i+1 on a copy "
                         "stream overlapping the compute of step i; logits D2H + stream sync every step"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "north_star": north_star(kernels, scopes, peaks),
        "parity": parity,
        "fp32_path": fp32_path,
        "cpu_baseline": cpu,
        "kernels": kernels,
        "sum_kernel_ms_per_step": total_ms / prof_steps,
    }
    emit(out)


if __name__ == "__main__":
    main()

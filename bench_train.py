"""`python bench.py --mode train` -- BASELINE.json configs[3]: the training step of train.py:332-378 (frozen extractor:
--freeze_backbone) on synthetic ForgeryNet-shaped clips, bf16 compute / fp32 master weights, SGD(lr 0.01, wd 1e-4),
BCEWithLogits(pos_weight); one process per GPU, gradients averaged by the per-layer all-reduce issued from inside the
backward (mintime_b200.training.GradSync, NCCL).  Same JSON contract as the inference line of bench.py; a step =
H2D of the uint8 clips + extractor forward + transformer forward/backward + optimizer step + loss D2H."""
import json
import os
import time

import torch


def cpu_train_videos_per_sec(batch, frames, identities, steps=2):
    """The reference's training step (train.py:332-378 with --freeze_backbone: extractor forward under no_grad + autograd
    through the transformer + SGD) on the host cores: the UNMODIFIED reference modules staged in oracle/_ref when they are
    there (kind "reference"), else the oracle port.  Returns (videos/s, s/step, threads, kind)."""
    from mintime_b200 import synth
    from mintime_b200.spec import default_tsf_config
    from oracle import mintime_oracle as orc
    from oracle import ref_arm
    torch.set_num_threads(os.cpu_count())
    cfg = default_tsf_config(num_frames=frames)
    esd = synth.make_effnet_state_dict(1234, conditioned=True)
    meta = synth.make_batch_meta(batch, frames, identities, seed=1234)
    clip = synth.make_frames(batch, frames, seed=1234, mask=meta["mask"])
    labels = (torch.arange(batch) % 2).float().view(batch, 1)
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([0.8169]))
    if ref_arm.available():
        from einops import rearrange
        ext, model = ref_arm.build_modules(esd, synth.make_tsf_state_dict(cfg, 4321), cfg, require_attention=False)
        model.train()
        ropt = torch.optim.SGD(model.parameters(), lr=0.01, weight_decay=1e-4)

        def rstep():
            with torch.no_grad():                                       # train.py:344-346
                feats = ext(rearrange(clip, "b f h w c -> (b f) c h w"))
            feats = rearrange(feats, "(b f) c h w -> b f c h w", b=batch)
            y = model(feats, mask=meta["mask"], size_embedding=meta["size_embedding"], identities_mask=meta["identities_mask"],
                      positions=meta["positions"])
            ropt.zero_grad()
            lossf(y, labels).backward()
            ropt.step()

        rstep()
        t0 = time.perf_counter()
        for _ in range(steps):
            rstep()
        sec = (time.perf_counter() - t0) / steps
        return batch / sec, sec, torch.get_num_threads(), "reference"
    tsd = {k: v.clone().requires_grad_(True) for k, v in synth.make_tsf_state_dict(cfg, 4321).items()}
    opt = torch.optim.SGD(list(tsd.values()), lr=0.01, weight_decay=1e-4)

    def step():
        with torch.no_grad():
            feats = orc.effnet_b0_forward(esd, clip.view(batch * frames, 224, 224, 3).permute(0, 3, 1, 2))
        feats = feats.view(batch, frames, *feats.shape[1:])
        opt.zero_grad()
        logits, _ = orc.tsf_forward(tsd, cfg, feats, meta["mask"], meta["identities_mask"], meta["size_embedding"],
                                    meta["positions"])
        lossf(logits, labels).backward()
        opt.step()

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    sec = (time.perf_counter() - t0) / steps
    return batch / sec, sec, torch.get_num_threads(), "port"


class NvmlSampler:
    """Clocks / throttle reasons DURING the timed region from an in-process NVML thread.  (bench.py's nvidia-smi poller
    re-initialises NVML on every sample; its driver locks stall the ~900 host-side launches of an eager training step:
    25.7 -> 34 ms at a 50 ms period, 39 ms on the polled GPU of a 2-GPU run.)"""

    def __init__(self, index: int, period_s: float = 0.05):
        import threading
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.window = "warm-up + timed region"
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self.ok = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self._thread.start()
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self._thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        s = sorted(self.samples)
        load = s[len(s) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "window": self.window, "source": "in-process NVML thread, 50 ms period"}


def run_train(args, emit, ClockSampler, load_peaks):
    import mintime_b200
    from mintime_b200 import _lib, synth, training
    from mintime_b200 import dist as mdist
    from mintime_b200.spec import default_tsf_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --mode train needs a GPU (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B, f = args.batch, args.frames
    cfg = default_tsf_config(num_frames=f)
    ext = mintime_b200.EfficientNet.from_name("efficientnet-b0", precision=args.precision)
    ext.load_state_dict(synth.make_effnet_state_dict(1234, conditioned=True))
    ext = ext.to(dev).eval()                                            # train.py:153-154 (freeze_backbone)
    unfrozen = bool(getattr(args, "unfrozen", False))
    if unfrozen:                                                        # train.py:155-170
        ext.train()
        k = getattr(args, "unfreeze_blocks", -1)
        if k >= 0:
            for name, p in ext.named_parameters():                      # train.py:157-167
                if name.startswith("_blocks."):
                    p.requires_grad_(int(name.split(".")[1]) >= 16 - k)
                else:
                    p.requires_grad_(name.startswith(("_conv_head", "_bn1")))
        if world > 1:
            training.attach_grad_sync(ext)                              # one averaged bucket for the extractor's gradients
    model = mintime_b200.SizeInvariantTimeSformer(config=cfg, precision=args.precision)
    model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321))
    model = model.to(dev).train()                                       # train.py:315
    if world > 1:
        training.attach_grad_sync(model)
    train_params = list(model.parameters()) + ([p for p in ext.parameters() if p.requires_grad] if unfrozen else [])
    opt = torch.optim.SGD(train_params, lr=cfg["training"]["lr"], weight_decay=cfg["training"]["weight-decay"])
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([0.8169], device=dev))

    meta = synth.make_batch_meta(B, f, args.identities, seed=1234 + rank)
    host = {"clip": synth.make_frames(B, f, seed=1234 + rank, mask=meta["mask"], dtype=torch.uint8).pin_memory(),
            "labels": (torch.rand((B, 1), generator=torch.Generator().manual_seed(rank)) < 0.55).float().pin_memory(),
            **{k: v.pin_memory() for k, v in meta.items()}}
    resident = {k: v.to(dev) for k, v in host.items()}
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def train_step(m):
        if unfrozen:                                                    # train.py:347-348
            feats = ext(m["clip"].view(B * f, 224, 224, 3).permute(0, 3, 1, 2)).reshape(B, f, 1280, 7, 7)
        else:
            with torch.no_grad():                                       # train.py:344-346
                feats = ext(m["clip"].view(B * f, 224, 224, 3).permute(0, 3, 1, 2)).reshape(B, f, 1280, 7, 7)
        opt.zero_grad(set_to_none=True)
        y = model(feats, mask=m["mask"], size_embedding=m["size_embedding"], identities_mask=m["identities_mask"],
                  positions=m["positions"])
        loss = lossf(y, m["labels"])
        loss.backward()
        opt.step()
        return loss

    def step_eager():
        return train_step(resident)

    def step_resident():
        return train_step(resident)

    def step_e2e():
        m = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss_host.copy_(train_step(m).detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return loss_host

    # The step (extractor, forward, backward incl. the per-layer NCCL all-reduces under data parallelism) replays as ONE
    # CUDA graph (mintime_b200.graphed.GraphedTrainStep) followed by the eager optimizer step; the fully eager step is
    # bound by its ~900 host-side launches.
    use_graph = not args.no_graph
    graph_kernels = 0
    if use_graph:
        from mintime_b200.graphed import GraphedTrainStep
        gs = GraphedTrainStep(ext, model, opt, lossf, B, f, frame_dtype=torch.uint8, device=dev)
        names = {"clip": "videos"}
        for k, v in resident.items():
            gs.static[names.get(k, k)].copy_(v.view_as(gs.static[names.get(k, k)]))
        gs.capture()
        graph_kernels = gs.kernels_per_replay

        def step_resident():                                            # noqa: F811
            return gs.replay()

        def step_e2e():                                                 # noqa: F811
            for k, v in host.items():
                dst = gs.static[names.get(k, k)]
                dst.copy_(v.view_as(dst), non_blocking=True)
            loss_host.copy_(gs.replay(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return loss_host

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
        return mdist.max_over_ranks(ms, device=dev)

    lib = _lib.load()
    sampler = NvmlSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (the first steps of a data-parallel run still set up NCCL channels and grow the side-stream allocator pools)
    n_warm = 1 if unfrozen else max(args.warmup, 8)
    if unfrozen:
        args.steps = min(args.steps, 3)
        args.warmup = 1
    for _ in range(n_warm):
        step_resident()
    torch.cuda.synchronize()
    launches0 = lib.mt_prof_launch_count()
    t_wall = time.perf_counter()
    ms = timed(step_resident, args.steps, 0)
    launches = lib.mt_prof_launch_count() - launches0
    if use_graph:
        launches = graph_kernels * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    lib.mt_prof_reset()
    lib.mt_prof_enable(1)
    model._serial_wgrad = True                                          # per-kernel times without side-stream overlap
    step_eager()                                                        # (event-bracketed launches: eager calls)
    torch.cuda.synchronize()
    model._serial_wgrad = False
    lib.mt_prof_enable(0)
    prof = _lib.profile_collect()
    lib.mt_prof_reset()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores, kind = cpu_train_videos_per_sec(2, f, args.identities)
        cpu = {"value": v, "unit": "videos/s", "cores": cores, "kind": kind,
               "sample": f"2 clips x {f} frames per training step ("
                         + ("unmodified reference modules staged in oracle/_ref" if kind == "reference" else "oracle port")
                         + f": extractor forward under no_grad + torch autograd + SGD), 2 timed steps after 1 warm-up, {sec:.2f} s/step"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = load_peaks()
    ridge = peaks["tensor"] * 1e12 / (peaks["hbm"] * 1e9)
    prof = [p for p in prof if "(strict bytes)" not in p[0]]          # nested scopes over groups of launches, not kernels
    total_ms = sum(p[1] for p in prof) or 1.0
    kernels = []
    for name, ms_t, fl, by, cnt in prof:
        sec = ms_t * 1e-3
        bound = "tensor" if (by > 0 and fl / by >= ridge and name.startswith("gemm")) else "hbm"
        ach = fl / sec / 1e12 if bound == "tensor" else by / sec / 1e9
        peak = peaks["tensor"] if bound == "tensor" else peaks["hbm"]
        kernels.append({"name": name, "share": ms_t / total_ms, "ms_per_launch": ms_t / cnt, "launches_per_step": cnt,
                        "bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                        "frac": ach / peak})
    dom = kernels[0] if kernels else None
    roofline = None
    traffic = {}
    try:     # per-launch dram bytes (ncu --set full, profiles/r1_train_kernels_ncu_summary.txt)
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic_train.json")) as fh:
            traffic = json.load(fh)
    except Exception:
        pass
    if dom:
        roofline = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": traffic.get(dom["name"]), "kernel": dom["name"], "share_of_step": dom["share"],
                    "ms_per_launch": dom["ms_per_launch"], "peak_source": peaks["source"]}
    h2d = int(sum(v.numel() * v.element_size() for v in host.values()))
    emit({
        "metric": "train_videos_per_sec_16f_224px", "value": world * B / (ms * 1e-3), "unit": "videos/s", "n_gpus": world,
        "steps": args.steps, "warmup": n_warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": ("BASELINE.json configs[3]: train.py step, TRAINABLE EfficientNet-B0 (.train(): batch-stat BN, "
                                "drop-connect, fp32 forward + backward, csrc/effnet_train.cu"
                                + (f", last {args.unfreeze_blocks} blocks unfrozen" if getattr(args, "unfreeze_blocks", -1) >= 0 else "")
                                + ") -> SizeInvariantTimeSformer forward + backward + SGD over both modules"
                                if unfrozen else
                                "BASELINE.json configs[3]: train.py step, frozen EfficientNet-B0 (eval, no_grad) -> "
                                "SizeInvariantTimeSformer forward + backward + SGD") + ", synthetic ForgeryNet-shaped clips",
                   "batch_per_gpu": B, "frames": f, "identities": ",".join(map(str, args.identities)),
                   "precision": args.precision + " compute, fp32 master weights and gradients",
                   "launch": ("one CUDA-graph replay per step (GraphedTrainStep: extractor, forward, backward) + eager " + ("NCCL gradient exchange + " if world > 1 else "") + "optimizer.step()") if use_graph
                             else "eager nn.Module / autograd calls",
                   "grad_exchange": ("per-layer flat fp32 buckets (9 x 29 MB + 1), NCCL all-reduce + average after the graph replay"
                                     if use_graph else "per-layer flat fp32 buckets, NCCL all-reduce issued inside the backward")
                                    if world > 1 else "none (1 GPU)",
                   "timing": "CUDA events on the launch stream, max over ranks; activations per step (6.7 GB) exceed the L2"},
        "clocks": clocks,
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "videos/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e, "input": "uint8 NHWC clips, masks, positions, labels from pinned host memory every "
                                                "step; loss D2H + stream sync every step"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels[:24],
        "sum_kernel_ms_per_step": total_ms,
    })

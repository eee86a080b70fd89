"""CUDA-graph replay of the hot path for a fixed batch shape.

The reference serves one video at a time (predict.py:394-417: b = 1, 8..32 faces).  At that size the ~185 kernel
launches of EfficientNet-B0 -> SizeInvariantTimeSformer cost more host time (launch + TMA descriptor encoding) than
GPU time, so the latency-critical entry point captures them ONCE into a CUDA graph -- possible because nothing behind
the C ABI allocates or synchronises -- and replays the graph per clip.  Results are bit-identical to the eager
modules (tests/test_gpu_parity.py::test_graphed_hot_path).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib


class GraphedHotPath:
    """extractor + model captured for (batch, num_frames) clips of 224x224 faces.

    ``__call__(videos, mask, identities_mask, size_embedding, positions)`` takes the tensors of the reference loop
    (videos (b,f,224,224,3) raw 0..255, uint8 or float32, host or device) and returns ``logits`` or
    ``(logits, [space_attn, time_attn])`` like ``SizeInvariantTimeSformer.forward``; the returned tensors are the
    graph's static outputs and are overwritten by the next call.
    """

    def __init__(self, extractor, model, batch: int, num_frames: int, frame_dtype=torch.uint8, device="cuda:0",
                 num_patches: int = 49):
        self.device = torch.device(device)
        _lib.require_device(self.device)
        self.ext, self.model = extractor, model
        self.b, self.f = batch, num_frames
        d = self.device
        self.static: Dict[str, torch.Tensor] = {
            "videos": torch.zeros((batch, num_frames, 224, 224, 3), dtype=frame_dtype, device=d),
            "mask": torch.ones((batch, num_frames), dtype=torch.bool, device=d),
            "identities_mask": torch.ones((batch, num_frames, num_frames), dtype=torch.bool, device=d),
            "size_embedding": torch.ones((batch, num_frames), dtype=torch.int32, device=d),
            "positions": torch.arange(1 + num_frames * num_patches, dtype=torch.int64, device=d).repeat(batch, 1),
        }
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out = None
        self.kernels_per_replay = 0      # kernels of this library inside the graph
        self._capture()

    def _forward(self):
        s = self.static
        x = s["videos"].view(self.b * self.f, 224, 224, 3).permute(0, 3, 1, 2)          # train.py:341
        feats = self.ext(x)
        feats = feats.reshape(self.b, self.f, *feats.shape[1:])                           # train.py:354
        return self.model(feats, mask=s["mask"], size_embedding=s["size_embedding"],
                          identities_mask=s["identities_mask"], positions=s["positions"])

    def _capture(self):
        with torch.no_grad(), torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                  # warm-up: weight packing, one-time function attributes
                    self._forward()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.load().mt_prof_launch_count()
            with torch.cuda.graph(self.graph):
                self.out = self._forward()
            self.kernels_per_replay = int(_lib.load().mt_prof_launch_count() - n0)

    def replay(self):
        """Replay the captured graph on whatever ``self.static`` holds (pipelined feeding: fill ``static`` from a
        copy stream, make the current stream wait for it, then call this); returns the static outputs."""
        self.graph.replay()
        return self.out

    def __call__(self, videos: torch.Tensor, mask: torch.Tensor, identities_mask: torch.Tensor,
                 size_embedding: torch.Tensor, positions: torch.Tensor):
        s = self.static
        if tuple(videos.shape) != tuple(s["videos"].shape):
            raise ValueError(f"graph captured for {tuple(s['videos'].shape)}, got {tuple(videos.shape)}")
        s["videos"].copy_(videos, non_blocking=True)
        s["mask"].copy_(mask, non_blocking=True)
        s["identities_mask"].copy_(identities_mask, non_blocking=True)
        s["size_embedding"].copy_(size_embedding, non_blocking=True)
        s["positions"].copy_(positions, non_blocking=True)
        self.graph.replay()
        return self.out


class GraphedTrainStep:
    """One training step of train.py:332-378 (frozen extractor forward, model forward, loss, ``loss.backward()``,
    ``optimizer.step()``) captured as ONE CUDA graph for a fixed batch shape.

    The eager step issues ~500 library launches plus ~400 small torch kernels from Python and is bound by that host
    work; the replay is bound by the GPU.  Everything the step does is capturable: the library never allocates or
    synchronises, the autograd node of ``training.py`` forks its weight-gradient GEMMs to side streams that join back
    before it returns, and the bf16 re-packing of the updated parameters is part of the captured forward.  Under data
    parallelism (``training.attach_grad_sync``) the backward's gradient buckets are static tensors of the graph: the
    replay is followed by ONE eager round of NCCL all-reduces over them (``GradSync.exchange``: 9 per-layer buckets of
    29 MB + one small one) and the optimizer step.  (Capturing the collectives themselves was tried in round 2 and hung
    at 2 GPUs inside the capture; the exchange is ~1 % of the step over NVLink, so it is issued after the replay.)
    Fill ``static`` (videos, mask, identities_mask, size_embedding, positions, labels), call ``replay()``; ``loss`` is the
    step's static output.

    ``capture_optimizer`` (default False): ``optimizer.step()`` runs EAGERLY after every replay on the graph's static
    gradient tensors, so learning-rate schedulers (train.py:380 calls ``lr_scheduler.step_update`` every iteration;
    the shipped config uses a cosine schedule) and any optimizer work unchanged.  With True the update is captured too,
    which freezes every host-side hyper-parameter (lr, weight decay) at its capture-time value: only for constant-lr SGD.
    """

    def __init__(self, extractor, model, optimizer, loss_fn, batch: int, num_frames: int, frame_dtype=torch.uint8,
                 device="cuda:0", num_patches: int = 49, warmup: int = 3, capture_optimizer: bool = False):
        self.device = torch.device(device)
        _lib.require_device(self.device)
        self.capture_optimizer = capture_optimizer
        if capture_optimizer and getattr(model, "_grad_sync", None) is not None:
            raise ValueError("capture_optimizer=True cannot be combined with a gradient exchange (it runs after the replay)")
        self.ext, self.model, self.opt, self.loss_fn = extractor, model, optimizer, loss_fn
        self.b, self.f = batch, num_frames
        d = self.device
        self.static: Dict[str, torch.Tensor] = {
            "videos": torch.zeros((batch, num_frames, 224, 224, 3), dtype=frame_dtype, device=d),
            "mask": torch.ones((batch, num_frames), dtype=torch.bool, device=d),
            "identities_mask": torch.ones((batch, num_frames, num_frames), dtype=torch.bool, device=d),
            "size_embedding": torch.ones((batch, num_frames), dtype=torch.int32, device=d),
            "positions": torch.arange(1 + num_frames * num_patches, dtype=torch.int64, device=d).repeat(batch, 1),
            "labels": torch.zeros((batch, model.num_classes), dtype=torch.float32, device=d),
        }
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss = None
        self.logits = None
        self.kernels_per_replay = 0
        self._warmup = warmup

    def _step(self):
        s = self.static
        sync = getattr(self.model, "_grad_sync", None)
        if sync is not None:
            sync.deferred = self.graph is not None             # inside the capture: record the buckets, exchange later
            sync.buckets = []
        # frozen extractor (eval, train.py:344-346): no graph of its own; unfrozen (train.py:347-348): its train-mode
        # autograd node (batch-statistic BatchNorm, drop-connect draws from the graph-safe CUDA generator) is captured too
        with torch.set_grad_enabled(bool(self.ext.training)):
            x = s["videos"].view(self.b * self.f, 224, 224, 3).permute(0, 3, 1, 2)
            feats = self.ext(x)
            feats = feats.reshape(self.b, self.f, *feats.shape[1:])
        esync = getattr(self.ext, "_grad_sync", None)
        if esync is not None:
            esync.deferred = self.graph is not None
            esync.buckets = []
        self.model._train_pack = None          # the re-pack of the (just updated) parameters belongs to every step
        y = self.model(feats, mask=s["mask"], size_embedding=s["size_embedding"], identities_mask=s["identities_mask"],
                       positions=s["positions"])
        if isinstance(y, tuple):
            y = y[0]
        loss = self.loss_fn(y, s["labels"])
        loss.backward()
        if self.capture_optimizer or self.graph is None:       # (the eager warm-up steps always update)
            self.opt.step()
        return loss.detach(), y.detach()

    def capture(self):
        """Run `warmup` eager steps on the CURRENT contents of ``static`` (they do update the parameters), then
        capture the step."""
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self._warmup):
                    self.opt.zero_grad(set_to_none=True)
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            n0 = _lib.load().mt_prof_launch_count()
            # MINTIME_B200_TRAIN_PRIO=1: capture on a high-priority stream, so the kernel nodes of the critical path
            # outrank the weight-gradient work forked to the (default-priority) side streams
            import os
            cap = torch.cuda.Stream(priority=-1) if os.environ.get("MINTIME_B200_TRAIN_PRIO", "0") == "1" else None
            self.graph = graph                                 # (set before the capture: _step skips the eager update)
            try:
                with torch.cuda.graph(graph, stream=cap):
                    self.loss, self.logits = self._step()
            except Exception:
                self.graph = None
                raise
            self.kernels_per_replay = int(_lib.load().mt_prof_launch_count() - n0)
        return self

    def replay(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        for owner in (self.model, self.ext):
            sync = getattr(owner, "_grad_sync", None)
            if sync is not None:
                sync.exchange()        # NCCL all-reduce + average of the graph's static gradient buckets
        if not self.capture_optimizer:
            self.opt.step()            # eager, on the graph's static .grad tensors: schedulers / any optimizer work
        return self.loss

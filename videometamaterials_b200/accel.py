"""A minimal `accelerate.Accelerator`-compatible object over torch.distributed (NCCL on GPUs, gloo on CPU).

The reference's Trainer talks to HuggingFace Accelerate (not installed here) through
exactly these members (`grep accelerator\\.` over VDDP:1400-1919): init_trackers, log, prepare, device, print,
num_processes, process_index, is_main_process, wait_for_everyone, accumulate, backward, sync_gradients,
clip_grad_norm_, gather, gather_for_metrics, pad_across_processes, unwrap_model, end_training - plus
accelerate.utils.broadcast_object_list (VDDP:24).  This class provides them with the same meaning.

Data parallelism ("Accelerate's role" in the north star): one process per GPU, replicated parameters, the batch
sharded over ranks, and ONE all-reduce per step over the flat fp32 gradient arena (blocks_bwd.GradArena) on NCCL -
the reference gets the same effect from DDP's bucketed all-reduce over Gloo (main.py:31-34).  The 31 parameters
that never receive a gradient just carry zeros in the arena, so no unused-parameter search is needed.
"""
from __future__ import annotations

import contextlib
import datetime
import os
from typing import Any, List, Optional

import torch
import torch.distributed as dist
from torch.utils import data as tdata


class _DeviceLoader:
    """DataLoader wrapper that moves batches to the device (what accelerate's prepared loaders do)."""

    def __init__(self, loader, device, sampler=None, total=None):
        self.loader, self.device = loader, device
        self.sampler = sampler        # DistributedSampler of a multi-process run (None otherwise)
        self.total = total            # dataset length: gather_for_metrics trims the sampler's padding duplicates with it
        self.epoch = 0
        self.remainder = -1           # samples of the LAST gathered batch that are not padding (-1: no trimming needed)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        if self.sampler is not None:
            # a new permutation per pass, as accelerate's prepared loader reshuffles every epoch (the sampler seeds with seed + epoch)
            self.sampler.set_epoch(self.epoch)
            self.epoch += 1
        n_batches = len(self.loader)
        # On a CUDA device the copy of batch i + 1 is issued on a copy stream when batch i is handed out, so that it runs under the
        # step that consumes batch i (the compute stream only waits for the copy's event: no host synchronisation).
        side = self.device.type == "cuda" and torch.cuda.is_available()
        if side and getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        pending = None
        for i, batch in enumerate(self.loader):
            # accelerate's GradientState: on the last batch of a sharded loader `remainder` = the number of genuine samples in the
            # gathered global batch (the DistributedSampler pads every rank's shard to equal length with repeated samples)
            rem = -1
            if self.sampler is not None and i == n_batches - 1 and self.total is not None and self.loader.batch_size:
                rem = self.total % (self.loader.batch_size * self.sampler.num_replicas)
            if side:
                with torch.cuda.stream(self._copy_stream):
                    dev_batch = self._to_device(batch)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
            else:
                dev_batch, ev = self._to_device(batch), None
            if pending is not None:
                yield self._hand_out(*pending)
            pending = (dev_batch, ev, rem)
        if pending is not None:
            yield self._hand_out(*pending)
        self.remainder = -1

    def _hand_out(self, batch, ev, rem):
        self.remainder = rem
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in (batch if isinstance(batch, (list, tuple)) else (batch,)):
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(cur)          # allocated on the copy stream, consumed on the compute stream
        return batch

    def _to_device(self, batch):
        if isinstance(batch, (list, tuple)):
            return type(batch)(b.to(self.device, non_blocking=True) if torch.is_tensor(b) else b for b in batch)
        if torch.is_tensor(batch):
            return batch.to(self.device, non_blocking=True)
        return batch


class Accelerator:
    def __init__(self, mixed_precision: str = "bf16", kwargs_handlers=None, log_with=None, cpu: bool = False, backend: Optional[str] = None):
        self.mixed_precision = mixed_precision
        self.log_with = log_with
        self.sync_gradients = True
        self.logs: List[dict] = []
        world = int(os.environ.get("WORLD_SIZE", "1"))
        self.use_cuda = torch.cuda.is_available() and not cpu
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.use_cuda:
            torch.cuda.set_device(local_rank % max(torch.cuda.device_count(), 1))
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device("cpu")
        self.grad_group = None
        if world > 1:
            want = backend or ("nccl" if self.use_cuda else "gloo")
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29500")
                dist.init_process_group(backend=want, timeout=datetime.timedelta(seconds=600))
            elif dist.get_backend() != want:
                # reference main.py creates a gloo group first (main.py:31); gradients still go over NCCL
                self.grad_group = dist.new_group(backend=want)
        self.num_processes = dist.get_world_size() if dist.is_initialized() else 1
        self.process_index = dist.get_rank() if dist.is_initialized() else 0
        self._models: List[torch.nn.Module] = []
        self._loaders: List[_DeviceLoader] = []

    # ------------------------------------------------------------------ properties
    @property
    def is_main_process(self) -> bool:
        return self.process_index == 0

    @property
    def compute_dtype(self) -> torch.dtype:
        """16-bit activation format of the training kernels.  The reference asks accelerate for 'fp16' (main.py:34: autocast plus
        a GradScaler); the B200 path stores the same 16 bits as bf16, which needs no loss scale, so 'fp16' and 'bf16' both train
        in bf16 here.  Sampling runs the EMA copy in fp16 (three more mantissa bits, no gradients involved), see Trainer."""
        return torch.bfloat16

    # ------------------------------------------------------------------ logging
    def init_trackers(self, project_name=None, init_kwargs=None, **kw):
        self.project_name = project_name

    def log(self, values: dict, step: Optional[int] = None):
        if self.is_main_process:
            self.logs.append(dict(values, step=step))

    def end_training(self):
        pass

    def print(self, *a, **k):
        if self.is_main_process:
            print(*a, **k)

    # ------------------------------------------------------------------ preparation
    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                inner = getattr(o, "denoise_fn", o)
                if hasattr(inner, "set_compute_dtype"):
                    inner.set_compute_dtype(self.compute_dtype)
                if self.num_processes > 1:
                    self.broadcast_parameters(o)
                self._models.append(o)
            elif isinstance(o, tdata.DataLoader):
                sampler = None
                if self.num_processes > 1:
                    shuffle = isinstance(o.sampler, tdata.RandomSampler)
                    sampler = tdata.distributed.DistributedSampler(o.dataset, num_replicas=self.num_processes, rank=self.process_index,
                                                                   shuffle=shuffle)
                    extra = dict(collate_fn=o.collate_fn, worker_init_fn=o.worker_init_fn, timeout=o.timeout, generator=o.generator)
                    if o.num_workers > 0:
                        extra.update(persistent_workers=o.persistent_workers, prefetch_factor=o.prefetch_factor)
                    o = tdata.DataLoader(o.dataset, batch_size=o.batch_size, sampler=sampler, pin_memory=o.pin_memory,
                                         num_workers=o.num_workers, drop_last=o.drop_last, **extra)
                o = _DeviceLoader(o, self.device, sampler=sampler, total=len(o.dataset))
                self._loaders.append(o)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model):
        return model

    def broadcast_parameters(self, module: torch.nn.Module) -> None:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=self.grad_group)

    @contextlib.contextmanager
    def accumulate(self, model):
        yield

    # ------------------------------------------------------------------ backward + gradient all-reduce
    def backward(self, loss: torch.Tensor):
        loss.backward()
        self.all_reduce_gradients()

    def all_reduce_gradients(self, average: bool = True) -> float:
        """SUM the gradient arenas over ranks and divide by the world size (what DDP's averaging does).  With average=False the
        division is left to the caller (the fused optimiser kernel takes it as `grad_scale`, which saves one pass over the
        158 MB arena); the factor still to be applied is returned."""
        if self.num_processes <= 1:
            return 1.0
        pending = 1.0
        for m in self._models:
            inner = getattr(m, "denoise_fn", m)
            arena = getattr(inner, "_vmm_arena", None)
            if arena is not None:
                dist.all_reduce(arena.flat_grad, op=dist.ReduceOp.SUM, group=self.grad_group)
                if average:
                    arena.flat_grad.mul_(1.0 / self.num_processes)
                else:
                    pending = 1.0 / self.num_processes
            else:
                for p in m.parameters():
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.grad_group)
                        p.grad.mul_(1.0 / self.num_processes)
        return pending

    def reduce_gradient_chunks(self, arena, n_chunks: int = 4):
        """SUM all-reduce of one gradient arena as `n_chunks` asynchronous collectives over contiguous pieces.  Returns
        ([(start, end, work), ...], factor still to be applied = 1 / world size).  The caller waits on piece i and runs the optimiser
        over it while piece i + 1 is still on the wire (Trainer.train_step), so that only the first piece's transfer and the last
        piece's update are exposed instead of the whole all-reduce followed by the whole update."""
        n = arena.flat_grad.numel()
        step = -(-n // max(1, n_chunks))
        step = -(-step // 1024) * 1024                       # 4 KB boundaries keep every piece 16-byte aligned
        out = []
        for a in range(0, n, step):
            b = min(n, a + step)
            out.append((a, b, dist.all_reduce(arena.flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.grad_group, async_op=True)))
        return out, 1.0 / self.num_processes

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)

    # ------------------------------------------------------------------ collectives
    def wait_for_everyone(self):
        if self.num_processes > 1:
            dist.barrier()

    def gather(self, tensor: torch.Tensor) -> torch.Tensor:
        if self.num_processes <= 1:
            return tensor
        t = tensor.reshape(1) if tensor.dim() == 0 else tensor.contiguous()
        outs = [torch.empty_like(t) for _ in range(self.num_processes)]
        dist.all_gather(outs, t)
        return torch.cat(outs, dim=0)

    def gather_for_metrics(self, tensor: torch.Tensor) -> torch.Tensor:
        """gather() minus the samples a sharded loader duplicated to fill its last global batch (accelerate truncates the gathered
        tensor to `remainder` entries on that batch; a per-rank scalar counts as one entry, as in accelerate)."""
        out = self.gather(tensor)
        if self.num_processes > 1:
            for ld in self._loaders:
                if ld.remainder > 0 and out.dim() > 0 and out.shape[0] > ld.remainder:
                    return out[:ld.remainder]
        return out

    def pad_across_processes(self, tensor: torch.Tensor, dim: int = 0, pad_index: float = 0) -> torch.Tensor:
        if self.num_processes <= 1:
            return tensor
        size = torch.tensor([tensor.shape[dim]], device=tensor.device)
        sizes = self.gather(size)
        mx = int(sizes.max())
        if mx == tensor.shape[dim]:
            return tensor
        shape = list(tensor.shape)
        shape[dim] = mx
        out = tensor.new_full(shape, pad_index)
        idx = [slice(None)] * tensor.dim()
        idx[dim] = slice(0, tensor.shape[dim])
        out[tuple(idx)] = tensor
        return out


def broadcast_object_list(objects: List[Any], from_process: int = 0) -> List[Any]:
    """accelerate.utils.broadcast_object_list (VDDP:24, 1725, 1816)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast_object_list(objects, src=from_process)
    return objects


class DistributedDataParallelKwargs:
    """accelerate.DistributedDataParallelKwargs stand-in (main.py:32).  find_unused_parameters is moot here:
    unused parameters simply keep zero gradients in the arena."""

    def __init__(self, **kw):
        self.kw = kw


class InitProcessGroupKwargs:
    """accelerate.InitProcessGroupKwargs stand-in (main.py:33)."""

    def __init__(self, **kw):
        self.kw = kw

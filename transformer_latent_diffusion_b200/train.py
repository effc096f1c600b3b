"""Training step — drop-in for ``tld.train`` on the hot path (reference tld/train.py).

* ``denoiser_autograd_forward`` makes ``Denoiser.forward`` differentiable: forward = ``tld_train_forward`` (keeps the
  activations in library-owned buffers), backward = ``tld_train_backward`` (tcgen05 dgrad/wgrad GEMMs + the backward
  kernels of ``csrc/backward.cu`` / ``attention_bwd.cu``), gradients handed to autograd per parameter, so the caller's
  optimiser / EMA / gradient clipping code is unchanged.
* ``main(config)`` mirrors ``tld.train.main`` (tld/train.py:62-176): Beta(1,2.5) noise levels, linear-interp noising,
  15 % label dropout, MSE against the clean latent, Adam(lr), EMA(alpha) on rank 0.  ``accelerate``/``wandb`` are not
  available offline: the data-parallel part is plain ``torch.distributed`` (one process per GPU, NCCL all-reduce of
  the flattened gradients = what DDP does inside ``accelerator.backward``), logging goes to stdout.
  bf16 tensor-core operands with fp32 accumulation/master weights replace the reference's fp16 autocast + GradScaler
  (no loss scaling needed).
* ``eval_gen`` (tld/train.py:23-40) and the checkpoint save / resume of ``main`` (tld/train.py:92-102,140-158) use the
  reference's file layout ``{"model_ema", "opt_state", "global_step"}``, so a checkpoint written by either side loads in
  the other.
"""
from __future__ import annotations

import copy
import weakref
from dataclasses import asdict
from typing import Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib
from .configs import ModelConfig


class _DenoiserFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, noise_level, label, *params):
        B = x.shape[0]
        h = module._ensure_handle(x.device)
        xf = x.detach().to(torch.float32).contiguous()
        tf = noise_level.detach().to(device=x.device, dtype=torch.float32).reshape(B).contiguous()
        lf = label.detach().to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(xf)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().tld_train_forward(h, _lib.ptr(xf), _lib.ptr(tf), _lib.ptr(lf), _lib.ptr(out), B,
                                                     _lib.current_stream_ptr(x.device)), "tld_train_forward")
        ctx.module, ctx.batch, ctx.handle = module, B, h
        ctx.serial = _lib.load().tld_forward_serial(h)   # the activations live in per-handle buffers (see backward)
        ctx.keys = [k for k, _ in module.named_parameters()]
        ctx.meta = [(p.shape, p.dtype) for p in params]
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        dev = d_out.device
        g = d_out.detach().to(torch.float32).contiguous()
        if lib.tld_forward_serial(ctx.handle) != ctx.serial:
            raise _lib.TldError(
                "backward through a Denoiser forward whose saved activations were overwritten by a later forward of the "
                "same module (the activation buffers belong to the module's library handle, not to the autograd node): "
                "run backward() before the next forward / generate call of this module")
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr(dev)
            _lib.check(lib.tld_train_backward(ctx.handle, _lib.ptr(g), ctx.batch, st), "tld_train_backward")
            ctx.module._tld_grads_allreduced = _overlapped_allreduce(ctx.module, ctx.handle, dev)
            # Default: ONE device copy snapshots the whole gradient arena (fresh storage per backward, so gradients that the
            # caller keeps or accumulates never alias the library's buffers); each parameter gets a view of the snapshot.
            # With `module.grad_views = True` (set by train.main for the fused optimiser) autograd receives views of the
            # arena itself: no 404 MB copy per step, valid until the next backward of this module - the contract of a loop
            # that calls optimizer.zero_grad() (set_to_none) every step, as tld/train.py:164 does.
            view, offsets = _grad_arena(ctx.module, ctx.handle, dev, ctx.keys)
            snap = view if getattr(ctx.module, "grad_views", False) else view.clone()
            grads = []
            for (off, n), (shape, dtype) in zip(offsets, ctx.meta):
                t = snap[off:off + n].view(shape)
                grads.append(t if dtype == torch.float32 else t.to(dtype))
        return (None, None, None, None, *grads)


_ARENA_CACHE = weakref.WeakKeyDictionary()   # module -> {arena pointer, tensor view, segment table, side stream}


class _DeviceArray:
    """zero-copy view of a device buffer for ``torch.as_tensor`` (CUDA array interface v2)"""

    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _grad_arena(module, handle, dev, keys=None):
    """(fp32 tensor view of the library's gradient arena, [(offset, numel) per key]) - cached per module / arena"""
    import ctypes as C

    lib = _lib.load()
    L = module.n_layers
    arena, total = C.c_void_p(), C.c_int64()
    seg = (C.c_int64 * (2 * (L + 2)))()
    _lib.check(lib.tld_train_grad_layout(handle, C.byref(arena), C.byref(total), seg, L + 2), "tld_train_grad_layout")
    cache = _ARENA_CACHE.setdefault(module, {})   # outside the module: copy.deepcopy(model) must not drag views/streams along
    if cache.get("ptr") != arena.value:
        cache.clear()
        cache["ptr"] = arena.value
        cache["view"] = torch.as_tensor(_DeviceArray(arena.value, total.value), device=dev)
        cache["segments"] = [(int(seg[2 * i]), int(seg[2 * i + 1])) for i in range(L + 2)]
    if keys is not None and cache.get("keys") != keys:
        offs = []
        for key in keys:
            off, n = C.c_int64(), C.c_int64()
            _lib.check(lib.tld_train_grad_offset(handle, key.encode(), C.byref(off), C.byref(n)), f"grad_offset({key})")
            offs.append((off.value, n.value))
        cache["keys"], cache["offsets"] = list(keys), offs
    return cache["view"], cache.get("offsets")


def _overlapped_allreduce(module, handle, dev) -> bool:
    """Data-parallel gradient averaging overlapped with the backward (what DDP does inside ``accelerator.backward``,
    tld/train.py:109,169).  ``tld_train_backward`` has only been ENQUEUED at this point; every decoder block's gradients
    form one contiguous range of the library's gradient arena and become final at a recorded event, so a side stream
    waits for block l's event and all-reduces that range in place (NCCL, average) while the compute stream is still
    differentiating blocks l-1 .. 0.  Returns False (nothing done) without an NCCL process group of more than one rank."""
    import ctypes as C

    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"):
        return False
    if not getattr(module, "overlap_grad_allreduce", True):
        return False
    lib = _lib.load()
    L = module.n_layers
    view, _ = _grad_arena(module, handle, dev)
    cache = _ARENA_CACHE[module]
    seg = [v for pair in cache["segments"] for v in pair]
    if "stream" not in cache:
        cache["stream"] = torch.cuda.Stream(device=dev)
    side = cache["stream"]
    with torch.cuda.stream(side):
        for l in range(L - 1, -1, -1):   # the backward finishes the blocks in this order
            _lib.check(lib.tld_train_wait_grad(handle, l, side.cuda_stream), "tld_train_wait_grad")
            dist.all_reduce(view[seg[2 * l]: seg[2 * l] + seg[2 * l + 1]], op=dist.ReduceOp.AVG)
        _lib.check(lib.tld_train_wait_grad(handle, L, side.cuda_stream), "tld_train_wait_grad")
        for i in (L, L + 1):             # all kv_linear weights; embedding / conditioning / out-projection
            dist.all_reduce(view[seg[2 * i]: seg[2 * i] + seg[2 * i + 1]], op=dist.ReduceOp.AVG)
    torch.cuda.current_stream(dev).wait_stream(side)   # the per-parameter copies below read averaged gradients
    return True


def denoiser_autograd_forward(module, x: Tensor, noise_level: Tensor, label: Tensor) -> Tensor:
    """Differentiable ``Denoiser.forward`` (w.r.t. the parameters; the latents/labels get no gradient)."""
    n_tok = (module.image_size // module.patch_size) ** 2
    if n_tok > 256 and n_tok % 256:
        raise _lib.TldError("training path: tokens per sample must be <= 256 or a multiple of 256")
    if module.embed_dim % 128:
        raise _lib.TldError("training path: embed_dim must be a multiple of 128")
    params = [p for _, p in module.named_parameters()]
    return _DenoiserFn.apply(module, x, noise_level, label, *params)


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def count_parameters_per_layer(model: nn.Module) -> None:
    for name, param in model.named_parameters():
        print(f"{name}: {param.numel()} parameters")


def eval_gen(diffuser, labels: Tensor, img_size: int, out_path: Optional[str] = None):
    """Fixed-seed sample grid from the EMA weights (tld/train.py:23-40): every validation embedding twice, 16 images,
    guidance 4.5, seed 10, 40 DPM-Solver++ steps, sharpness shift 0.1, 8 per row with 4 px padding.  Returns the PIL image
    and saves it under the reference's file name (or ``out_path``)."""
    from PIL import Image

    class_guidance, seed = 4.5, 10
    arr, _ = diffuser.generate_grid_uint8(torch.repeat_interleave(labels, 2, dim=0), nrow=8, padding=4, num_imgs=16,
                                          class_guidance=class_guidance, seed=seed, n_iter=40, exponent=1, sharp_f=0.1,
                                          img_size=img_size)       # (x+1)/2, clip, *255 -> uint8 on the device
    img = Image.fromarray(arr)
    img.save(out_path or f"emb_val_cfg:{class_guidance}_seed:{seed}.png")
    return img


def update_ema(ema_model: nn.Module, model: nn.Module, alpha: float = 0.999) -> None:
    """ema = alpha * ema + (1 - alpha) * param (tld/train.py:55-58), one fused foreach pass"""
    with torch.no_grad():
        ema_p = [p.data for p in ema_model.parameters()]
        cur_p = [p.data for p in model.parameters()]
        torch._foreach_mul_(ema_p, alpha)
        torch._foreach_add_(ema_p, cur_p, alpha=1 - alpha)


def noise_batch(x: Tensor, y: Tensor, noise_level: Tensor, noise: Tensor, drop_mask: Tensor, vae_scale_factor: float):
    """tld/train.py:122-138 with the random draws passed in: returns (clean target, noisy input, sigma[B,1], labels)."""
    x = x / vae_scale_factor
    s = noise_level.double().view(-1, 1, 1, 1)
    x_noisy = (s * noise.double() + (1 - s) * x.double()).float()
    label = y.clone()
    label[drop_mask] = 0
    return x, x_noisy, noise_level.float().view(-1, 1), label


def allreduce_gradients(model: nn.Module) -> None:
    """Average the gradients over the data-parallel ranks with ONE flattened NCCL all-reduce (what DDP does in
    accelerator.backward, tld/train.py:169); no-op without an initialised process group."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    if getattr(model, "_tld_grads_allreduced", False):   # already averaged, overlapped with the backward
        model._tld_grads_allreduced = False
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat)
    flat.div_(dist.get_world_size())
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)


def shard_indices(perm: Tensor, rank: int, world: int) -> Tensor:
    """This rank's share of a shuffled index list, padded by wrapping around to ceil(n / world) entries exactly like
    torch.utils.data.DistributedSampler (which the reference's accelerate-prepared DataLoader uses, tld/train.py:109):
    every rank runs the same number of steps, so no rank waits forever in an all-reduce or barrier."""
    n = perm.numel()
    per_rank = -(-n // world)
    total = per_rank * world
    if total > n:
        perm = perm.repeat(-(-total // n))[:total]
    return perm[rank::world]


def train_step(model: nn.Module, optimizer, x: Tensor, x_noisy: Tensor, sigma: Tensor, label: Tensor) -> Tensor:
    """zero_grad -> forward -> MSE -> backward -> (all-reduce) -> step (tld/train.py:160-170). Returns the loss tensor."""
    model.train()
    optimizer.zero_grad()
    pred = model(x_noisy, sigma, label)
    loss = torch.nn.functional.mse_loss(pred, x)
    loss.backward()
    allreduce_gradients(model)
    optimizer.step()
    return loss.detach()


def main(config: ModelConfig, device: Optional[torch.device] = None, log_every: int = 50, vae: Optional[nn.Module] = None,
         eval_dir: Optional[str] = None, fused_optimizer: bool = True) -> nn.Module:
    """Training loop with the reference's semantics (tld/train.py:62-176). Returns the EMA model (rank 0) / model.

    ``vae``: decoder used by the periodic ``eval_gen`` on rank 0 (the reference downloads it with
    ``AutoencoderKL.from_pretrained``, tld/train.py:78 - no network here, so the caller injects it; without one the
    evaluation images are skipped and only the checkpoint is written).  ``from_scratch=False`` resumes from
    ``train_config.model_name`` exactly as tld/train.py:92-102 (EMA weights into the model, optimizer state, step).
    ``fused_optimizer``: Adam + EMA as one kernel over flat arenas (``optim.FusedAdamEMA``, same arithmetic and checkpoint
    layout as ``torch.optim.Adam`` + ``update_ema``); False keeps the reference's two separate torch calls."""
    import os

    import torch.distributed as dist

    from .denoiser import Denoiser
    from .diffusion import DiffusionGenerator

    denoiser_config, train_config, dataconfig = config.denoiser_config, config.train_config, config.data_config
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = device or torch.device("cuda", torch.cuda.current_device())

    latent_train_data = torch.tensor(np.load(dataconfig.latent_path), dtype=torch.float32)
    train_label_embeddings = torch.tensor(np.load(dataconfig.text_emb_path), dtype=torch.float32)
    emb_val = None
    if dataconfig.val_path and os.path.exists(dataconfig.val_path):
        emb_val = torch.tensor(np.load(dataconfig.val_path), dtype=torch.float32).to(device)
    n = latent_train_data.shape[0]
    model = Denoiser(**asdict(denoiser_config)).to(device)
    global_step = 0
    full_state_dict = None
    if not train_config.from_scratch:  # tld/train.py:92-102 (the wandb.restore download is the caller's business)
        full_state_dict = torch.load(train_config.model_name, map_location=device)
        model.load_state_dict(full_state_dict["model_ema"])
        global_step = int(full_state_dict["global_step"])
    if world > 1:  # identical initial weights on every rank (DDP broadcasts rank 0's)
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    ema_model = copy.deepcopy(model) if rank == 0 else None
    if fused_optimizer:
        from .optim import FusedAdamEMA

        optimizer = FusedAdamEMA(model, lr=train_config.lr, ema_model=ema_model, alpha=train_config.alpha)
        model.grad_views = True   # the loop below zeroes the gradients every step: no snapshot copy of the arena needed
    else:
        optimizer = torch.optim.Adam(model.parameters(), lr=train_config.lr)
    if full_state_dict is not None:
        optimizer.load_state_dict(full_state_dict["opt_state"])
    diffuser = None
    if rank == 0:
        print(count_parameters(model))
        if vae is not None:
            diffuser = DiffusionGenerator(ema_model, vae, device, torch.float32)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)   # noise / label-dropout draws on the device (tld/train.py:128,133)
    for epoch in range(1, train_config.n_epoch + 1):
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(epoch))  # same shuffle on every rank
        shard = shard_indices(perm, rank, world)
        for i in range(0, len(shard), train_config.batch_size):
            idx = shard[i:i + train_config.batch_size]
            x, y = latent_train_data[idx].to(device), train_label_embeddings[idx].to(device)
            noise_level = torch.tensor(np.random.beta(train_config.beta_a, train_config.beta_b, len(x)), device=device)
            noise = torch.randn(x.shape, generator=gen, device=device)
            mask = torch.rand(y.size(0), generator=gen, device=device) < 0.15
            xs, x_noisy, sigma, label = noise_batch(x, y, noise_level, noise, mask, config.vae_cfg.vae_scale_factor)
            if global_step % train_config.save_and_eval_every_iters == 0:  # tld/train.py:140-158, before the step
                if world > 1:
                    dist.barrier()
                if rank == 0:
                    if diffuser is not None and emb_val is not None:
                        ema_model.eval()
                        out = eval_gen(diffuser, emb_val, denoiser_config.image_size,
                                       out_path=os.path.join(eval_dir, f"eval_step{global_step}.png") if eval_dir else None)
                        out.save(os.path.join(eval_dir, "img.jpg") if eval_dir else "img.jpg")
                    if train_config.save_model and train_config.model_name:
                        torch.save({"model_ema": ema_model.state_dict(), "opt_state": optimizer.state_dict(),
                                    "global_step": global_step}, train_config.model_name)
            loss = train_step(model, optimizer, xs, x_noisy, sigma, label)
            if rank == 0:
                if not fused_optimizer:   # the fused step has already folded the EMA update in
                    update_ema(ema_model, model, alpha=train_config.alpha)
                if global_step % log_every == 0:
                    print(f"epoch {epoch} step {global_step} train_loss {float(loss):.5f}")
            global_step += 1
    if rank == 0:
        main.last_global_step = global_step  # convenience for callers / tests (the reference keeps it local)
    return ema_model if rank == 0 else model

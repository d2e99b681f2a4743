"""Optimizer for the trainable tail: detectron2/solver/build.py:93-137 `build_optimizer` mirrored, with the SGD
update running as one C-ABI kernel per parameter that also rewrites the bf16 kernel-layout copy of the weight
(drn_sgd_step), so an optimizer step does not trigger a separate re-pack of fc6's 205 M weights.

`FusedSGD` is a torch.optim.Optimizer (param_groups / state_dict / lr schedulers work as usual) with the
arithmetic of torch.optim.SGD; parameters that are not on a CUDA device take torch's own update."""
import torch

from . import ops


class FusedSGD(torch.optim.Optimizer):
    """`model` (optional) lets the update kernel rewrite the model's bf16 kernel-layout weight copies in the same pass.
    Without it the parameters are still updated correctly: every stepped parameter's autograd version is bumped, so the
    model's derived layouts (keyed on the versions) re-pack themselves on the next forward.
    `clip` = (type, value, norm_type): per-parameter gradient clipping before the update, as
    detectron2/solver/build.py:19-92 wraps the optimizer when SOLVER.CLIP_GRADIENTS.ENABLED."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, nesterov=False, model=None, clip=None):
        if nesterov and momentum <= 0:
            raise ValueError("Nesterov momentum requires a momentum")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov))
        self.model = model
        self._sharded = {}  # parameter -> distributed.ShardedLinearTrainer (its gradient never reaches p.grad)
        if clip is not None and clip[0] not in ("value", "norm"):
            raise ValueError(f"SOLVER.CLIP_GRADIENTS.CLIP_TYPE must be 'value' or 'norm', got {clip[0]!r}")
        self.clip = clip

    def attach_sharded(self, param, sharder):
        """`param` (fc6.weight) is stepped shard by shard by `sharder` (distributed.ShardedLinearTrainer): its gradient was
        reduce-scattered into the owners' windows by the weight-gradient GEMM, every rank updates the rows it owns and
        broadcasts the refreshed bf16 kernel rows."""
        if self.clip is not None:
            raise NotImplementedError("gradient clipping of a sharded parameter needs its global norm; not implemented")
        self._sharded[param] = sharder

    def _clip(self, p):
        kind, value, norm_type = self.clip
        if kind == "value":
            torch.nn.utils.clip_grad_value_(p, value)
        else:
            torch.nn.utils.clip_grad_norm_(p, value, norm_type)

    def _packed_copies(self):
        """parameter -> (bf16 kernel-layout buffer, c49) for the single-layer tensor-core packs of the model."""
        out = {}
        if self.model is None:
            return out
        for m in self.model.modules():
            cache = getattr(m, "_cache", None)
            if not isinstance(m, torch.nn.Linear) or not cache:
                continue
            for (precision, perm), hit in cache.items():
                w = hit.get("w")
                if precision != "fp32" and torch.is_tensor(w) and w.dtype == torch.bfloat16 and tuple(w.shape) == tuple(m.weight.shape):
                    out[m.weight] = (hit, perm or 0, m)
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        packed = self._packed_copies()
        touched = False
        stepped_shards = []
        for group in self.param_groups:
            lr, mom, wd, nest = group["lr"], group["momentum"], group["weight_decay"], group["nesterov"]
            for p in group["params"]:
                sharder = self._sharded.get(p)
                if sharder is not None:
                    sharder.fence()  # every rank's weight-gradient GEMM has finished: the slots are final
                    sharder.step(lr, mom, wd, nest)
                    torch.autograd.graph.increment_version(p)
                    stepped_shards.append(sharder)
                    touched = True
                    continue
                if p.grad is None:
                    continue
                if self.clip is not None:
                    self._clip(p)
                st = self.state[p]
                first = "momentum_buffer" not in st
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    d = p.grad.add(p, alpha=wd) if wd != 0 else p.grad
                    if mom != 0:
                        buf = st["momentum_buffer"] = torch.clone(d).detach() if first else st["momentum_buffer"].mul_(mom).add_(d)
                        d = d.add(buf, alpha=mom) if nest else buf
                    p.add_(d, alpha=-lr)
                    continue
                if first and mom != 0:
                    st["momentum_buffer"] = torch.empty_like(p)
                hit = packed.get(p)
                ops.sgd_step(p, p.grad.contiguous(), st.get("momentum_buffer"), hit[0]["w"] if hit else None, hit[1] if hit else 0,
                             lr, mom, wd, nest, first)
                # the kernel wrote through raw pointers: tell autograd (and every cache keyed on the version) so
                torch.autograd.graph.increment_version(p)
                touched = True
        for sharder in stepped_shards:
            sharder.fence()  # every rank's refreshed rows have landed in this rank's weight buffer (and its slots are free again)
        if touched and self.model is not None:
            self._after_step(packed)
        return loss

    def _after_step(self, packed):
        """Bring the derived layouts the kernel did not rewrite itself up to date, in place and now (on the step's
        stream) -- bias copies of the fused layers, every other pack of a stepped layer (fp32-mode layouts, the
        concatenated heads) -- and re-key the fused packs to the bumped parameter versions so that the next forward
        does not pack fc6's 205 M weights a second time."""
        from .modeling import _versions

        fused = {id(hit[0]) for hit in packed.values()}
        for m in self.model.modules():
            cache = getattr(m, "_cache", None)
            if not isinstance(m, torch.nn.Linear) or not cache:
                continue
            for (precision, perm), hit in list(cache.items()):
                if id(hit) in fused:
                    hit["bias"].copy_(m.bias.detach())
                    hit["key"] = (precision, perm, _versions([m.weight, m.bias]))
                else:
                    hit["key"] = None
                    m.packed(precision, permute_c49=perm)  # re-packed in place, now
        rh = getattr(self.model, "roi_heads", None)
        if rh is not None and getattr(rh, "_heads_cache", None) is not None:
            rh._heads_cache["key"] = None
            rh._heads_packed()


def build_optimizer(cfg, model):
    """detectron2/solver/build.py:93-137: one param group per trainable parameter, lr = BASE_LR (x BIAS_LR_FACTOR for
    biases), weight decay = WEIGHT_DECAY / WEIGHT_DECAY_BIAS / WEIGHT_DECAY_NORM; SGD with MOMENTUM / NESTEROV."""
    s = cfg.SOLVER
    norm_types = (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d, torch.nn.SyncBatchNorm, torch.nn.GroupNorm,
                  torch.nn.InstanceNorm1d, torch.nn.InstanceNorm2d, torch.nn.InstanceNorm3d, torch.nn.LayerNorm,
                  torch.nn.LocalResponseNorm)
    params, memo = [], set()
    for module in model.modules():
        for key, value in module.named_parameters(recurse=False):
            if not value.requires_grad or value in memo:
                continue
            memo.add(value)
            lr, wd = s.BASE_LR, s.WEIGHT_DECAY
            if isinstance(module, norm_types):
                wd = s.WEIGHT_DECAY_NORM
            elif key == "bias":
                lr, wd = s.BASE_LR * s.BIAS_LR_FACTOR, s.WEIGHT_DECAY_BIAS
            params.append({"params": [value], "lr": lr, "weight_decay": wd})
    clip = None
    cg = s.get("CLIP_GRADIENTS") if hasattr(s, "get") else getattr(s, "CLIP_GRADIENTS", None)
    if cg is not None and cg.ENABLED:  # solver/build.py:65-92 maybe_add_gradient_clipping
        clip = (cg.CLIP_TYPE, cg.CLIP_VALUE, cg.NORM_TYPE)
    return FusedSGD(params, s.BASE_LR, momentum=s.MOMENTUM, nesterov=s.NESTEROV, model=model, clip=clip)


def run_step(model, optimizer, batched_inputs, iteration, iter_size=1, start_iter=0, grad_sync=None):
    """One iteration of projects/WSL/tools/train_net.py:65-117 (`Trainer.run_step`) around the B200 model: forward, the sum
    of the loss dict divided by `WSL.ITER_SIZE`, backward, and -- only on iterations that are a multiple of ITER_SIZE --
    the optimizer step followed by zero_grad; gradients accumulate in `p.grad` over the iterations in between
    (zero_grad also runs once before the first iteration, as there).  With a captured forward plan the backward of
    iteration i has to run before the forward of iteration i+1 (it reads the plan's buffers): this order does.
    grad_sync: a distributed.GradientSynchronizer attached to the model (data-parallel runs) -- its `finish()` hands the
    averaged gradients of THIS backward to `p.grad` (accumulating) before the step, as DistributedDataParallel would.
    Returns the (un-divided) loss dict, detached."""
    assert model.training, "[run_step] model was changed to eval mode!"
    assert iter_size >= 1
    loss_dict = model(batched_inputs)
    losses = sum(loss_dict.values())
    if not torch.isfinite(losses).all():  # detectron2/engine/train_loop.py:_detect_anomaly
        raise FloatingPointError(f"Loss became infinite or NaN at iteration={iteration}!\nloss_dict = {loss_dict}")
    if iteration == start_iter:
        optimizer.zero_grad()
    (losses / iter_size).backward()
    if grad_sync is not None:
        grad_sync.finish()
    if iteration % iter_size == 0:
        optimizer.step()
        optimizer.zero_grad()
    return {k: v.detach() for k, v in loss_dict.items()}

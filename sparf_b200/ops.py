"""torch.autograd adapters over the C ABI (include/sparf_b200.h).

PyTorch is plumbing here: it owns device memory, the stream and the autograd tape; every number is
produced by the kernels in csrc/.  All functions take / return fp32 CUDA tensors and raise if the
native library is unavailable (no fallback).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import SparfMLP, SparfMLPGrad, check

_ENGINE = [_lib.ENGINE_AUTO]
# keep the training forward's operand images for the backward (tcgen05 engine); False = always recompute
USE_TAPE = [True]
# Opt-in: let the MLP backward accumulate straight into existing `param.grad` storage (the C ABI accumulates, +=)
# instead of returning fresh gradient tensors for autograd to add.  Saves ~20 tiny kernels and a 2 MB memset per
# render pass.  Only valid for plain `loss.backward()` training loops (no torch.autograd.grad / grad hooks / DDP
# hooks on these parameters).  Per parameter set: sparf_b200.distributed.FlatGradients marks its parameters
# (`p._sparf_inplace_grad`); the global switch forces it for every model of the process.
ACCUMULATE_INTO_PARAM_GRAD = [False]

# optional device-side timing of the MLP kernels (bench.py roofline): CUDA events on the launching stream
PROFILE_ON = [False]
PROFILE = []
# MLP sample-evaluations issued through this module (forward calls, backward calls): bench.py's FLOP accounting
EVALS = {"fwd": 0, "bwd": 0}


class _timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if PROFILE_ON[0]:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE_ON[0]:
            self.e1.record()
            PROFILE.append((self.tag, self.e0, self.e1))


def profile_total_ms(tag_prefix="mlp"):
    """Sum of the recorded intervals (call after torch.cuda.synchronize())."""
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for tag, e0, e1 in PROFILE if tag.startswith(tag_prefix))


def set_engine(name_or_id) -> None:
    """Select the MLP engine: 'auto' | 'simt_fp32' | 'tc_3x' | 'tc_1x'."""
    _ENGINE[0] = _lib.ENGINES[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)


def get_engine() -> int:
    return _ENGINE[0]


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_tensor_device(fn):
    """Run an op body with the CUDA device of its first tensor argument current: the kernels launch on the current
    device's current stream, so a Graph on cuda:1 works without a global torch.cuda.set_device (one process driving
    several devices)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in args:
            if torch.is_tensor(a) and a.is_cuda:
                dev = a.device
                break
            saved = getattr(a, "saved_tensors", None) if not torch.is_tensor(a) and hasattr(a, "needs_input_grad") else None
            if saved:
                cand = [t for t in saved if torch.is_tensor(t) and t.is_cuda]
                if cand:
                    dev = cand[0].device
                    break
        if dev is None:
            for a in kwargs.values():
                if torch.is_tensor(a) and a.is_cuda:
                    dev = a.device
                    break
        if dev is None and kwargs.get("device") is not None:
            dev = torch.device(kwargs["device"])
        if dev is None or (torch.cuda.current_device() == (dev.index if dev.index is not None else torch.cuda.current_device())):
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "sparf_b200 ops need CUDA tensors (there is no CPU path)"
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# MLP description
# ------------------------------------------------------------------------------------------------
class MLPSpec:
    """Static description of one NeRF network (shapes + c2f schedule), independent of the tensors."""

    def __init__(self, n_trunk=8, width=256, head_width=128, skip_layer=4, L_xyz=10, L_view=4, barf_c2f=None):
        self.n_trunk, self.width, self.head_width, self.skip_layer = n_trunk, width, head_width, skip_layer
        self.L_xyz, self.L_view = L_xyz, L_view
        self.barf_c2f = tuple(barf_c2f) if barf_c2f is not None else None

    def n_params(self) -> int:
        return 2 * self.n_trunk + 4

    def fill(self, params: Sequence[torch.Tensor], progress: Optional[torch.Tensor]) -> Tuple[SparfMLP, list]:
        """params = [trunk_w0, trunk_b0, ..., head_w0, head_b0, head_w1, head_b1] (nn.Linear tensors)."""
        keep = [_f32c(p.detach()) for p in params]
        m = SparfMLP()
        m.n_trunk, m.width, m.head_width, m.skip_layer = self.n_trunk, self.width, self.head_width, self.skip_layer
        m.L_xyz, m.L_view = self.L_xyz, self.L_view
        m.use_c2f = 1 if self.barf_c2f is not None else 0
        if self.barf_c2f is not None:
            start, end = self.barf_c2f
            m.c2f_start = float(start)
            m.c2f_range = float(end - start)  # python-double subtraction, like the reference
            assert progress is not None
            prog = _f32c(progress.detach()).reshape(1)
            keep.append(prog)
            m.progress = prog.data_ptr()
        else:
            m.progress = 0
        for i in range(self.n_trunk):
            m.trunk_w[i] = keep[2 * i].data_ptr()
            m.trunk_b[i] = keep[2 * i + 1].data_ptr()
        o = 2 * self.n_trunk
        m.head_w[0], m.head_b[0] = keep[o].data_ptr(), keep[o + 1].data_ptr()
        m.head_w[1], m.head_b[1] = keep[o + 2].data_ptr(), keep[o + 3].data_ptr()
        return m, keep

    def grad_struct(self, grads: Sequence[torch.Tensor]) -> SparfMLPGrad:
        g = SparfMLPGrad()
        for i in range(self.n_trunk):
            g.trunk_w[i] = grads[2 * i].data_ptr()
            g.trunk_b[i] = grads[2 * i + 1].data_ptr()
        o = 2 * self.n_trunk
        g.head_w[0], g.head_b[0] = grads[o].data_ptr(), grads[o + 1].data_ptr()
        g.head_w[1], g.head_b[1] = grads[o + 2].data_ptr(), grads[o + 3].data_ptr()
        return g


_WS = {}
_WS_RETIRED = []   # outgrown buffers stay alive: a captured CUDA graph may have their addresses baked into its kernels


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only per-(device, stream) scratch buffer.  Reuse is stream-ordered, so one buffer per stream is safe; a
    buffer that is outgrown is retired, never freed (CUDA graphs captured earlier keep replaying into it)."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _WS_RETIRED.append(buf)
            nbytes = max(nbytes, int(1.5 * buf.numel()))   # geometric growth bounds the retired bytes
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------
# MLP: sigma, rgb = NeRF.forward_samples(o, d, t)
# ------------------------------------------------------------------------------------------------
class MLPFunction(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, spec: MLPSpec, engine: int, grad_mode: bool, origins, dirs, t, noise, progress, *params):
        L = _lib.lib()
        origins, dirs, t = _f32c(origins), _f32c(dirs), _f32c(t)
        R, S = t.shape
        assert origins.shape == (R, 3) and dirs.shape == (R, 3)
        noise_c = _f32c(noise) if noise is not None else None
        m, keep = spec.fill(params, progress)
        sigma = torch.empty(R, S, device=t.device, dtype=torch.float32)
        rgb = torch.empty(R, S, 3, device=t.device, dtype=torch.float32)
        nbytes = L.sparf_mlp_workspace_bytes(ctypes.byref(m), R, S, 0, engine)
        ws = _workspace(nbytes, t.device)
        # Training forward: when a gradient will be asked for and the engine offers it, keep a "tape" (the
        # per-layer operand images) so that the backward skips the forward recompute.
        # (grad_mode: autograd is recording at the call site -- under torch.no_grad() nothing is kept)
        EVALS["fwd"] += R * S
        wants_grad = grad_mode and (any(ctx.needs_input_grad[i] for i in (3, 4)) or any(ctx.needs_input_grad[8:]))
        tape_bytes = L.sparf_mlp_tape_bytes(ctypes.byref(m), engine, R, S) if (wants_grad and USE_TAPE[0]) else 0
        ctx.tape = None
        with _timed("mlp_forward"):
            if tape_bytes:
                ctx.tape = torch.empty(tape_bytes, dtype=torch.uint8, device=t.device)
                check(L.sparf_mlp_forward_tape(ctypes.byref(m), engine, R, S, _ptr(origins), _ptr(dirs), _ptr(t),
                                               _ptr(noise_c), _ptr(sigma), _ptr(rgb), _ptr(ctx.tape), tape_bytes, _ptr(ws),
                                               ws.numel(), _stream()), "mlp_forward_tape")
            else:
                check(L.sparf_mlp_forward(ctypes.byref(m), engine, R, S, _ptr(origins), _ptr(dirs), _ptr(t), _ptr(noise_c),
                                          _ptr(sigma), _ptr(rgb), _ptr(ws), ws.numel(), _stream()), "mlp_forward")
        ctx.spec, ctx.engine = spec, engine
        ctx.noise = noise_c
        ctx.progress = progress
        ctx.param_refs = params if (ACCUMULATE_INTO_PARAM_GRAD[0] or
                                    all(getattr(p, "_sparf_inplace_grad", False) for p in params)) else None
        if ctx.tape is not None:
            ctx.save_for_backward(origins, dirs, t, sigma, rgb, *params)
        else:
            ctx.save_for_backward(origins, dirs, t, *params)
        return sigma, rgb

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g_sigma, g_rgb):
        L = _lib.lib()
        if ctx.tape is not None:
            origins, dirs, t, sigma_f, rgb_f, *params = ctx.saved_tensors
        else:
            origins, dirs, t, *params = ctx.saved_tensors
        spec = ctx.spec
        R, S = t.shape
        EVALS["bwd"] += R * S
        g_sigma = _f32c(g_sigma) if g_sigma is not None else torch.zeros(R, S, device=t.device)
        g_rgb = _f32c(g_rgb) if g_rgb is not None else torch.zeros(R, S, 3, device=t.device)
        m, keep = spec.fill(params, ctx.progress)
        inplace = ctx.param_refs is not None and all(
            p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 for p in ctx.param_refs)
        if inplace:
            grads = [p.grad for p in ctx.param_refs]
        else:
            sizes = [p.numel() for p in params]
            flat = torch.zeros(sum(sizes), device=t.device, dtype=torch.float32)
            grads, o = [], 0
            for p, n in zip(params, sizes):
                grads.append(flat[o:o + n].view(p.shape))
                o += n
        gs = spec.grad_struct(grads)
        need_o, need_d = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        d_o = torch.zeros_like(origins) if (need_o or need_d) else None
        d_d = torch.zeros_like(dirs) if (need_o or need_d) else None
        nbytes = L.sparf_mlp_workspace_bytes(ctypes.byref(m), R, S, 2 if ctx.tape is not None else 1, ctx.engine)
        ws = _workspace(nbytes, t.device)
        with _timed("mlp_backward"):
            if ctx.tape is not None:
                check(L.sparf_mlp_backward_tape(ctypes.byref(m), ctx.engine, R, S, _ptr(origins), _ptr(dirs), _ptr(t),
                                                _ptr(sigma_f), _ptr(rgb_f), _ptr(g_sigma), _ptr(g_rgb), ctypes.byref(gs),
                                                _ptr(d_o), _ptr(d_d), _ptr(ctx.tape), ctx.tape.numel(), _ptr(ws), ws.numel(),
                                                _stream()), "mlp_backward_tape")
                ctx.tape = None
            else:
                check(L.sparf_mlp_backward(ctypes.byref(m), ctx.engine, R, S, _ptr(origins), _ptr(dirs), _ptr(t),
                                           _ptr(ctx.noise), _ptr(g_sigma), _ptr(g_rgb), ctypes.byref(gs), _ptr(d_o),
                                           _ptr(d_d), _ptr(ws), ws.numel(), _stream()), "mlp_backward")
        if inplace:
            grads = [None] * len(grads)
        return (None, None, None, d_o if need_o else None, d_d if need_d else None, None, None, None, *grads)


def mlp_forward(spec: MLPSpec, origins, dirs, t, params: Sequence[torch.Tensor], *, noise=None, progress=None,
                engine: Optional[int] = None):
    """origins/dirs [R,3], t [R,S] -> (sigma [R,S], rgb [R,S,3]); differentiable w.r.t. origins, dirs, params."""
    eng = get_engine() if engine is None else engine
    return MLPFunction.apply(spec, eng, torch.is_grad_enabled(), origins, dirs, t, noise, progress, *params)


# ------------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------------
class CompositeFunction(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, sigma, rgb, t, dirs, white_bg: bool):
        L = _lib.lib()
        sigma, rgb, t, dirs = _f32c(sigma), _f32c(rgb), _f32c(t), _f32c(dirs)
        R, S = t.shape
        dev = t.device
        rgb_map = torch.empty(R, 3, device=dev)
        depth, opacity = torch.empty(R, device=dev), torch.empty(R, device=dev)
        depth_var, rgb_var = torch.empty(R, device=dev), torch.empty(R, device=dev)
        weights, all_cum = torch.empty(R, S, device=dev), torch.empty(R, device=dev)
        check(L.sparf_composite_forward(R, S, _ptr(sigma), _ptr(rgb), _ptr(t), _ptr(dirs), int(white_bg), _ptr(rgb_map),
                                        _ptr(depth), _ptr(opacity), _ptr(depth_var), _ptr(rgb_var), _ptr(weights),
                                        _ptr(all_cum), _stream()), "composite_forward")
        ctx.white_bg = bool(white_bg)
        ctx.save_for_backward(sigma, rgb, t, dirs)
        ctx.mark_non_differentiable(depth_var, rgb_var, all_cum)
        # outputs nobody differentiated arrive as None in backward (the C ABI takes NULL) instead of as freshly
        # zero-filled tensors: six fill kernels and a [R,S] write + read less per render call
        ctx.set_materialize_grads(False)
        return rgb_map, depth, opacity, weights, depth_var, rgb_var, all_cum

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g_rgb, g_depth, g_opacity, g_weights, *_unused):
        L = _lib.lib()
        sigma, rgb, t, dirs = ctx.saved_tensors
        R, S = t.shape
        opt = lambda g: _f32c(g) if g is not None else None
        g_rgb, g_depth, g_opacity, g_weights = opt(g_rgb), opt(g_depth), opt(g_opacity), opt(g_weights)
        d_sigma = torch.empty_like(sigma)
        d_rgb = torch.empty_like(rgb)
        d_dirs = torch.zeros_like(dirs) if ctx.needs_input_grad[3] else None
        check(L.sparf_composite_backward(R, S, _ptr(sigma), _ptr(rgb), _ptr(t), _ptr(dirs), int(ctx.white_bg),
                                         _ptr(g_rgb), _ptr(g_depth), _ptr(g_opacity), _ptr(g_weights), _ptr(d_sigma),
                                         _ptr(d_rgb), _ptr(d_dirs), _stream()), "composite_backward")
        return d_sigma, d_rgb, None, d_dirs, None


def composite(sigma, rgb, t, dirs, white_bg=False):
    """-> rgb_map [R,3], depth [R], opacity [R], weights [R,S], depth_var [R], rgb_var [R], all_cumulated [R]."""
    return CompositeFunction.apply(sigma, rgb, t, dirs, bool(white_bg))


# ------------------------------------------------------------------------------------------------
# rays
# ------------------------------------------------------------------------------------------------
class RayGenFunction(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, pose_w2c, intr_inv, W: int, ray_idx, pixels):
        L = _lib.lib()
        pose_w2c, intr_inv = _f32c(pose_w2c), _f32c(intr_inv)
        B = pose_w2c.shape[0]
        if pixels is not None:
            pixels = _f32c(pixels)
            per_image = int(pixels.dim() == 3)
            n = pixels.shape[-2]
            idx = None
        else:
            idx = ray_idx.to(torch.int64).contiguous()
            per_image = int(idx.dim() == 2 and idx.shape[0] == B)
            n = idx.shape[-1]
        origins = torch.empty(B, n, 3, device=pose_w2c.device)
        dirs = torch.empty(B, n, 3, device=pose_w2c.device)
        check(L.sparf_raygen_forward(B, n, int(W), _ptr(pose_w2c), _ptr(intr_inv), _ptr(idx), _ptr(pixels), per_image,
                                     _ptr(origins), _ptr(dirs), _stream()), "raygen_forward")
        ctx.args = (B, n, int(W), per_image)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pose_w2c, intr_inv, idx if idx is not None else torch.empty(0), pixels if pixels is not None else torch.empty(0))
        ctx.has_idx = idx is not None
        return origins, dirs

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g_o, g_d):
        L = _lib.lib()
        pose_w2c, intr_inv, idx, pixels = ctx.saved_tensors
        B, n, W, per_image = ctx.args
        if g_o is None and g_d is None:
            return None, None, None, None, None
        d_pose = torch.zeros_like(pose_w2c)
        g_o = _f32c(g_o) if g_o is not None else None
        g_d = _f32c(g_d) if g_d is not None else None
        # pixel locations are differentiable in the reference (camera.py:400-416); the depth-consistency loss renders at
        # pixels projected from a rendered depth and back-propagates through them (depth_cons_loss.py:247-283)
        d_px = torch.zeros_like(pixels) if (not ctx.has_idx and ctx.needs_input_grad[4]) else None
        check(L.sparf_raygen_backward(B, n, W, _ptr(pose_w2c), _ptr(intr_inv), _ptr(idx if ctx.has_idx else None),
                                      _ptr(None if ctx.has_idx else pixels), per_image, _ptr(g_o), _ptr(g_d),
                                      _ptr(d_pose), _ptr(d_px), _stream()), "raygen_backward")
        return d_pose, None, None, None, d_px


_HOST_CACHE = {}


def cached(tag: str, tensor: torch.Tensor, fn):
    """Memoise a small derived quantity of a (device) tensor that rarely changes (intrinsics, depth range), keyed
    by storage address + in-place version, so that the hot loop neither recomputes it nor synchronises."""
    key = (tag, tensor.data_ptr(), tensor._version, tuple(tensor.shape), tensor.device)
    hit = _HOST_CACHE.get(key)
    if hit is None and tensor.is_cuda and torch.cuda.is_current_stream_capturing():
        return fn(tensor)          # graph-pool temporaries must not be memoised (their addresses are recycled)
    if hit is None:
        if len(_HOST_CACHE) > 256:
            _HOST_CACHE.clear()
        # the entry keeps `tensor` alive, so its address cannot be handed to another tensor while cached
        hit = (fn(tensor), tensor)
        _HOST_CACHE[key] = hit
    return hit[0]


def raygen(pose_w2c, intr, W: int, *, ray_idx=None, pixels=None):
    """pose_w2c [B,3,4], intr [B,3,3] -> (center, ray) [B,n,3] at ray_idx ((n,)/(B,n) int) or float pixels."""
    assert (ray_idx is None) != (pixels is None)
    # camera.py:318-319; intrinsics carry no gradient and are constant over training: invert once
    # (inv_ex: no host-side singularity check, i.e. no synchronisation)
    intr_inv = cached("Kinv", intr, lambda k: torch.linalg.inv_ex(k.detach().float()).inverse)
    return RayGenFunction.apply(pose_w2c, intr_inv, W, ray_idx, pixels)


# ------------------------------------------------------------------------------------------------
# sampling (no gradient)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
@_on_tensor_device
def sample_depth(R: int, S: int, near: float, rng: float, *, inverse=False, rand=None, far_per_ray=None, device=None):
    L = _lib.lib()
    device = device or (rand.device if rand is not None else far_per_ray.device)
    t = torch.empty(R, S, device=device, dtype=torch.float32)
    rand_c = _f32c(rand).reshape(R, S) if rand is not None else None
    far_c = _f32c(far_per_ray).reshape(R) if far_per_ray is not None else None
    check(L.sparf_sample_depth(R, S, float(near), float(rng), int(inverse), _ptr(rand_c), _ptr(far_c), _ptr(t), _stream()),
          "sample_depth")
    return t


@torch.no_grad()
@_on_tensor_device
def sample_pdf_merge(weights, t_coarse, u_mid, near: float, far: float):
    """weights, t_coarse [R,S]; u_mid [S_fine] -> (t_fine [R,S_fine], t_all [R,S+S_fine] ascending)."""
    L = _lib.lib()
    weights, t_coarse, u_mid = _f32c(weights), _f32c(t_coarse), _f32c(u_mid)
    R, S = weights.shape
    Sf = u_mid.numel()
    t_fine = torch.empty(R, Sf, device=weights.device)
    t_all = torch.empty(R, S + Sf, device=weights.device)
    check(L.sparf_sample_pdf_merge(R, S, Sf, float(near), float(far), _ptr(weights), _ptr(t_coarse), _ptr(u_mid),
                                   _ptr(t_fine), _ptr(t_all), _stream()), "sample_pdf_merge")
    return t_fine, t_all


# ------------------------------------------------------------------------------------------------
# photometric Huber loss
# ------------------------------------------------------------------------------------------------
class Huber2Function(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, pred, target):
        L = _lib.lib()
        pred_c, target_c = _f32c(pred), _f32c(target)
        loss = torch.zeros((), device=pred.device)
        d_pred = torch.empty_like(pred_c)
        check(L.sparf_huber2_fwd_bwd(pred_c.numel(), _ptr(pred_c), _ptr(target_c), 1.0, _ptr(loss), _ptr(d_pred), _stream()),
              "huber2")
        ctx.save_for_backward(d_pred)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        (d_pred,) = ctx.saved_tensors
        return (d_pred * g).view(ctx.shape), None


def huber2(pred, target):
    """2 * mean Huber(delta=0.5)(pred - target)   (base_losses.py:155-156)."""
    return Huber2Function.apply(pred, target.expand_as(pred))


# ------------------------------------------------------------------------------------------------
# stand-alone positional encoding (the MLP kernels fuse it; this is the tensor-level op)
# ------------------------------------------------------------------------------------------------
class PosEncFunction(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, L: int, barf_c2f, progress):
        lib = _lib.lib()
        x_c = _f32c(x)
        C = x_c.shape[-1]
        n = x_c.numel() // C
        out = torch.empty(*x_c.shape[:-1], 2 * C * L, device=x_c.device, dtype=torch.float32)
        use = barf_c2f is not None
        start, rng = (float(barf_c2f[0]), float(barf_c2f[1] - barf_c2f[0])) if use else (0.0, 1.0)
        prog = _f32c(progress.detach()).reshape(1) if use else None
        check(lib.sparf_posenc_forward(n, C, int(L), _ptr(x_c), int(use), start, rng, _ptr(prog), _ptr(out), _stream()), "posenc_forward")
        ctx.args = (n, C, int(L), use, start, rng)
        ctx.save_for_backward(x_c, prog if use else torch.empty(0))
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        lib = _lib.lib()
        x_c, prog = ctx.saved_tensors
        n, C, L, use, start, rng = ctx.args
        d_x = torch.empty_like(x_c)
        check(lib.sparf_posenc_backward(n, C, L, _ptr(x_c), int(use), start, rng, _ptr(prog if use else None), _ptr(_f32c(g)),
                                        _ptr(d_x), _stream()), "posenc_backward")
        return d_x, None, None, None


def posenc(x, L: int, barf_c2f=None, progress=None):
    """[..., C] -> [..., 2*C*L]: per channel L sines then L cosines of x * 2^j * pi, times the BARF coarse-to-fine weight of
    band j when barf_c2f = (start, end) is given (frequency_nerf.py:47-69, 248-257).  Differentiable w.r.t. x."""
    return PosEncFunction.apply(x, L, tuple(barf_c2f) if barf_c2f is not None else None, progress)


# ------------------------------------------------------------------------------------------------
# distortion regulariser (default off in the reference's configs)
# ------------------------------------------------------------------------------------------------
class DistortionFunction(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, t, w):
        L = _lib.lib()
        t_c, w_c = _f32c(t), _f32c(w)
        S = t_c.shape[-2] if t_c.shape[-1] == 1 else t_c.shape[-1]
        R = t_c.numel() // S
        loss = torch.zeros((), device=t.device)
        d_w = torch.empty_like(w_c)
        need_t = ctx.needs_input_grad[0]
        d_t = torch.empty_like(t_c) if need_t else None
        check(L.sparf_distortion_fwd_bwd(R, S, _ptr(t_c), _ptr(w_c), 1.0, _ptr(loss), _ptr(d_w), _ptr(d_t), _stream()),
              "distortion")
        ctx.save_for_backward(d_w, d_t if need_t else d_w)
        ctx.need_t = need_t
        ctx.shapes = (t.shape, w.shape)
        return loss

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        d_w, d_t = ctx.saved_tensors
        return ((d_t * g).view(ctx.shapes[0]) if ctx.need_t else None), (d_w * g).view(ctx.shapes[1])


def distortion_loss(t, w):
    """mip-NeRF-360 distortion loss of the renderer's `t`, `weights` [B,R,S,1] (regularization_losses.py:20-48), O(S)."""
    return DistortionFunction.apply(t, w)
